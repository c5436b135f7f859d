"""Streaming session (adsb_stream_*): a stream pushed in arbitrary pieces through the pinned ring, scanned chunk by chunk
with the neighbouring chunks read in place, must give exactly what ONE work() call over the whole stream gives."""
import numpy as np
import pytest

import gr_adsb_b200 as g
from gr_adsb_b200 import synth
from oracle import c_oracle as co

pytestmark = pytest.mark.gpu


def run_stream(buf, fmt, sps, thr, chunk_items, pushes, in_place=False):
    ctx = g.Context(sps, max(chunk_items, 1 << 16))
    ctx.stream_open(fmt, thr, chunk_items)
    pos, got = 0, []
    rng = np.random.default_rng(len(buf) + chunk_items)
    while pos < len(buf):
        n = min(len(buf) - pos, int(pushes) if not isinstance(pushes, tuple) else int(rng.integers(pushes[0], pushes[1])))
        if in_place:
            done = 0
            while done < n:
                dst = ctx.stream_buffer()
                m = min(n - done, len(dst))
                dst[:m] = buf[pos + done: pos + done + m]
                ctx.stream_commit(m)
                done += m
        else:
            ctx.stream_push(buf[pos:pos + n])
        pos += n
        got.append(ctx.stream_pop())
    ctx.stream_flush()
    got.append(ctx.stream_pop())
    st = ctx.stream_stats()
    hits = np.concatenate(got)
    assert st["items"] == len(buf) and st["hits"] == len(hits)
    ctx.stream_close()
    ctx.close()
    return hits, st


def check(hits, x, sps, thr):
    ref = co.run_stream(x, sps * 1e6, thr, None)
    H1 = 8 * sps - 1
    assert np.array_equal(hits["idx"] - H1, ref["tag_off"])
    a, b = hits["snr"], ref["tag_snr"]
    fin = np.isfinite(b)
    assert np.array_equal(np.isnan(a), np.isnan(b)) and np.all(np.abs(a[fin] - b[fin]) <= 1e-4)
    valid = (hits["flags"] & g.HIT_DEMOD_VALID) != 0
    assert np.array_equal(hits["idx"][valid] - H1, ref["pdu_off"])
    assert np.array_equal(g.unpack_bits(hits[valid]), ref["pdu_bits"])
    assert int(((hits["flags"] & g.HIT_CRC_OK) != 0).sum()) == co.crc_pass_count(ref["pdu_bits"])


@pytest.mark.parametrize("sps,fmt,n,chunk,pushes,in_place", [
    (2, g.FMT_C64, 200000, 16384, 8192, False),            # pushes smaller than a chunk
    (2, g.FMT_C64, 200000, 16384, 50000, False),           # pushes larger than a chunk
    (2, g.FMT_F32, 150001, 20000, (1, 30000), False),      # ragged pushes, ragged tail
    (8, g.FMT_C64, 300000, 65536, (1000, 90000), True),    # filled in place
    (4, g.FMT_F32, 131072, 32768, 32768, False),           # stream length a multiple of the chunk
    (2, g.FMT_C64, 100000, 1 << 17, 7777, False),          # the whole stream inside one chunk
    (2, g.FMT_F32, 65536 + 7, 65536, 65536, False),        # a tail shorter than the 8*sps-1 look-ahead
    (6, g.FMT_C64, 120000, 30000, (500, 40000), False)])   # unspecialised sample rate
def test_stream_equals_whole_buffer(sps, fmt, n, chunk, pushes, in_place):
    d = synth.synth_numpy(n, sps=sps, n_bursts=int(n * 2e-3 * 2 / sps), seed=700 + sps + n % 13, amp_db=(-20.0, 0.0))
    buf = d["iq"] if fmt == g.FMT_C64 else synth.mag2(d["iq"])
    hits, st = run_stream(buf, fmt, sps, 0.01, chunk, pushes, in_place)
    assert len(hits) > 20
    check(hits, synth.mag2(d["iq"]), sps, 0.01)


@pytest.mark.parametrize("fmt", [g.FMT_SC16, g.FMT_CU8])
def test_stream_integer_formats(fmt):
    n = 180000
    d = synth.synth_numpy(n, sps=2, n_bursts=400, seed=42)
    ri = d["iq"].view(np.float32).reshape(-1, 2)
    if fmt == g.FMT_SC16:
        raw = np.clip(np.round(ri * 32768.0), -32768, 32767).astype(np.int16)
        v = raw.astype(np.float32) * np.float32(2.0 ** -15)
    else:
        raw = np.clip(np.round(ri * 128.0 + 127.5), 0, 255).astype(np.uint8)
        v = (raw.astype(np.float32) - np.float32(127.5)) * np.float32(1.0 / 128)
    x = (v[:, 0] * v[:, 0] + v[:, 1] * v[:, 1]).astype(np.float32)
    hits, st = run_stream(raw, fmt, 2, 0.02, 40000, (100, 50000))
    check(hits, x, 2, 0.02)


def test_stream_carries_blanking_and_long_runs_across_chunks():
    """Preamble-only pulses every ~100 samples (every chunk's first candidates are blanked by the previous chunk's last
    hit) and an above-threshold run of 3000 samples over a chunk boundary."""
    n, chunk = 100000, 20000
    rng = np.random.default_rng(5)
    x = rng.exponential(1e-4, n).astype(np.float32)
    pos = 300
    while pos < n - 300:
        a = np.float32(rng.uniform(0.1, 1.0))
        for hs in (0, 2, 7, 9):
            x[pos + hs] += a
        pos += int(rng.integers(40, 120))
    x[2 * chunk - 2000: 2 * chunk + 1000] = 0.3
    hits, st = run_stream(x, g.FMT_F32, 2, 0.01, chunk, 9999)
    check(hits, x, 2, 0.01)


def test_stream_errors_and_reuse():
    ctx = g.Context(2, 1 << 16)
    with pytest.raises(g.AdsbError):
        ctx.stream_open(g.FMT_C64, 0.01, 1 << 17)          # chunk larger than max_items
    with pytest.raises(g.AdsbError):
        ctx.stream_open(g.FMT_C64, 0.01, 100)              # shorter than a burst
    with pytest.raises(g.AdsbError):
        ctx._stream_fmt = g.FMT_C64
        ctx.stream_push(np.zeros(10, np.complex64))        # not open
    d = synth.synth_numpy(50000, sps=2, n_bursts=100, seed=1)
    for _ in range(2):                                     # a second session on the same context starts from scratch
        ctx.stream_open(g.FMT_C64, 0.01, 1 << 14)
        ctx.stream_push(d["iq"])
        ctx.stream_flush()
        hits = ctx.stream_pop()
        check(hits, synth.mag2(d["iq"]), 2, 0.01)
        with pytest.raises(g.AdsbError):
            ctx.stream_push(d["iq"][:10])                  # after flush
    # the block entry points still work on this context afterwards
    ctx.stream_close()
    h = ctx.burst_work(d["iq"], threshold=0.01)
    assert np.array_equal(h["idx"], hits["idx"])
    ctx.close()


def test_stream_with_error_correction_equals_fused_call():
    """The conservative error correction (decoder.py:738-763) applied chunk by chunk gives the same records as one fused call."""
    n = 150000
    d = synth.synth_numpy(n, sps=2, n_bursts=300, seed=77, amp_db=(-22.0, -10.0))      # weak bursts: some single-bit errors
    a = g.Context(2, n)
    a.set_error_correction(g.FEC_CONSERVATIVE)
    whole = a.burst_work(d["iq"], threshold=0.004).copy()
    b = g.Context(2, 1 << 15)
    b.set_error_correction(g.FEC_CONSERVATIVE)
    b.stream_open(g.FMT_C64, 0.004, 1 << 15)
    b.stream_push(d["iq"])
    b.stream_flush()
    got = b.stream_pop()
    assert len(whole) > 50 and whole.tobytes() == got.tobytes()
    assert int(((got["flags"] & g.HIT_FEC_FIXED) != 0).sum()) > 0
    a.close()
    b.close()
