"""Parity of the CUDA path (through the C ABI) with the golden vectors of the unmodified reference
and with the oracle.  Bit-exact for tag offsets, PDU order and bits, carried state; SNR within
1e-4 dB (BASELINE.json north_star), nan/inf matching as such."""
import numpy as np
import pytest

import devmem

import golden_io as G
import gr_adsb_b200 as g
from gr_adsb_b200 import synth
from harness import run_blocks
from oracle import adsb_oracle as ao
from oracle import c_oracle as co

pytestmark = pytest.mark.gpu

_ctx_cache = {}


def ctx_for(sps, n):
    key = sps
    c = _ctx_cache.get(key)
    if c is None or c.max_items < n:
        if c is not None:
            c.close()
        c = g.Context(sps, max(n, 1 << 20))
        _ctx_cache[key] = c
    c.reset_state()
    return c


# ---------------------------------------------------------------------------------- blocks, per call
@pytest.mark.parametrize("name,ci", G.all_chunkings())
def test_blocks_match_reference_per_call(name, ci):
    c = G.Case(name)
    gold = c.chunkings[ci]
    fr = g.framer(c.fs, c.threshold)
    dm = g.demod(c.fs)
    r = run_blocks(fr, dm, c.x, gold.sizes, gold.thresholds)
    assert [t[0] for t in r["tags"]] == list(gold.tag_off)
    assert G.snr_close([t[1] for t in r["tags"]], gold.tag_snr)
    assert [p[0] for p in r["pdus"]] == list(gold.pdu_off)
    assert G.snr_close([p[1] for p in r["pdus"]], gold.pdu_snr)
    bits = np.array([p[2] for p in r["pdus"]], np.uint8).reshape(-1, 112)
    assert np.array_equal(bits, gold.pdu_bits)
    assert [s[1] for s in r["state"]] == list(gold.state_eob)
    assert np.array_equal(np.array([s[0] for s in r["state"]], np.float32).view(np.uint32),
                          gold.state_in0.view(np.uint32))


# ---------------------------------------------------------------------------------- fused, whole buffer
def check_fused(hits, sps, tag_off, tag_snr, pdu_off, pdu_bits, pdu_crc):
    H1 = 8 * sps - 1
    assert np.array_equal(hits["idx"] - H1, tag_off)
    assert G.snr_close(hits["snr"], tag_snr)
    valid = (hits["flags"] & g.HIT_DEMOD_VALID) != 0
    assert np.array_equal(hits["idx"][valid] - H1, pdu_off)
    assert np.array_equal(g.unpack_bits(hits[valid]), pdu_bits)
    assert np.array_equal(((hits["flags"][valid] & g.HIT_CRC_OK) != 0).astype(np.uint8), pdu_crc)
    assert not np.any(hits["flags"][~valid] & g.HIT_CRC_OK)


@pytest.mark.parametrize("name", G.case_names())
def test_fused_whole_buffer_matches_reference(name):
    c = G.Case(name)
    whole = [ch for ch in c.chunkings if len(ch.sizes) == 1]
    if not whole:
        pytest.skip("no whole-buffer chunking in this case")
    gold = whole[0]
    ctx = ctx_for(c.sps, len(c.x))
    hits = ctx.burst_work(c.x, threshold=c.threshold)
    check_fused(hits, c.sps, gold.tag_off, gold.tag_snr, gold.pdu_off, gold.pdu_bits, gold.pdu_crc)
    st = ctx.get_state()
    assert st[1] == gold.state_eob[-1]
    assert np.float32(st[0]).view(np.uint32) == gold.state_in0[-1].view(np.uint32)
    if c.iq is not None:
        ctx.reset_state()
        hits = ctx.burst_work(c.iq, threshold=c.threshold)
        check_fused(hits, c.sps, gold.tag_off, gold.tag_snr, gold.pdu_off, gold.pdu_bits, gold.pdu_crc)


# ---------------------------------------------------------------------------------- vs the C oracle, larger
def _c_oracle_check(iq, sps, thr, ctx, use_iq=True):
    x = co.mag2(iq)
    ref = co.run_stream(x, sps * 1e6, thr, None)
    ctx.reset_state()
    hits = ctx.burst_work(iq if use_iq else x, threshold=thr)
    crc = np.array([co.crc_pass(b) for b in ref["pdu_bits"]], np.uint8)
    check_fused(hits, sps, ref["tag_off"], ref["tag_snr"], ref["pdu_off"], ref["pdu_bits"], crc)
    assert ctx.get_state()[1] == ref["state"][1]
    return len(hits)


@pytest.mark.parametrize("sps,seed,noise,thr,dens", [
    (2, 11, 1e-3, 0.01, 5e-3), (2, 12, 1e-3, 0.01, 5e-4), (2, 13, 6e-3, 0.01, 2e-3),
    (2, 14, 1e-3, 3e-4, 1e-3), (2, 15, 1e-3, 0.05, 5e-2), (4, 16, 1e-3, 0.01, 2e-3),
    (6, 17, 2e-3, 0.01, 1e-3), (8, 18, 1e-3, 0.01, 1.25e-3), (8, 19, 1e-3, 2e-4, 5e-4),
    (16, 20, 1e-3, 0.01, 5e-4), (2, 21, 1e-3, 1e-5, 1e-3),
])
def test_fused_matches_c_oracle_medium(sps, seed, noise, thr, dens):
    n = (1 << 21) * (2 if sps >= 8 else 1)
    d = synth.synth_numpy(n, sps=sps, n_bursts=int(n * dens), seed=seed, noise_var=noise, amp_db=(-25.0, 0.0))
    ctx = ctx_for(sps, n)
    nh = _c_oracle_check(d["iq"], sps, thr, ctx, use_iq=(seed % 2 == 1))
    assert nh > 0 or thr < 1e-4


def test_long_runs_slow_path():
    """Runs of 32..5000 above-threshold samples (bitmap look-back exhausted -> global-memory path),
    some with a preamble-shaped centre."""
    rng = np.random.default_rng(5)
    n = 1 << 18
    x = rng.exponential(1e-3, n).astype(np.float32)
    pos = 500
    while pos < n - 12000:
        L = int(rng.integers(30, 5000))
        x[pos:pos + L] = np.maximum(x[pos:pos + L], 0.011)
        p = pos + L // 2                        # centre of the run: plant a preamble shape around it
        a = np.float32(rng.uniform(0.2, 1.0))
        x[p:p + 16] = 0.011
        for k in (0, 2, 7, 9):
            x[p + k] = a
        x[pos + L] = 0.0
        pos += L + int(rng.integers(50, 400))
    ref = co.run_stream(x, 2e6, 0.01, None)
    assert len(ref["tag_off"]) > 5
    ctx = ctx_for(2, n)
    hits = ctx.burst_work(x, threshold=0.01)
    crc = np.array([co.crc_pass(b) for b in ref["pdu_bits"]], np.uint8)
    check_fused(hits, 2, ref["tag_off"], ref["tag_snr"], ref["pdu_off"], ref["pdu_bits"], crc)
    for chunk in (8192, 3000):
        a = co.run_stream(x, 2e6, 0.01, chunk)
        fr, dm = g.framer(2e6, 0.01), g.demod(2e6)
        r = run_blocks(fr, dm, x, ao.chunk_sizes(n, chunk))
        assert [t[0] for t in r["tags"]] == list(a["tag_off"])
        assert [p[0] for p in r["pdus"]] == list(a["pdu_off"])
        assert fr.prev_eob_idx == a["state"][1]


def test_demod_work_straddle_and_empty():
    ctx = ctx_for(2, 1 << 16)
    x = np.random.default_rng(1).random(5000).astype(np.float32)
    bits, valid = ctx.demod_work(x, np.array([], np.int64))
    assert bits.shape == (0, 112) and valid.shape == (0,)
    sob = np.array([16, 4000, 5000 - 223, 5000 - 224, 4999], np.int64)
    bits, valid = ctx.demod_work(x, sob)
    assert list(valid) == [1, 1, 0, 1, 0]              # eob_idx = sob + 223 must be < len (demod.py:82)
    for k in np.flatnonzero(valid):
        assert np.array_equal(bits[k], ao.demod_slice(x, int(sob[k]), 2))


def test_device_resident_entry_point():
    torch = pytest.importorskip("torch")
    n, sps, thr = 1 << 20, 2, 0.01
    d = synth.synth_numpy(n, sps=sps, n_bursts=2000, seed=77)
    ref = co.run_stream(co.mag2(d["iq"]), sps * 1e6, thr, None)
    t = devmem.to_dev(torch.from_numpy(d["iq"].view(np.float32).reshape(-1, 2)))
    ctx = ctx_for(sps, n)
    s = devmem.stream()
    for _ in range(2):          # second call re-uses every buffer (epoch handling)
        ctx.burst_work_dev(t.data_ptr(), 1, n, n, 8 * sps - 1, thr, base=0, eob_in=-1, stream=s)
        devmem.sync()
        cnt, hits = ctx.burst_finish(cap=len(ref["tag_off"]) + 8)
        crc = np.array([co.crc_pass(b) for b in ref["pdu_bits"]], np.uint8)
        check_fused(hits, sps, ref["tag_off"], ref["tag_snr"], ref["pdu_off"], ref["pdu_bits"], crc)
        assert cnt["n_hits"] == len(hits) and cnt["n_crc_ok"] == int(crc.sum())
        assert cnt["n_demod"] == len(ref["pdu_off"])


def test_full_size_properties():
    """BASELINE config-2 shape at 2^26 samples, generated in HBM: every injected burst that is clear of
    its neighbours must come back with exactly its bits and a passing CRC; hits are ordered and obey
    the 63-symbol blanking; all of it twice (determinism); and the ~100 000 hits equal the C oracle's exactly."""
    torch = pytest.importorskip("torch")
    n, sps, thr = 1 << (21 if devmem.EMU else 26), 2, 0.01      # the emulation runs ~0.4 Msamples/s
    nb = int(n * 5e-3)
    iq, truth = synth.synth_torch(n, sps, nb, 20260925, devmem.DEVICE)
    ctx = g.Context(sps, n)
    s = devmem.stream()
    outs = []
    for _ in range(2):
        ctx.burst_work_dev(iq.data_ptr(), 1, n, n, 8 * sps - 1, thr, stream=s)
        devmem.sync()
        cnt, hits = ctx.burst_finish(cap=n // 126 + 8)
        outs.append((cnt, hits.copy()))
    assert outs[0][0] == outs[1][0] and np.array_equal(outs[0][1], outs[1][1])
    cnt, hits = outs[0]
    idx = hits["idx"] - (8 * sps - 1)
    assert np.all(np.diff(idx) > 63 * sps)
    starts = truth["starts"].cpu().numpy()
    bits = truth["bits"].cpu().numpy()
    amp2 = truth["amp2"].cpu().numpy()
    gap_prev = np.diff(starts, prepend=-10**9)
    gap_next = np.diff(starts, append=10**12)
    clear = (gap_prev > 63 * sps + 260) & (gap_next > 250) & (amp2 > 0.15)
    pos = np.searchsorted(idx, starts[clear])
    found = (pos < len(idx)) & (idx[np.minimum(pos, len(idx) - 1)] == starts[clear])
    assert found.mean() > 0.999
    got = g.unpack_bits(hits[pos[found]])
    same = np.all(got == bits[clear][found], axis=1)
    assert same.mean() > 0.98           # noise can still flip a bit of a -8 dB burst
    ok = (hits["flags"][pos[found]] & g.HIT_CRC_OK) != 0
    assert np.array_equal(ok, same | ok) and ok.sum() >= same.sum()
    assert cnt["n_crc_ok"] >= int(same.sum())
    # and exact equality with the oracle at this size: every tag offset, SNR, PDU and parity verdict of the whole buffer
    x = co.mag2(iq.cpu().numpy().view(np.complex64).reshape(-1))
    ref = co.run_stream(x, sps * 1e6, thr, None)
    H1 = 8 * sps - 1
    assert np.array_equal(hits["idx"] - H1, ref["tag_off"])
    a, b = hits["snr"], ref["tag_snr"]
    fin = np.isfinite(b)
    assert np.array_equal(np.isnan(a), np.isnan(b)) and np.array_equal(np.isinf(a), np.isinf(b))
    assert np.all(np.abs(a[fin] - b[fin]) <= 1e-4)          # BASELINE.json north_star: SNR within 1e-4 dB
    valid = (hits["flags"] & g.HIT_DEMOD_VALID) != 0
    assert np.array_equal(hits["idx"][valid] - H1, ref["pdu_off"])
    assert np.array_equal(g.unpack_bits(hits[valid]), ref["pdu_bits"])
    assert cnt["n_crc_ok"] == co.crc_pass_count(ref["pdu_bits"]) and cnt["n_demod"] == len(ref["pdu_off"])
    ctx.close()


def test_host_path_multi_chunk_pipeline():
    """adsb_burst_work copies the host buffer in 2^22-item chunks and launches the scan for each chunk as it
    lands: a stream spanning several chunks (ragged tail) must equal the oracle exactly."""
    n, sps, thr = (1 << 23) + 12345, 2, 0.01
    d = synth.synth_numpy(n, sps=sps, n_bursts=int(n * 3e-3), seed=4242)
    ctx = ctx_for(sps, n)
    _c_oracle_check(d["iq"], sps, thr, ctx, use_iq=True)
    _c_oracle_check(d["iq"], sps, thr, ctx, use_iq=False)


def test_threshold_sweep_same_context():
    """set_threshold between calls on one context: results follow the threshold of each call."""
    n, sps = 1 << 20, 2
    d = synth.synth_numpy(n, sps=sps, n_bursts=3000, seed=99, amp_db=(-30.0, 0.0))
    ctx = ctx_for(sps, n)
    for thr in (0.05, 0.002, 0.3, 0.01):
        _c_oracle_check(d["iq"], sps, thr, ctx)


def _run_burst_block(blk, x, sizes, thresholds=None):
    """Scheduler stand-in for the fused block: history, counters, tag + PDU collection."""
    H = blk.N_hist
    dt = x.dtype
    hist = np.zeros(H - 1, dtype=dt)
    pos = 0
    tags, pdus = [], []
    blk.start_timestamp = 0.0
    for k, n in enumerate(sizes):
        if thresholds is not None:
            blk.set_threshold(thresholds[k])
        in0 = np.concatenate((hist, x[pos:pos + n]))
        out0 = np.empty(n, dtype=np.float32)
        blk._nitems_written = pos
        blk._nitems_read = pos
        nt0, nm0 = len(blk._tags_out), len(blk._msgs["demodulated"])
        assert blk.work([in0], [out0]) == n
        # output port: the |IQ|^2 stream (complex input) / the input itself (framer.py:181), bit for bit
        want = ao.mag2(x[pos:pos + n]) if dt == np.complex64 else x[pos:pos + n]
        assert np.array_equal(out0.view(np.uint32), np.ascontiguousarray(want).view(np.uint32))
        for t in blk._tags_out[nt0:]:
            assert t.key == "burst" and t.value[0] == "SOB" and t.srcid == "framer"
            tags.append((int(t.offset), t.value[1]))
        for meta, vec in blk._msgs["demodulated"][nm0:]:
            pdus.append((int(round(meta["timestamp"] * blk.fs)), float(meta["snr"]), np.array(vec)))
        hist = in0[len(in0) - (H - 1):].copy()
        pos += n
    return tags, pdus


@pytest.mark.parametrize("name,ci", G.all_chunkings())
def test_fused_block_matches_reference_chain(name, ci):
    """gr_adsb_b200.burst == reference framer -> demod in lock step, for every golden chunking.  (A late tag --
    offset before the chunk -- is attached by the framer but never demodulated: same here.)"""
    c = G.Case(name)
    gold = c.chunkings[ci]
    blk = g.burst(c.fs, c.threshold)
    tags, pdus = _run_burst_block(blk, c.x, gold.sizes, gold.thresholds)
    assert [t[0] for t in tags] == list(gold.tag_off)
    assert G.snr_close([t[1] for t in tags], gold.tag_snr)
    assert [p[0] for p in pdus] == list(gold.pdu_off)
    bits = np.array([p[2] for p in pdus], np.uint8).reshape(-1, 112)
    assert np.array_equal(bits, gold.pdu_bits)
    assert blk.prev_eob_idx == (gold.state_eob[-1] if len(gold.state_eob) else -1)
    if c.iq is not None:
        blk = g.burst(c.fs, c.threshold, complex_input=True)
        tags2, pdus2 = _run_burst_block(blk, c.iq, gold.sizes, gold.thresholds)
        assert [t[0] for t in tags2] == list(gold.tag_off)
        assert [p[0] for p in pdus2] == list(gold.pdu_off)
        assert np.array_equal(np.array([p[2] for p in pdus2], np.uint8).reshape(-1, 112), gold.pdu_bits)


def test_conservative_error_correction_matches_reference_decoder():
    """the 600 golden messages (no error / 1 bit / 2 contiguous / 2 scattered / wrong DF / 3 flips)
    are modulated as clean, well separated bursts, so the demodulator returns exactly the corrupted bits; with
    ADSB_FEC_CONSERVATIVE the hits must carry the bits and the verdict the reference decoder
    (error_corr="Conservative") arrives at."""
    bits_in, passed, bits_out = G.fec_vectors()
    nb, sps = len(bits_in), 2
    n = nb * 720 + 4096
    starts = 500 + 720 * np.arange(nb)
    d = synth.synth_numpy(n, sps=sps, seed=606, bits=bits_in, starts=starts, noise_var=1e-5, amp_db=(-3.0, 0.0))
    ctx = g.Context(sps, n)
    ctx.set_error_correction(g.FEC_CONSERVATIVE)
    hits = ctx.burst_work(d["iq"], threshold=0.01)
    assert np.array_equal(hits["idx"] - 15, starts)
    got = g.unpack_bits(hits)
    assert np.array_equal(got, bits_out)
    ok = ((hits["flags"] & g.HIT_CRC_OK) != 0).astype(np.uint8)
    assert np.array_equal(ok, passed)
    fixed = (hits["flags"] & g.HIT_FEC_FIXED) != 0
    assert np.array_equal(fixed, np.any(bits_in != bits_out, axis=1))
    c = ctx.counters()
    assert c["n_fec_fixed"] == int(fixed.sum()) and c["n_crc_ok"] == int(passed.sum()) - int(fixed.sum())
    # default mode leaves the bits alone
    ctx.set_error_correction(g.FEC_NONE)
    ctx.reset_state()
    hits0 = ctx.burst_work(d["iq"], threshold=0.01)
    assert np.array_equal(g.unpack_bits(hits0), bits_in)
    assert not np.any(hits0["flags"] & g.HIT_FEC_FIXED)
    ctx.close()


@pytest.mark.parametrize("sps,seed", [(2, 31), (8, 32), (6, 33)])
def test_sc16_input_matches_oracle(sps, seed):
    """interleaved int16 IQ widened on the device (value = int * 2^-15) must give exactly what
    the complex64 path gives on the widened samples."""
    n = (1 << 20) * (2 if sps >= 6 else 1)
    d = synth.synth_numpy(n, sps=sps, n_bursts=int(n * 2e-3 * 2 / sps), seed=seed, noise_var=1e-3, amp_db=(-20.0, 0.0))
    q = np.clip(np.round(d["iq"].view(np.float32).reshape(-1, 2) * 32768.0), -32768, 32767).astype(np.int16)
    iq = (q.astype(np.float32) * np.float32(2.0 ** -15)).view(np.complex64).reshape(-1)
    ref = co.run_stream(co.mag2(iq), sps * 1e6, 0.01, None)
    ctx = ctx_for(sps, n)
    hits = ctx.burst_work(q, threshold=0.01)
    crc = np.array([co.crc_pass(b) for b in ref["pdu_bits"]], np.uint8)
    check_fused(hits, sps, ref["tag_off"], ref["tag_snr"], ref["pdu_off"], ref["pdu_bits"], crc)
    assert len(hits) > 20

