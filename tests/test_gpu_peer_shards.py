"""Multi-GPU peer mode (adsb_shard_*) checked on ONE device: `world` contexts in one process, every rank's shard in
its own symmetric allocation, the neighbours mapped by pointer (adsb_shard_attach_ptrs) -- on the B200 box the scan
kernel's edge tiles then read the neighbouring allocations exactly as they read a peer GPU's memory over NVLink.
The phases of a step are interleaved across the ranks (one stream cannot hold two ranks that wait for each other).
The concatenated hits must equal the whole-buffer oracle."""
import ctypes as C

import numpy as np
import pytest

import devmem

import gr_adsb_b200 as g
from gr_adsb_b200 import sharding, synth
from oracle import c_oracle as co

pytestmark = pytest.mark.gpu

PHASES = (g.SHARD_PH_SIGNAL, g.SHARD_PH_SCAN, g.SHARD_PH_EXPORT, g.SHARD_PH_PLAN, g.SHARD_PH_CHECK)


def to_fmt(iq, fmt):
    if fmt == g.FMT_C64:
        return np.ascontiguousarray(iq)
    if fmt == g.FMT_F32:
        return co.mag2(iq)
    raise ValueError(fmt)


def run_peer(buf, fmt, sps, thr, world, max_rounds=None):
    n = len(buf)
    L = n // world
    assert L * world == n
    isz = buf.dtype.itemsize * (buf.shape[1] if buf.ndim == 2 else 1)      # bytes per item (int16 / uint8 IQ: two components)
    ctxs = [g.Context(sps, L) for _ in range(world)]
    ptrs = []
    for r in range(world):
        ptr, handle = ctxs[r].shard_alloc(L * isz)
        assert len(handle) == 64
        ptrs.append(ptr)
        part = np.ascontiguousarray(buf[r * L:(r + 1) * L])
        ctxs[r].shard_write(part.ctypes.data, L * isz)
    for r in range(world):
        ctxs[r].shard_attach_ptrs(r, world, ptrs)
    s = devmem.stream()
    for ph in PHASES:
        for r in range(world):
            ctxs[r].shard_step(fmt, L, thr, phases=ph, stream=s)
    rounds = 1
    while True:
        devmem.sync()
        unstable = [ctxs[r].burst_finish(fetch_hits=False)[0]["carry_unstable"] for r in range(world)]
        assert len(set(unstable)) == 1          # every rank evaluates the same plan
        if not unstable[0]:
            break
        rounds += 1
        assert rounds <= (max_rounds or world + 1)
        for ph in (g.SHARD_PH_EXPORT, g.SHARD_PH_PLAN, g.SHARD_PH_CHECK):
            for r in range(world):
                ctxs[r].shard_step(fmt, L, thr, phases=ph, stream=s)
    hits = []
    for r in range(world):
        c, h = ctxs[r].burst_finish(cap=L // (63 * sps) + 8)
        lo, hi = sharding.peer_window(r, world, L, sps)
        assert len(h) == 0 or (h["idx"][0] >= lo - 64 * sps and h["idx"][-1] < hi)
        hits.append(h.copy())
        ctxs[r].close()
    return np.concatenate(hits), rounds


def check_against_oracle(hits, x, sps, thr):
    ref = co.run_stream(x, sps * 1e6, thr, None)
    H1 = 8 * sps - 1
    assert np.array_equal(hits["idx"] - H1, ref["tag_off"])
    a, b = hits["snr"], ref["tag_snr"]
    fin = np.isfinite(b)
    assert np.array_equal(np.isnan(a), np.isnan(b)) and np.array_equal(np.isinf(a), np.isinf(b))
    assert np.all(np.abs(a[fin] - b[fin]) <= 1e-4)
    valid = (hits["flags"] & g.HIT_DEMOD_VALID) != 0
    assert np.array_equal(hits["idx"][valid] - H1, ref["pdu_off"])
    assert np.array_equal(g.unpack_bits(hits[valid]), ref["pdu_bits"])
    assert int(((hits["flags"] & g.HIT_CRC_OK) != 0).sum()) == co.crc_pass_count(ref["pdu_bits"])


@pytest.mark.parametrize("sps,world,seed,fmt", [(2, 1, 0, g.FMT_C64), (2, 2, 1, g.FMT_C64), (2, 4, 2, g.FMT_C64),
                                                 (2, 8, 3, g.FMT_C64), (8, 2, 4, g.FMT_C64), (8, 4, 5, g.FMT_F32),
                                                 (4, 3, 6, g.FMT_C64), (6, 2, 7, g.FMT_C64), (2, 8, 8, g.FMT_F32)])
def test_peer_shards_equal_whole_buffer(sps, world, seed, fmt):
    L = 24576 * sps // 2
    n = L * world
    # dense enough that bursts (and blanking intervals) straddle every shard boundary
    d = synth.synth_numpy(n, sps=sps, n_bursts=int(n * 2e-3), seed=300 + seed, amp_db=(-20.0, 0.0))
    hits, rounds = run_peer(to_fmt(d["iq"], fmt), fmt, sps, 0.01, world)
    check_against_oracle(hits, co.mag2(d["iq"]), sps, 0.01)


def _preamble_train(n, seed):
    rng = np.random.default_rng(seed)
    x = rng.exponential(1e-4, n).astype(np.float32)
    pos = 300
    while pos < n - 300:
        a = np.float32(rng.uniform(0.1, 1.0))
        for hs in (0, 2, 7, 9):
            x[pos + hs] += a
        pos += int(rng.integers(40, 120))
    return x


def test_peer_carry_cascade():
    """Preamble-only pulses every ~100 samples: every shard's first candidates sit inside the blanking interval of
    the previous shard's last hit, so the carry fix-up must run and cascade through the shards."""
    sps, world, L = 2, 4, 8192
    x = _preamble_train(L * world, 9)
    hits, rounds = run_peer(x, g.FMT_F32, sps, 0.01, world)
    check_against_oracle(hits, x, sps, 0.01)


def test_peer_long_run_across_the_cut():
    """An above-threshold run far longer than any halo that crosses a shard boundary: the owner of its falling edge
    walks back into the left neighbour's memory (halo-copy sharding returned ADSB_E_HALO here)."""
    sps, world, L = 2, 4, 16384
    n = L * world
    d = synth.synth_numpy(n, sps=sps, n_bursts=int(n * 2e-3), seed=77, amp_db=(-20.0, 0.0))
    x = co.mag2(d["iq"]).copy()
    x[L - 900: L + 37] = 0.5                     # 937-sample run over the first cut: centre in shard 0, fall in shard 1
    x[2 * L - 5: 2 * L + 3] = 0.25               # short run over the second cut
    x[3 * L - 2000: 3 * L] = 0.75                # run whose first sample below threshold is shard 3's first sample
    # make the middle of the long run look like a preamble start so that its centre is a candidate worth slicing
    hits, rounds = run_peer(x, g.FMT_F32, sps, 0.01, world)
    check_against_oracle(hits, x, sps, 0.01)


def test_peer_threshold_below_noise_floor():
    """Threshold inside the noise: thousands of pulses and candidates per shard, runs everywhere across the cuts."""
    sps, world, L = 2, 4, 16384
    n = L * world
    d = synth.synth_numpy(n, sps=sps, n_bursts=int(n * 1e-3), seed=78)
    x = co.mag2(d["iq"])
    hits, rounds = run_peer(x, g.FMT_F32, sps, 0.0015, world)
    check_against_oracle(hits, x, sps, 0.0015)


def test_peer_step_repeats_with_graph_flag():
    """The same step issued twice (second time asking for graph replay; the emulation ignores the flag) gives the same hits."""
    sps, world, L = 2, 1, 1 << 16
    d = synth.synth_numpy(L, sps=sps, n_bursts=200, seed=5)
    ctx = g.Context(sps, L)
    ptr, _ = ctx.shard_alloc(L * 8)
    ctx.shard_write(d["iq"].ctypes.data, L * 8)
    ctx.shard_attach_ptrs(0, 1, [ptr])
    s = devmem.stream()
    out = []
    for use_graph in (False, True, True):
        ctx.shard_step(g.FMT_C64, L, 0.01, use_graph=use_graph, stream=s)
        devmem.sync()
        out.append(ctx.burst_finish()[1].copy())
    assert len(out[0]) >= 50
    assert out[0].tobytes() == out[1].tobytes() == out[2].tobytes()
    check_against_oracle(out[0], co.mag2(d["iq"]), sps, 0.01)
    ctx.close()


def test_peer_errors():
    ctx = g.Context(2, 1 << 16)
    with pytest.raises(g.AdsbError):
        ctx.shard_step(g.FMT_C64, 1 << 16, 0.01)          # not attached
    ptr, _ = ctx.shard_alloc(8 << 16)
    ctx.shard_attach_ptrs(0, 1, [ptr])
    with pytest.raises(g.AdsbError):
        ctx.shard_step(g.FMT_C64, 100, 0.01)              # shorter than a burst
    with pytest.raises(g.AdsbError):
        ctx.shard_step(g.FMT_C64, 1 << 17, 0.01)          # more than max_items
    with pytest.raises(g.AdsbError):
        ctx.shard_write(ptr, 16, byte_offset=(8 << 16))   # outside the allocation
    ctx.close()


@pytest.mark.parametrize("fmt", [g.FMT_SC16, g.FMT_CU8])
def test_peer_shards_integer_formats(fmt):
    """int16 / uint8 IQ shards (widened on the device) across the cuts: equal to the whole-buffer oracle on the widened samples."""
    sps, world, L = 2, 4, 12288
    n = L * world
    d = synth.synth_numpy(n, sps=sps, n_bursts=int(n * 2e-3), seed=321)
    ri = d["iq"].view(np.float32).reshape(-1, 2)
    if fmt == g.FMT_SC16:
        raw = np.clip(np.round(ri * 32768.0), -32768, 32767).astype(np.int16)
        v = raw.astype(np.float32) * np.float32(2.0 ** -15)
    else:
        raw = np.clip(np.round(ri * 128.0 + 127.5), 0, 255).astype(np.uint8)
        v = (raw.astype(np.float32) - np.float32(127.5)) * np.float32(1.0 / 128)
    x = (v[:, 0] * v[:, 0] + v[:, 1] * v[:, 1]).astype(np.float32)
    hits, rounds = run_peer(raw, fmt, sps, 0.02, world)
    assert len(hits) > 20
    check_against_oracle(hits, x, sps, 0.02)
