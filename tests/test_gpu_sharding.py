"""Multi-GPU decomposition checked on ONE GPU: the stream is cut into G shards exactly as bench.py does
across ranks (gr_adsb_b200/sharding.py geometry, halos copied from the neighbours, blanking carry
resolved with carry_plan + adsb_burst_reresolve_dev) and the concatenated hits must equal the
whole-buffer oracle."""
import numpy as np
import pytest

import devmem

import gr_adsb_b200 as g
from gr_adsb_b200 import sharding, synth
from oracle import c_oracle as co

pytestmark = pytest.mark.gpu


def run_sharded(iq, sps, thr, world):
    torch = pytest.importorskip("torch")
    n = len(iq)
    L = n // world
    assert L * world == n
    full = devmem.to_dev(torch.from_numpy(iq.view(np.float32).reshape(-1, 2)))
    geos = [sharding.shard_geometry(r, world, L, sps) for r in range(world)]
    ctxs = [g.Context(sps, geo["n_in"]) for geo in geos]
    bufs = []
    for r, geo in enumerate(geos):
        lo = r * L - geo["left"]
        hi = (r + 1) * L + geo["right"]
        bufs.append(full[lo:hi].contiguous())          # what exchange_halos assembles on rank r
        assert bufs[-1].shape[0] == geo["n_in"]
    s = devmem.stream()
    used = [sharding.NONE] * world
    for r, geo in enumerate(geos):
        ctxs[r].burst_work_dev(bufs[r].data_ptr(), 1, geo["n_in"], geo["n_out"], geo["lead"], thr,
                               base=geo["base"], eob_in=-1, halo_left=geo["halo_left"], stream=s)
    rounds = reruns = 0
    while True:
        devmem.sync()
        infos = []
        for r, geo in enumerate(geos):
            c, _ = ctxs[r].burst_finish(fetch_hits=False)
            first_g = geo["base"] + c["first_candidate"] if c["first_candidate"] >= 0 else sharding.NONE
            last_g = geo["base"] + c["last_accept"] if c["last_accept"] >= 0 else sharding.NONE
            infos.append((first_g, last_g, used[r]))
        e_in, need = sharding.carry_plan(infos, geos[0]["blank"])
        rounds += 1
        if not any(need):
            break
        for r, geo in enumerate(geos):
            if need[r]:
                ctxs[r].burst_reresolve_dev(e_in[r] - geo["base"], stream=s)
                reruns += 1
            used[r] = e_in[r]
        assert rounds <= world + 1
    hits = []
    for r in range(world):
        _, h = ctxs[r].burst_finish(cap=L // (63 * sps) + 8)
        hits.append(h.copy())
        ctxs[r].close()
    return np.concatenate(hits), rounds, reruns


@pytest.mark.parametrize("sps,world,seed", [(2, 2, 1), (2, 4, 2), (2, 8, 3), (8, 2, 4), (8, 4, 5), (4, 3, 6)])
def test_sharded_equals_whole_buffer(sps, world, seed):
    L = 24576 * sps // 2
    n = L * world
    # dense enough that bursts (and blanking intervals) straddle every shard boundary
    d = synth.synth_numpy(n, sps=sps, n_bursts=int(n * 2e-3), seed=100 + seed, amp_db=(-20.0, 0.0))
    ref = co.run_stream(co.mag2(d["iq"]), sps * 1e6, 0.01, None)
    hits, rounds, reruns = run_sharded(d["iq"], sps, 0.01, world)
    H1 = 8 * sps - 1
    assert np.array_equal(hits["idx"] - H1, ref["tag_off"])
    valid = (hits["flags"] & g.HIT_DEMOD_VALID) != 0
    assert np.array_equal(hits["idx"][valid] - H1, ref["pdu_off"])
    assert np.array_equal(g.unpack_bits(hits[valid]), ref["pdu_bits"])
    fin = np.isfinite(ref["tag_snr"])
    assert np.all(np.abs(hits["snr"][fin] - ref["tag_snr"][fin]) <= 1e-4)


def test_carry_crosses_boundaries():
    """Preamble-only pulses every ~100 samples: every shard's first candidates sit inside the blanking
    interval of the previous shard's last hit, so the carry fix-up must run (and cascade)."""
    sps, world = 2, 4
    L = 8192
    n = L * world
    rng = np.random.default_rng(9)
    x = rng.exponential(1e-4, n).astype(np.float32)
    pos = 300
    while pos < n - 300:
        a = np.float32(rng.uniform(0.1, 1.0))
        for hs in (0, 2, 7, 9):
            x[pos + hs] += a
        pos += int(rng.integers(40, 120))
    iq = np.sqrt(x).astype(np.complex64)           # the oracle sees the same |iq|^2
    ref = co.run_stream(co.mag2(iq), 2e6, 0.01, None)
    hits, rounds, reruns = run_sharded(iq, sps, 0.01, world)
    assert np.array_equal(hits["idx"] - 15, ref["tag_off"])
    assert reruns >= 1


def run_sharded_device_carry(iq, sps, thr, world):
    """Same decomposition, but the carry is planned on the device (adsb_burst_carry_round_dev) the way
    bench.py does it under NCCL: export -> (all-gather) -> plan + conditional re-resolve, then a final check."""
    torch = pytest.importorskip("torch")
    n = len(iq)
    L = n // world
    full = devmem.to_dev(torch.from_numpy(iq.view(np.float32).reshape(-1, 2)))
    geos = [sharding.shard_geometry(r, world, L, sps) for r in range(world)]
    ctxs = [g.Context(sps, geo["n_in"]) for geo in geos]
    bufs = [full[r * L - geo["left"]: (r + 1) * L + geo["right"]].contiguous() for r, geo in enumerate(geos)]
    s = devmem.stream()
    gathered = torch.zeros(3 * world, dtype=torch.int64, device=devmem.DEVICE)
    for r, geo in enumerate(geos):
        ctxs[r].burst_work_dev(bufs[r].data_ptr(), 1, geo["n_in"], geo["n_out"], geo["lead"], thr,
                               base=geo["base"], eob_in=-1, halo_left=geo["halo_left"], stream=s)
    seen = [0] * world
    rounds = 0
    while True:
        for final in (False, True):
            for r in range(world):      # every rank's export lands in its row: what all_gather_into_tensor does
                ctxs[r].burst_export_carry_dev(g.CARRY_USED_ON_DEVICE, gathered.data_ptr() + 24 * r, stream=s)
            for r in range(world):
                ctxs[r].burst_carry_round_dev(gathered.data_ptr(), world, r, final_check=final, stream=s)
        devmem.sync()
        rounds += 1
        now = [ctxs[r].burst_finish(fetch_hits=False)[0]["carry_unstable"] for r in range(world)]
        settled = now == seen
        seen = now
        if settled:
            break
        assert rounds <= world + 1
    hits = []
    for r in range(world):
        _, h = ctxs[r].burst_finish(cap=L // (63 * sps) + 8)
        hits.append(h.copy())
        ctxs[r].close()
    return np.concatenate(hits), rounds


@pytest.mark.parametrize("sps,world,seed", [(2, 2, 11), (2, 8, 12), (8, 4, 13)])
def test_device_carry_equals_whole_buffer(sps, world, seed):
    L = 24576 * sps // 2
    n = L * world
    d = synth.synth_numpy(n, sps=sps, n_bursts=int(n * 2e-3), seed=200 + seed, amp_db=(-20.0, 0.0))
    ref = co.run_stream(co.mag2(d["iq"]), sps * 1e6, 0.01, None)
    hits, rounds = run_sharded_device_carry(d["iq"], sps, 0.01, world)
    H1 = 8 * sps - 1
    assert np.array_equal(hits["idx"] - H1, ref["tag_off"])
    valid = (hits["flags"] & g.HIT_DEMOD_VALID) != 0
    assert np.array_equal(g.unpack_bits(hits[valid]), ref["pdu_bits"])


def test_device_carry_cascade():
    sps, world = 2, 4
    L = 8192
    n = L * world
    rng = np.random.default_rng(9)
    x = rng.exponential(1e-4, n).astype(np.float32)
    pos = 300
    while pos < n - 300:
        a = np.float32(rng.uniform(0.1, 1.0))
        for hs in (0, 2, 7, 9):
            x[pos + hs] += a
        pos += int(rng.integers(40, 120))
    iq = np.sqrt(x).astype(np.complex64)
    ref = co.run_stream(co.mag2(iq), 2e6, 0.01, None)
    hits, rounds = run_sharded_device_carry(iq, sps, 0.01, world)
    assert np.array_equal(hits["idx"] - 15, ref["tag_off"])
    assert rounds >= 1
