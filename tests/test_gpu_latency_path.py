"""The one-kernel latency path (adsb_small.cuh) behind adsb_framer_work / adsb_demod_work / adsb_burst_work for calls of
GNU Radio size, the bits the framer leaves for the demod block, and the batched PDU hand-off: same results as the
throughput pipeline, the oracle and the per-burst PDUs."""
import numpy as np
import pytest

import gr_adsb_b200 as g
from gr_adsb_b200 import blocks, flowgraph, synth
from oracle import adsb_oracle as ao
from oracle import c_oracle as co

pytestmark = pytest.mark.gpu


def chunked_hits(ctx, buf, sizes, thr, sps):
    """Feed `buf` through ctx.burst_work in work()-like calls (history prepended); hits get absolute indices."""
    H1 = 8 * sps - 1
    hist = np.zeros(H1, dtype=buf.dtype)
    pos, out, states = 0, [], []
    for n in sizes:
        in0 = np.concatenate((hist, buf[pos:pos + n]))
        h = ctx.burst_work(in0, n_out=n, lead=0, threshold=thr).copy()
        h["idx"] += pos - H1
        out.append(h)
        states.append(ctx.get_state())
        hist = in0[len(in0) - H1:]
        pos += n
    return np.concatenate(out), states


@pytest.mark.parametrize("sps,fmt,chunk", [(2, "f32", 8192), (2, "c64", 8192), (2, "f32", 4096 + 37), (8, "c64", 65536),
                                           (4, "f32", 1000), (6, "c64", 5000), (2, "c64", 200000)])
def test_latency_path_equals_throughput_pipeline(sps, fmt, chunk):
    n = 1 << 19
    d = synth.synth_numpy(n, sps=sps, n_bursts=int(n * 2e-3 * 2 / sps), seed=500 + sps, amp_db=(-15.0, 0.0))
    buf = d["iq"] if fmt == "c64" else synth.mag2(d["iq"])
    sizes = [chunk] * (n // chunk) + ([n % chunk] if n % chunk else [])
    a = g.Context(sps, 1 << 20)
    b = g.Context(sps, 1 << 20)
    b.set_latency_path(False)
    ha, sa = chunked_hits(a, buf, sizes, 0.01, sps)
    hb, sb = chunked_hits(b, buf, sizes, 0.01, sps)
    calls, fallbacks = a.latency_stats()
    assert calls == len(sizes) and fallbacks == 0
    assert b.latency_stats() == (0, 0)
    assert len(ha) > 20
    assert ha.tobytes() == hb.tobytes()          # every field of every hit, bit for bit
    assert [(np.float32(x).tobytes(), e) for x, e in sa] == [(np.float32(x).tobytes(), e) for x, e in sb]
    # and both equal the oracle fed the same work() sizes
    ref = co.run_stream(synth.mag2(d["iq"]), sps * 1e6, 0.01, chunk)
    assert np.array_equal(ha["idx"], ref["tag_off"])
    valid = (ha["flags"] & g.HIT_DEMOD_VALID) != 0
    assert np.array_equal(g.unpack_bits(ha[valid]), ref["pdu_bits"])
    a.close()
    b.close()


def preamble_train(n, seed):
    rng = np.random.default_rng(seed)
    x = rng.exponential(1e-4, n).astype(np.float32)
    pos = 300
    while pos < n - 300:
        a = np.float32(rng.uniform(0.1, 1.0))
        for hs in (0, 2, 7, 9):
            x[pos + hs] += a
        pos += int(rng.integers(40, 120))
    return x


def test_latency_path_falls_back_when_candidates_overflow():
    """More than 2048 preamble matches in one small call: the in-kernel resolve hands over to the resolve kernels."""
    n = 220000
    x = preamble_train(n + 15, 21)
    ctx = g.Context(2, 1 << 18)
    hits = ctx.burst_work(np.concatenate((np.zeros(15, np.float32), x[:n])), n_out=n, lead=0, threshold=0.01).copy()
    calls, fallbacks = ctx.latency_stats()
    assert (calls, fallbacks) == (1, 1)
    assert ctx.counters()["n_candidates"] > 2048
    ref = co.run_stream(x[:n], 2e6, 0.01, None)
    assert np.array_equal(hits["idx"] - 15, ref["tag_off"])
    # the next small call works from a clean header again
    hits2 = ctx.burst_work(np.concatenate((x[n - 15:n], x[:8192])), n_out=8192, lead=0, threshold=0.01)
    big = g.Context(2, 1 << 18)
    big.set_latency_path(False)
    big.burst_work(np.concatenate((np.zeros(15, np.float32), x[:n])), n_out=n, lead=0, threshold=0.01)
    ref2 = big.burst_work(np.concatenate((x[n - 15:n], x[:8192])), n_out=8192, lead=0, threshold=0.01)
    assert hits2.tobytes() == ref2.tobytes() and ctx.get_state() == big.get_state()
    assert ctx.latency_stats() == (2, 1)
    ctx.close()
    big.close()


def test_framer_work_entry_point_on_both_paths():
    n = 1 << 16
    d = synth.synth_numpy(n + 15, sps=2, n_bursts=200, seed=31)
    x = synth.mag2(d["iq"])
    a, b = g.Context(2, 1 << 17), g.Context(2, 1 << 17)
    b.set_latency_path(False)
    ia, sa = a.framer_work(x, n, 0.01)
    ib, sb = b.framer_work(x, n, 0.01)
    assert len(ia) > 20 and np.array_equal(ia, ib) and sa.tobytes() == sb.tobytes()
    assert a.latency_stats()[0] == 1 and a.get_state() == b.get_state()
    # demod entry point: same tags through the latency path and the copy pipeline, straddlers left untouched
    sob = np.concatenate((ia + 16, [n - 100]))
    ba, va = a.demod_work(x, sob)
    bb, vb = b.demod_work(x, sob)
    assert np.array_equal(va, vb) and np.array_equal(ba, bb) and va[-1] == 0 and va[:-1].all()
    a.close()
    b.close()


def test_demod_uses_the_framers_bits_only_when_the_samples_match():
    n, chunk = 1 << 17, 8192
    d = synth.synth_numpy(n, sps=2, n_bursts=300, seed=41, nonoverlap=True)
    x = synth.mag2(d["iq"])
    fr, dm = g.framer(2e6, 0.01), g.demod(2e6)
    r = flowgraph.run_chain([fr, dm], x, chunk, collect=True)
    ref = ao.run_stream(x, 2e6, 0.01, chunk)
    assert [p[0] for p in r["pdus"]] == [p[0] for p in ref["pdus"]]
    assert all(np.array_equal(p[2], q[2]) for p, q in zip(r["pdus"], ref["pdus"]))
    assert dm.cache_hits > 200 and dm.cache_misses == 0
    # a demod that is handed DIFFERENT samples than the framer saw must slice its own input (device call), not the cache
    fr2, dm2 = g.framer(2e6, 0.01), g.demod(2e6)
    H1 = 15
    hist = np.zeros(H1, np.float32)
    rng = np.random.default_rng(1)
    pos = 0
    dm2.start_timestamp = 0.0
    while pos < n:
        in0 = np.concatenate((hist, x[pos:pos + chunk]))
        out0 = np.empty(chunk, np.float32)
        fr2._nitems_written = fr2._nitems_read = pos
        fr2._tags_out = []
        fr2.work([in0], [out0])
        other = (out0 * rng.uniform(0.5, 1.5, chunk).astype(np.float32)).astype(np.float32)
        dm2._nitems_written = dm2._nitems_read = pos
        dm2._tags_in = fr2._tags_out
        n0 = len(dm2._msgs["demodulated"])
        dm2.work([other], [np.empty(chunk, np.float32)])
        for meta, vec in dm2._msgs["demodulated"][n0:]:
            sob = int(round(meta["timestamp"] * 2e6)) + 16 - pos
            want = (other[sob: sob + 224: 2] > other[sob + 1: sob + 225: 2]).astype(np.uint8)     # demod.py:87-95
            assert np.array_equal(vec, want)
        hist = in0[len(in0) - H1:]
        pos += chunk
    assert dm2.cache_misses > 200 and dm2.cache_hits == 0
    for b in (fr, dm, fr2, dm2):
        b.close()


@pytest.mark.parametrize("complex_input", [False, True])
def test_batched_pdus_equal_per_burst_pdus(complex_input):
    n, chunk = 1 << 17, 16384
    d = synth.synth_numpy(n, sps=2, n_bursts=400, seed=51)
    x = d["iq"] if complex_input else synth.mag2(d["iq"])
    one = flowgraph.run_chain([g.burst(2e6, 0.01, complex_input=complex_input)], x, chunk, collect=True)
    bb = g.burst(2e6, 0.01, complex_input=complex_input, batch=True)
    many = flowgraph.run_chain([bb], x, chunk, collect=True)
    assert len(bb._msgs["demodulated"]) <= n // chunk < one["n_pdus"]          # one message per work() call at most
    assert one["n_pdus"] == many["n_pdus"] > 100 and one["tags"] == many["tags"]
    assert all(a[0] == b[0] and a[1] == b[1] and np.array_equal(a[2], b[2]) for a, b in zip(one["pdus"], many["pdus"]))
    # the message block that sits in front of an unchanged decoder
    ub = g.pdu_unbatch()
    for m in bb._msgs["demodulated"]:
        ub.handle(m)
    flat = ub._msgs["demodulated"]
    assert len(flat) == one["n_pdus"]
    assert all(set(m[0]) == {"timestamp", "snr"} and m[1].dtype == np.uint8 and m[1].shape == (112,) for m in flat)
    # two-block chain with a batching demod
    if not complex_input:
        ch = flowgraph.run_chain([g.framer(2e6, 0.01), g.demod(2e6, batch=True)], x, chunk, collect=True)
        assert [p[0] for p in ch["pdus"]] == [p[0] for p in one["pdus"]]

    class FakeDecoder(object):
        def __init__(self):
            self.got = []

        def decode_packet(self, pdu):
            self.got.append(pdu)

    dec = FakeDecoder()
    for m in bb._msgs["demodulated"]:
        g.feed_decoder(dec, m)
    assert len(dec.got) == one["n_pdus"]
