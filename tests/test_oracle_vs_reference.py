"""Fuzz the oracle restatements against the LIVE unmodified reference (build container only)."""
import numpy as np
import pytest

from oracle import adsb_oracle as ao
from oracle import c_oracle as co
from oracle import run_reference as rr
from gr_adsb_b200 import synth
import golden_io as G

pytestmark = pytest.mark.skipif(not rr.available(), reason="/root/reference not present (GPU box)")


@pytest.mark.parametrize("seed", range(12))
def test_fuzz_against_live_reference(seed):
    rng = np.random.default_rng(1000 + seed)
    sps = int(rng.choice([2, 2, 4, 6, 8]))
    n = int(rng.integers(2000, 40000)) * (sps // 2)
    nb = int(rng.integers(0, 60))
    noise = float(10.0 ** rng.uniform(-4, -2))
    thr = float(10.0 ** rng.uniform(-3.3, -1.3))
    d = synth.synth_numpy(n, sps=sps, n_bursts=nb, seed=seed, noise_var=noise, amp_db=(-25.0, 0.0))
    x = synth.mag2(d["iq"])
    chunk = [None, 8192, 4096, 1000, int(rng.integers(1, 3000))][seed % 5]
    a = rr.run_stream(x, sps * 1e6, thr, chunk)
    b = ao.run_stream(x, sps * 1e6, thr, chunk)
    assert [t[0] for t in a["tags"]] == [t[0] for t in b["tags"]]
    assert G.snr_close([t[1] for t in a["tags"]], [t[1] for t in b["tags"]], tol=0.0)
    assert [p[0] for p in a["pdus"]] == [p[0] for p in b["pdus"]]
    assert all(np.array_equal(p[2], q[2]) for p, q in zip(a["pdus"], b["pdus"]))
    assert [s[1] for s in a["state"]] == [s[1] for s in b["state"]]
    c = co.run_stream(x, sps * 1e6, thr, chunk)
    assert [t[0] for t in a["tags"]] == list(c["tag_off"])
    assert G.snr_close([t[1] for t in a["tags"]], c["tag_snr"])
    assert [p[0] for p in a["pdus"]] == list(c["pdu_off"])
    assert all(np.array_equal(p[2], c["pdu_bits"][i]) for i, p in enumerate(a["pdus"]))
    assert a["state"][-1][1] == c["state"][1]


def test_crc_against_live_reference():
    dec = rr.new_decoder()
    rng = np.random.default_rng(3)
    msgs = synth.df17_messages(rng, 16)
    msgs[8:, rng.integers(0, 112, 8)] ^= 1
    for m in msgs:
        assert rr.crc_pass(m, dec) == ao.crc_pass(m) == co.crc_pass(m)
        assert np.array_equal(rr.compute_crc(m[:88], dec), ao.compute_crc(m[:88]))
        assert np.array_equal(rr.compute_crc(m[:88], dec), synth.crc24_of_payload(m[:88]))
