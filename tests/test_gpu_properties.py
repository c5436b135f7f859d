"""Property tests of the CUDA path against the C oracle on adversarial float32 streams (zeros, ties,
negatives, inf, nan, subnormal-range thresholds) -- whole-buffer fused entry point and per-call blocks.
(-0.0 is left out: NumPy's median of a window holding both zeros may return either sign.)"""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings, strategies as st

import golden_io as G
import gr_adsb_b200 as g
from harness import run_blocks
from oracle import adsb_oracle as ao
from oracle import c_oracle as co

pytestmark = pytest.mark.gpu

sample = st.one_of(
    st.sampled_from([0.0, 0.005, 0.01, 0.02, 0.25, 0.5, 1.0, 1e-30, 3e38]),
    st.floats(min_value=0.0, max_value=2.0, width=32, allow_nan=False),
    st.sampled_from([float("nan"), float("inf"), -1.0, -0.5]),
)
_ctx = {}


def ctx_for(sps):
    if sps not in _ctx:
        _ctx[sps] = g.Context(sps, 1 << 16)
    _ctx[sps].reset_state()
    return _ctx[sps]


@st.composite
def streams(draw):
    sps = draw(st.sampled_from([2, 2, 4, 8, 6]))
    n = draw(st.integers(min_value=8 * sps, max_value=6000))
    rng = np.random.default_rng(draw(st.integers(0, 2**31 - 1)))
    kind = draw(st.sampled_from(["list", "noise", "high"]))
    if kind == "list":
        n = min(n, 600)
        x = np.array(draw(st.lists(sample, min_size=n, max_size=n)), dtype=np.float32)
    elif kind == "noise":
        x = rng.exponential(draw(st.sampled_from([1e-3, 6e-3, 0.05])), n).astype(np.float32)
    else:       # mostly above threshold: long runs
        x = (0.02 + rng.exponential(0.01, n)).astype(np.float32)
        x[rng.integers(0, n, max(1, n // 40))] = 0.0
    h = sps // 2
    for _ in range(draw(st.integers(0, 6))):
        p = draw(st.integers(0, max(0, n - 16 * h - 1)))
        a = np.float32(draw(st.sampled_from([0.02, 0.3, 1.0])))
        x[p:p + 16 * h] = np.minimum(x[p:p + 16 * h], a * 0.4)
        for k in (0, 2, 7, 9):
            x[p + k * h] = a
    thr = draw(st.sampled_from([0.01, 0.0, 0.5, 0.005, 1e-3, 1e-38]))
    return x, sps, thr


@settings(max_examples=250, deadline=None, suppress_health_check=[HealthCheck.too_slow])
@given(streams())
def test_fused_whole_buffer_equals_oracle(case):
    x, sps, thr = case
    ref = co.run_stream(x, sps * 1e6, thr, None, cap=len(x) + 8)
    hits = ctx_for(sps).burst_work(x, threshold=thr)
    H1 = 8 * sps - 1
    assert np.array_equal(hits["idx"] - H1, ref["tag_off"])
    assert G.snr_close(hits["snr"], ref["tag_snr"])
    valid = (hits["flags"] & g.HIT_DEMOD_VALID) != 0
    assert np.array_equal(hits["idx"][valid] - H1, ref["pdu_off"])
    assert np.array_equal(g.unpack_bits(hits[valid]), ref["pdu_bits"])
    assert ctx_for.__globals__["_ctx"][sps].get_state()[1] == ref["state"][1]


@settings(max_examples=60, deadline=None, suppress_health_check=[HealthCheck.too_slow])
@given(streams(), st.sampled_from([1, 7, 64, 100, 333, 2048]))
def test_blocks_per_call_equal_oracle(case, chunk):
    x, sps, thr = case
    with np.errstate(all="ignore"):
        ref = ao.run_stream(x, sps * 1e6, thr, chunk)
    fr, dm = g.framer(sps * 1e6, thr), g.demod(sps * 1e6)
    r = run_blocks(fr, dm, x, ao.chunk_sizes(len(x), chunk))
    assert [t[0] for t in r["tags"]] == [t[0] for t in ref["tags"]]
    assert G.snr_close([t[1] for t in r["tags"]], [t[1] for t in ref["tags"]])
    assert [p[0] for p in r["pdus"]] == [p[0] for p in ref["pdus"]]
    assert all(np.array_equal(p[2], q[2]) for p, q in zip(r["pdus"], ref["pdus"]))
    assert [s[1] for s in r["state"]] == [s[1] for s in ref["state"]]
    fr._ctx.close()
    dm._ctx.close()
