"""Property tests (hypothesis) that pin the two oracle restatements to each other -- and, in the build
container, to the live reference -- on adversarial small inputs: arbitrary float32 samples (zeros, ties,
negatives, inf, nan), arbitrary thresholds and ragged work() chunkings."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from oracle import adsb_oracle as ao
from oracle import c_oracle as co
from oracle import run_reference as rr
import golden_io as G

sample = st.one_of(
    st.sampled_from([0.0, 0.005, 0.01, 0.02, 0.25, 0.5, 1.0, 1e-30, 3e38]),
    st.floats(min_value=0.0, max_value=2.0, width=32, allow_nan=False),
    st.sampled_from([float("nan"), float("inf"), -1.0, -0.0]),
)


@st.composite
def streams(draw):
    sps = draw(st.sampled_from([2, 2, 4, 8]))
    n = draw(st.integers(min_value=8 * sps, max_value=700))
    base = draw(st.lists(sample, min_size=n, max_size=n))
    x = np.array(base, dtype=np.float32)
    # plant a few preamble shapes so that detections actually happen
    h = sps // 2
    for _ in range(draw(st.integers(0, 4))):
        p = draw(st.integers(0, max(0, n - 16 * h - 1)))
        a = np.float32(draw(st.sampled_from([0.02, 0.3, 1.0])))
        x[p:p + 16 * h] = np.minimum(x[p:p + 16 * h], a * 0.4)
        for k in (0, 2, 7, 9):
            x[p + k * h] = a
    thr = draw(st.sampled_from([0.01, 0.0, 0.5, 0.005, 1e-3]))
    chunk = draw(st.sampled_from([None, 1, 7, 64, 100, 333]))
    return x, sps, thr, chunk


@settings(max_examples=150, deadline=None)
@given(streams())
def test_numpy_and_c_ports_agree(case):
    x, sps, thr, chunk = case
    with np.errstate(all="ignore"):
        a = ao.run_stream(x, sps * 1e6, thr, chunk)
    c = co.run_stream(x, sps * 1e6, thr, chunk, cap=len(x) + 8)
    assert [t[0] for t in a["tags"]] == list(c["tag_off"])
    assert G.snr_close([t[1] for t in a["tags"]], c["tag_snr"])
    assert [p[0] for p in a["pdus"]] == list(c["pdu_off"])
    assert all(np.array_equal(p[2], c["pdu_bits"][i]) for i, p in enumerate(a["pdus"]))
    assert a["state"][-1][1] == c["state"][1]


@pytest.mark.skipif(not rr.available(), reason="/root/reference not present (GPU box)")
@settings(max_examples=60, deadline=None)
@given(streams())
def test_numpy_port_agrees_with_live_reference(case):
    x, sps, thr, chunk = case
    a = rr.run_stream(x, sps * 1e6, thr, chunk)
    with np.errstate(all="ignore"):
        b = ao.run_stream(x, sps * 1e6, thr, chunk)
    assert [t[0] for t in a["tags"]] == [t[0] for t in b["tags"]]
    assert G.snr_close([t[1] for t in a["tags"]], [t[1] for t in b["tags"]], tol=0.0)
    assert [p[0] for p in a["pdus"]] == [p[0] for p in b["pdus"]]
    assert all(np.array_equal(p[2], q[2]) for p, q in zip(a["pdus"], b["pdus"]))
    assert [s[1] for s in a["state"]] == [s[1] for s in b["state"]]
