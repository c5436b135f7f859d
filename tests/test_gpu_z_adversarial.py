"""Adversarial streams long enough to reach the INTERIOR-tile fast paths of the scan kernel (the hypothesis streams of
test_gpu_properties.py are shorter than two tiles and only see the generic path): plateaus of every length, preamble
shapes at and across tile boundaries, exact ties with half the centre amplitude, values at the threshold +- 1 ulp,
denormals, NaN / inf, negative values (float32 input), candidate trains that overflow a tile's fixed record block,
and hundreds of long runs per tile (the deferral list of the warp-cooperative pass nearly full).  Whole-buffer fused entry point against the C oracle, for
float32 |IQ|^2 input and for complex64 IQ input (where the dense pre-filter compares through the sign of a
subtraction and leans on the device's canonical NaN)."""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings, strategies as st

import golden_io as G
import gr_adsb_b200 as g
from oracle import c_oracle as co

pytestmark = pytest.mark.gpu

_ctx = {}


def ctx_for(sps, n):
    c = _ctx.get(sps)
    if c is None or c.max_items < n:
        if c is not None:
            c.close()
        c = g.Context(sps, max(n, 1 << 17))
        _ctx[sps] = c
    c.reset_state()
    return c


def adversarial_mag(n, sps, thr, rng, specials):
    """float32 |IQ|^2 design.  `specials` = values sprinkled over the stream."""
    h = sps // 2
    T = 4096 if sps in (2, 4) else 8192
    x = rng.exponential(rng.choice([1e-3, 4e-3]), n).astype(np.float32)
    # plateaus of every length up to a few tiles' look-back
    for _ in range(n // 300):
        p = int(rng.integers(0, n - 400))
        L = int(rng.integers(1, 70)) if rng.random() < 0.9 else int(rng.integers(70, 390))
        x[p:p + L] = np.maximum(x[p:p + L], np.float32(0.011) * np.float32(rng.choice([1, 1, 2, 50])))
    # preamble shapes: random places, and at / across every tile boundary
    pos = [int(v) for v in rng.integers(0, n - 16 * h - 1, n // 400)]
    for k in range(1, n // T + 1):
        for d in (-16 * h - 1, -9 * h, -3, -1, 0, 1, 5, 100, 127, 128, 129):
            if 0 <= k * T + d < n - 16 * h - 1:
                pos.append(k * T + d)
    for p in pos:
        a = np.float32(rng.choice([0.02, 0.3, 1.0]))
        seg = x[p:p + 16 * h]
        np.minimum(seg, a * np.float32(0.4), out=seg)
        for k in (0, 2, 7, 9):
            x[p + k * h] = a
        r = rng.random()
        if r < 0.15:
            x[p + int(rng.choice([1, 3, 4, 5, 6, 8, 10, 15])) * h] = a * np.float32(0.5)     # tie on a low tap: still low
        elif r < 0.25:
            x[p + int(rng.choice([2, 7, 9])) * h] = a * np.float32(0.5)                       # tie on a high tap: no match
    # a train of back-to-back preambles: far more candidates than a tile's fixed record block
    p0 = int(rng.integers(T, max(T + 1, n - 80 * 16 * h)))
    for q in range(60):
        p = p0 + q * 16 * h
        if p + 16 * h >= n:
            break
        x[p:p + 16 * h] = 0.001
        for k in (0, 2, 7, 9):
            x[p + k * h] = 0.5
    # values at the threshold and one ulp around it, and the special values
    t32 = np.float32(thr)
    idx = rng.integers(0, n, n // 150)
    x[idx] = rng.choice(np.array([t32, np.nextafter(t32, np.float32(0)), np.nextafter(t32, np.float32(1))], np.float32), len(idx))
    idx = rng.integers(0, n, n // 400)
    x[idx] = rng.choice(np.array(specials, np.float32), len(idx))
    return x


def check(hits, sps, ref, ctx):
    H1 = 8 * sps - 1
    assert np.array_equal(hits["idx"] - H1, ref["tag_off"])
    assert G.snr_close(hits["snr"], ref["tag_snr"])
    valid = (hits["flags"] & g.HIT_DEMOD_VALID) != 0
    assert np.array_equal(hits["idx"][valid] - H1, ref["pdu_off"])
    assert np.array_equal(g.unpack_bits(hits[valid]), ref["pdu_bits"])
    crc = np.array([co.crc_pass(b) for b in ref["pdu_bits"]], np.uint8)
    assert np.array_equal(((hits["flags"][valid] & g.HIT_CRC_OK) != 0).astype(np.uint8), crc)
    st = ctx.get_state()
    assert st[1] == ref["state"][1]
    a, b = np.float32(st[0]), np.float32(ref["state"][0])
    assert a.view(np.uint32) == b.view(np.uint32) or (np.isnan(a) and np.isnan(b))


@pytest.mark.parametrize("sps,seed,thr", [(2, 1, 0.01), (2, 2, 0.01), (2, 3, 0.004), (4, 4, 0.01), (8, 5, 0.01),
                                           (6, 6, 0.01), (2, 7, 0.0), (16, 8, 0.01)])
def test_fast_path_adversarial_float32(sps, seed, thr):
    rng = np.random.default_rng(1000 + seed)
    n = (9 if sps in (2, 4) else 6) * (4096 if sps in (2, 4) else 8192) + int(rng.integers(0, 3000))
    x = adversarial_mag(n, sps, thr, rng, [np.nan, np.inf, 0.0, 1e-42, 3e-39, -1.0, -0.25, 3e38])
    with np.errstate(all="ignore"):
        ref = co.run_stream(x, sps * 1e6, thr, None, cap=n + 8)
    assert len(ref["tag_off"]) > 30 or thr == 0.0           # thr 0: one run with holes at the negatives / NaNs
    ctx = ctx_for(sps, n)
    hits = ctx.burst_work(x, threshold=thr)
    check(hits, sps, ref, ctx)


@pytest.mark.parametrize("sps,seed,thr", [(2, 11, 0.01), (2, 12, 0.01), (2, 13, 0.004), (4, 14, 0.01), (8, 15, 0.01),
                                           (6, 16, 0.01)])
def test_fast_path_adversarial_complex64(sps, seed, thr):
    rng = np.random.default_rng(2000 + seed)
    n = (9 if sps in (2, 4) else 6) * (4096 if sps in (2, 4) else 8192) + int(rng.integers(0, 3000))
    design = adversarial_mag(n, sps, thr, rng, [0.0, 1e-42, 3e-39, 3e38, 0.25])
    with np.errstate(all="ignore"):
        amp = np.sqrt(design).astype(np.float32)
    ph = rng.uniform(0, 2 * np.pi, n).astype(np.float32)
    ph[rng.random(n) < 0.5] = 0.0                              # half the samples purely real: |IQ|^2 = fl(re*re) exactly
    iq = (amp * np.cos(ph) + 1j * amp * np.sin(ph)).astype(np.complex64)
    # exact ties in |IQ|^2: (1 + 0j) against (0.5 + 0.5j) -> 1.0 and 0.5
    for p in rng.integers(0, n - 16 * sps, 40):
        h = sps // 2
        iq[p:p + 16 * h] = 0.05
        for k in (0, 2, 7, 9):
            iq[p + k * h] = 1.0
        iq[p + int(rng.choice([1, 2, 4, 7, 12])) * h] = 0.5 + 0.5j
    # non-finite samples: NaN / inf in either part, overflowing squares
    idx = rng.integers(0, n, n // 300)
    iq[idx] = rng.choice(np.array([np.nan, np.inf, -np.inf, 1j * np.nan, 1j * np.inf, np.nan + 1j * np.inf,
                                   3e19 + 3e19j, 1e-23 + 0j], np.complex64), len(idx))
    with np.errstate(all="ignore"):
        x = co.mag2(iq)
        ref = co.run_stream(x, sps * 1e6, thr, None, cap=n + 8)
    assert len(ref["tag_off"]) > 30
    ctx = ctx_for(sps, n)
    hits = ctx.burst_work(iq, threshold=thr)
    check(hits, sps, ref, ctx)


@pytest.mark.parametrize("sps", [6, 16, 2])
def test_hundreds_of_long_runs_per_tile(sps):
    """Runs of 33..40 samples separated by single low samples: on the generic path (unspecialised rates, edge tiles)
    every one of them leaves the 32-bit window and is deferred to the warp-cooperative pass -- about 220 per
    8192-sample tile, close to the deferral list's capacity (kSlowCap = 256, which a tile cannot exceed by much:
    a deferred run is at least 16 samples long)."""
    rng = np.random.default_rng(77 + sps)
    n = 3 * 8192 + 1234
    x = np.full(n, 0.02, np.float32)
    p = 5
    while p < n:
        x[p] = 0.0
        p += int(rng.integers(34, 42))
    h = sps // 2
    for p in rng.integers(0, n - 16 * h - 1, 60):              # some preamble-shaped centres inside the long runs
        a = np.float32(rng.choice([0.5, 1.0]))
        x[p:p + 16 * h] = 0.011
        for k in (0, 2, 7, 9):
            x[p + k * h] = a
    ref = co.run_stream(x, sps * 1e6, 0.01, None, cap=n + 8)
    ctx = ctx_for(sps, n)
    hits = ctx.burst_work(x, threshold=0.01)
    check(hits, sps, ref, ctx)


@settings(max_examples=60, deadline=None, suppress_health_check=[HealthCheck.too_slow])
@given(st.integers(0, 2**31 - 1), st.sampled_from([2, 2, 2, 4, 8, 6]), st.sampled_from([0.01, 0.01, 0.004, 0.02, 1e-3]),
       st.booleans())
def test_fast_path_adversarial_random(seed, sps, thr, complex_input):
    """The two generators above under hypothesis-drawn seeds, rates and thresholds."""
    rng = np.random.default_rng(seed)
    T = 4096 if sps in (2, 4) else 8192
    n = int(rng.integers(3, 7)) * T + int(rng.integers(0, T))
    if complex_input:
        design = adversarial_mag(n, sps, thr, rng, [0.0, 1e-42, 3e38, 0.25])
        with np.errstate(all="ignore"):
            amp = np.sqrt(design).astype(np.float32)
        ph = rng.uniform(0, 2 * np.pi, n).astype(np.float32)
        ph[rng.random(n) < 0.5] = 0.0
        buf = (amp * np.cos(ph) + 1j * amp * np.sin(ph)).astype(np.complex64)
        idx = rng.integers(0, n, n // 400)
        buf[idx] = rng.choice(np.array([np.nan, np.inf, 1j * np.inf, 3e19 + 3e19j], np.complex64), len(idx))
        with np.errstate(all="ignore"):
            x = co.mag2(buf)
    else:
        buf = x = adversarial_mag(n, sps, thr, rng, [np.nan, np.inf, 0.0, 1e-42, -1.0, 3e38])
    with np.errstate(all="ignore"):
        ref = co.run_stream(x, sps * 1e6, thr, None, cap=n + 8)
    ctx = ctx_for(sps, n)
    hits = ctx.burst_work(buf, threshold=thr)
    check(hits, sps, ref, ctx)
