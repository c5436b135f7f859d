"""Float32 identities the CUDA kernels lean on, checked with NumPy's IEEE single arithmetic (no GPU needed)."""
import numpy as np


def _check_sign_identity(x, m3):
    """adsb_scan.cuh gt_sign(): for x, m3 in {+0, positive, +inf} the sign bit of fl32(x*0.5 - m3) equals
    (m3 > x/2) -- the dense pre-filter's compare (framer.py:140-141) moved from the ALU to the FMA pipe.
    (inf - inf gives a NaN: the GPU's canonical NaN 0x7fffffff has no sign bit, i.e. "not greater", which is
    what inf > inf says; NumPy's NaN sign differs, so that case is asserted by value here.)"""
    half = (x * np.float32(0.5)).astype(np.float32)
    with np.errstate(invalid="ignore"):
        d = (half - m3).astype(np.float32)
    sign = (d.view(np.uint32) >> 31).astype(bool)
    nan = np.isnan(d)
    assert not np.any((m3 > half)[nan])
    assert np.array_equal(sign[~nan], (m3 > half)[~nan])


def test_prefilter_compare_equals_subtraction_sign():
    rng = np.random.default_rng(7)
    edge = np.array([0.0, 1e-45, 2e-45, 3e-45, 5e-45, 1.1754942e-38, 1.1754944e-38, 2.3509887e-38, 1e-30, 0.5, 1,
                     1.0000001, 2, 3.4028235e38, 1.7014118e38, np.inf], dtype=np.float32)
    a, b = np.meshgrid(edge, edge)
    _check_sign_identity(a.ravel(), b.ravel())
    bits = rng.integers(0, 0x7f800001, size=1_000_000, dtype=np.int64).astype(np.uint32)
    v = bits.view(np.float32)
    _check_sign_identity(v[:500_000], v[500_000:])
    # x within a few ulps of 2*m3, where rounding of the difference would matter if it could
    m3 = v[:500_000]
    with np.errstate(over="ignore"):
        x = (m3 * np.float32(2)).astype(np.float32)
    xb = (x.view(np.uint32).astype(np.int64) + rng.integers(-3, 4, size=x.size)).clip(0, 0x7f800000).astype(np.uint32)
    _check_sign_identity(xb.view(np.float32), m3)
    # every pair of the 2048 smallest non-negative floats (all denormal: halving rounds, subtraction is exact)
    s = np.arange(0, 2048, dtype=np.uint32).view(np.float32)
    a, b = np.meshgrid(s, s)
    _check_sign_identity(a.ravel(), b.ravel())


def test_median_padding_ranks():
    """slice_pair_snr pads the 100 noise samples with 14 keys below and 14 above everything: the median's
    order statistics (ranks 49, 50 of 100; framer.py:157 np.median) become ranks 63, 64 of the 128, and those are
    max_i min(A[i], B[63-i]) / min_i max(A[i], B[63-i]) for the two sorted 64-key halves A, B (half_mid128)."""
    rng = np.random.default_rng(3)
    for _ in range(200):
        x = rng.exponential(1e-3, 100).astype(np.float32)
        if rng.random() < 0.3:
            x[rng.integers(0, 100, 30)] = 0.0                 # ties with the low pads
        keys = np.concatenate([x.view(np.uint32), np.zeros(14, np.uint32), np.full(14, 0x7f800000, np.uint32)])
        rng.shuffle(keys)
        A, B = np.sort(keys[:64]), np.sort(keys[64:])
        lo = np.minimum(A, B[::-1]).max()
        hi = np.maximum(A, B[::-1]).min()
        s = np.sort(x)
        assert lo == s[49].view(np.uint32) and hi == s[50].view(np.uint32)
        med = np.float32(np.float32(lo.view(np.float32) + hi.view(np.float32)) * np.float32(0.5))
        assert med.view(np.uint32) == np.median(x).astype(np.float32).view(np.uint32)
