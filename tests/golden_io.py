"""Loader for the golden vectors (outputs of the unmodified reference, see golden/make_golden.py)."""
import glob
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SNR_TOL_DB = 1e-4     # BASELINE.json north_star: "reported SNR matches within 1e-4 dB"


def case_names():
    names = (os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))
    return sorted(n for n in names if n != "fec_vectors")


class Chunking(object):
    def __init__(self, z, ci):
        g = lambda k: z["c%d_%s" % (ci, k)]
        self.sizes = [int(v) for v in g("sizes")]
        self.tag_off = g("tag_off")
        self.tag_snr = g("tag_snr")
        self.pdu_off = g("pdu_off")
        self.pdu_snr = g("pdu_snr")
        self.pdu_bits = np.unpackbits(g("pdu_bits"), axis=1)[:, :112] if len(g("pdu_off")) else np.zeros((0, 112), np.uint8)
        self.pdu_crc = g("pdu_crc")
        self.state_in0 = g("state_in0")
        self.state_eob = g("state_eob")
        key = "c%d_thresholds" % ci
        self.thresholds = [float(v) for v in z[key]] if key in z.files else None


class Case(object):
    def __init__(self, name):
        z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
        self.name = name
        self.x = z["x"]
        self.iq = z["iq"] if "iq" in z.files else None
        self.fs = float(z["fs"])
        self.sps = int(self.fs // 1e6)
        self.threshold = float(z["threshold"])
        self.chunkings = [Chunking(z, ci) for ci in range(int(z["n_chunkings"]))]


def all_chunkings():
    out = []
    for n in case_names():
        c = Case(n)
        for ci in range(len(c.chunkings)):
            out.append((n, ci))
    return out


def snr_close(a, b, tol=SNR_TOL_DB):
    """|a-b| <= tol where finite; nan must match nan, +-inf must match exactly."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    if a.shape != b.shape:
        return False
    fin = np.isfinite(a)
    if not np.array_equal(fin, np.isfinite(b)):
        return False
    if not np.array_equal(np.isnan(a), np.isnan(b)):
        return False
    inf = np.isinf(a)
    if inf.any() and not np.array_equal(a[inf], b[inf]):
        return False
    return bool(np.all(np.abs(a[fin] - b[fin]) <= tol))


def fec_vectors():
    """(bits_in uint8[n,112], passed uint8[n], bits_out uint8[n,112]) from the reference decoder run with
    error_corr="Conservative" (golden/make_golden.py:emit_fec)."""
    z = np.load(os.path.join(GOLDEN_DIR, "fec_vectors.npz"))
    unpack = lambda a: np.unpackbits(a, axis=1)[:, :112]
    return unpack(z["bits_in"]), z["passed"], unpack(z["bits_out"])
