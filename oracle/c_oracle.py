"""ctypes binding of oracle/adsb_oracle.c — TEST INFRASTRUCTURE ONLY.

`build()` compiles `liboracle_adsb.so` next to this file with gcc (see oracle/Makefile).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_DIR = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_DIR, "liboracle_adsb.so")
_lib = None


def build(force=False):
    src = os.path.join(_DIR, "adsb_oracle.c")
    if force or not os.path.isfile(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _DIR, "-B", "liboracle_adsb.so"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        p = C.c_void_p
        L.adsb_oracle_run_stream.restype = C.c_int
        L.adsb_oracle_run_stream.argtypes = [p, C.c_int64, C.c_int, C.c_float, C.c_int64,
                                             p, p, p, p, p, p, p, C.c_int64, p, p]
        L.adsb_oracle_mag2.restype = None
        L.adsb_oracle_mag2.argtypes = [p, C.c_int64, p]
        L.adsb_oracle_crc24_syndrome.restype = C.c_uint32
        L.adsb_oracle_crc24_syndrome.argtypes = [p, C.c_int]
        L.adsb_oracle_crc_pass.restype = C.c_int
        L.adsb_oracle_crc_pass.argtypes = [p]
        _lib = L
    return _lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def run_stream(x, fs, threshold, chunk=None, cap=None):
    """Same contract as adsb_oracle.run_stream (uniform chunk size only); arrays instead of lists:
    returns dict(tag_off int64[], tag_snr f32[], pdu_off int64[], pdu_snr f32[], pdu_bits u8[n,112],
    state=(prev_in0, prev_eob_idx) after the last call)."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    sps = int(fs // 1e6)
    n = len(x)
    cap = cap or (n // (63 * sps) + 1024)
    tag_off = np.empty(cap, np.int64); tag_snr = np.empty(cap, np.float32)
    pdu_off = np.empty(cap, np.int64); pdu_snr = np.empty(cap, np.float32)
    pdu_bits = np.empty((cap, 112), np.uint8)
    nt = C.c_int64(0); npd = C.c_int64(0); pin = C.c_float(0); peob = C.c_int64(0)
    rc = lib().adsb_oracle_run_stream(_ptr(x), n, sps, float(np.float32(threshold)),
                                      -1 if chunk is None else int(chunk),
                                      _ptr(tag_off), _ptr(tag_snr), C.addressof(nt),
                                      _ptr(pdu_off), _ptr(pdu_snr), _ptr(pdu_bits), C.addressof(npd),
                                      cap, C.addressof(pin), C.addressof(peob))
    if rc != 0:
        raise RuntimeError("oracle capacity overflow")
    return {"tag_off": tag_off[:nt.value].copy(), "tag_snr": tag_snr[:nt.value].copy(),
            "pdu_off": pdu_off[:npd.value].copy(), "pdu_snr": pdu_snr[:npd.value].copy(),
            "pdu_bits": pdu_bits[:npd.value].copy(), "state": (pin.value, peob.value)}


def mag2(iq):
    iq = np.ascontiguousarray(iq, dtype=np.complex64)
    out = np.empty(len(iq), np.float32)
    lib().adsb_oracle_mag2(_ptr(iq), len(iq), _ptr(out))
    return out


def crc_pass(bits112):
    b = np.ascontiguousarray(bits112, dtype=np.uint8)
    return int(lib().adsb_oracle_crc_pass(_ptr(b)))


def crc_pass_count(bits):
    b = np.ascontiguousarray(bits, dtype=np.uint8).reshape(-1, 112)
    return sum(int(lib().adsb_oracle_crc_pass(_ptr(r))) for r in b)
