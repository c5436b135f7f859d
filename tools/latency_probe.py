#!/usr/bin/env python
"""Microseconds per C-ABI call at GNU Radio buffer sizes, without the Python block layer: adsb_burst_work /
adsb_framer_work / adsb_demod_work on a fixed pageable host buffer, latency path vs throughput pipeline.
    python tools/latency_probe.py            # on the GPU box
"""
import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gr_adsb_b200 as g
from gr_adsb_b200 import _native, synth

L = _native.lib()
d = synth.synth_numpy(1 << 18, sps=2, n_bursts=13, seed=3)          # BASELINE configs[0] density
x = synth.mag2(d["iq"])
iq = d["iq"]
for n in (8192, 65536):
    for fmt, buf in ((g.FMT_F32, x), (g.FMT_C64, iq)):
        for small in (True, False):
            ctx = g.Context(2, 1 << 18)
            ctx.set_latency_path(small)
            in0 = np.ascontiguousarray(buf[4096: 4096 + n + 15])
            hits = np.zeros(n // 126 + 2, g.HIT_DTYPE)
            nh = C.c_size_t(0)
            args = (ctx._h, in0.ctypes.data_as(C.c_void_p), fmt, n + 15, 0, n, C.c_float(0.01), 0,
                    hits.ctypes.data_as(C.c_void_p), len(hits), C.byref(nh))
            for _ in range(20):
                assert L.adsb_burst_work(*args) == 0
            reps = 300
            t0 = time.perf_counter()
            for _ in range(reps):
                L.adsb_burst_work(*args)
            dt = (time.perf_counter() - t0) / reps
            print("adsb_burst_work  n=%6d %-9s %-10s %7.1f us/call  (%d hits)" % (
                n, "float32" if fmt == g.FMT_F32 else "complex64", "latency" if small else "pipeline", dt * 1e6, nh.value))
            if small:
                dev = (C.c_uint64 * 8)()
                host = (C.c_double * 3)()
                acc = np.zeros(3)
                accd = np.zeros(5)
                for _ in range(50):
                    L.adsb_burst_work(*args)
                    L.adsb_latency_trace(ctx._h, dev, host)
                    acc += np.array(host[:])
                    if dev[0]:
                        accd += np.array([dev[2] - dev[0], dev[3] - dev[2], dev[4] - dev[3], dev[5] - dev[4], dev[5] - dev[0]]) * 1e-3
                print("    host us: stage %.1f, launch %.1f, wait %.1f" % tuple(acc / 50 * 1e6), end="")
                print("; device us (last CTA): scan %.1f, slice %.1f, ticket %.1f, resolve %.1f, total %.1f" % tuple(accd / 50)
                      if dev[0] else "")
            if fmt == g.FMT_F32:
                idx = np.zeros(len(hits), np.int64)
                snr = np.zeros(len(hits), np.float32)
                fa = (ctx._h, in0.ctypes.data_as(C.c_void_p), n + 15, n, C.c_float(0.01), idx.ctypes.data_as(C.c_void_p),
                      snr.ctypes.data_as(C.c_void_p), len(hits), C.byref(nh))
                for _ in range(20):
                    assert L.adsb_framer_work(*fa) == 0
                t0 = time.perf_counter()
                for _ in range(reps):
                    L.adsb_framer_work(*fa)
                dt = (time.perf_counter() - t0) / reps
                print("adsb_framer_work n=%6d %-9s %-10s %7.1f us/call" % (n, "float32", "latency" if small else "pipeline", dt * 1e6))
                sob = (idx[:max(1, nh.value)] + 16).copy()
                bits = np.zeros((len(sob), 112), np.uint8)
                valid = np.zeros(len(sob), np.uint8)
                da = (ctx._h, in0.ctypes.data_as(C.c_void_p), n + 15, sob.ctypes.data_as(C.c_void_p), len(sob),
                      bits.ctypes.data_as(C.c_void_p), valid.ctypes.data_as(C.c_void_p))
                for _ in range(20):
                    assert L.adsb_demod_work(*da) == 0
                t0 = time.perf_counter()
                for _ in range(reps):
                    L.adsb_demod_work(*da)
                dt = (time.perf_counter() - t0) / reps
                print("adsb_demod_work  n=%6d %d tags    %-10s %7.1f us/call" % (n, len(sob), "latency" if small else "pipeline", dt * 1e6))
            ctx.close()
