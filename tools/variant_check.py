#!/usr/bin/env python
"""A/B harness for compile-time variants of the scan kernel: build each variant into its own shared library,
then -- on the B200 -- check every library against the C oracle and time the headline workload with it.

    python tools/variant_check.py build                         # here (nvcc, no GPU): csrc/variants/*.so
    python tools/variant_check.py run [--log2n 28] [--iters 30] # on the GPU box: one JSON line per variant,
                                                                # also appended to gpurun_out/variants.jsonl

No torch: device memory comes from libcudart through ctypes so that a cold box does not pay torch's import.
The oracle is used as the checker only (tools/ is not a product path)."""
import ctypes as C
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
CSRC = os.path.join(ROOT, "gr_adsb_b200", "csrc")
VDIR = os.path.join(CSRC, "variants")

# compile-time switches currently under test (adsb_scan.cuh / adsb_device.cuh); "default" = the source as it is
_OPTS = []     # blocks guarded by `#if ADSB_OPT_<name>` in csrc/


def _defs(*on):
    return ["-DADSB_OPT_%s=%d" % (o, 1 if o in on else 0) for o in _OPTS]


VARIANTS = {
    "default": _defs(),
    # "cw16": _defs() + ["-DADSB_CW=16"],     # 16-sample detect chunks (round-1 geometry: bank-conflicted pre-filter loads)
    # "cw8_16": _defs() + ["-DADSB_CW8=16"],  # 8 Msps / generic kernels with the round-1 chunk width
    # "cw12": _defs() + ["-DADSB_CW=12", "-DADSB_MIN_CTAS=8"],                     # 3072-sample tiles
    # "cw12x2": _defs() + ["-DADSB_CW=12", "-DADSB_NCH2=2", "-DADSB_MIN_CTAS=6"],  # 6144-sample tiles, two chunks per thread
    "paironly": _defs() + ["-DADSB_QUAD_MEDIANS_SPS=-1"],   # phase-3 medians two candidates per warp at every rate
    # latency path with %globaltimer stamps per phase (tools/latency_probe.py prints them)
    "timing": _defs() + ["-DADSB_SMALL_TIMING=1"],
}


def lib_of(name):
    return os.path.join(VDIR, "libadsb_%s.so" % name)


def build():
    from gr_adsb_b200 import _native
    os.makedirs(VDIR, exist_ok=True)
    procs = []
    for name, defs in VARIANTS.items():
        cmd = [os.environ.get("NVCC", "nvcc")] + _native.NVCC_FLAGS + defs + ["-o", lib_of(name), os.path.join(CSRC, "adsb_api.cu")]
        procs.append((name, subprocess.Popen(cmd)))
    for name, p in procs:
        if p.wait() != 0:
            raise SystemExit("variant %s failed to build" % name)
        print("built", lib_of(name))


# ------------------------------------------------------------------------------------------------ one variant
def check_one(name, log2n, iters):
    import numpy as np
    import gr_adsb_b200 as g
    from gr_adsb_b200 import synth
    from oracle import c_oracle as co

    out = {"variant": name, "parity": {}, "ok": True}

    def parity(tag, sps, buf, thr):
        n = len(buf)
        ctx = g.Context(sps, n)
        hits = ctx.burst_work(buf, threshold=thr)
        x = buf if buf.dtype == np.float32 else co.mag2(buf)
        ref = co.run_stream(x, sps * 1e6, thr, None)
        H1 = 8 * sps - 1
        ok = np.array_equal(hits["idx"] - H1, ref["tag_off"])
        if ok:
            a, b = hits["snr"], ref["tag_snr"]
            fin = np.isfinite(b)
            ok = bool(np.array_equal(np.isnan(a), np.isnan(b)) and np.array_equal(np.isinf(a), np.isinf(b)) and
                      np.all(np.abs(a[fin] - b[fin]) <= 1e-4))
        valid = (hits["flags"] & g.HIT_DEMOD_VALID) != 0
        ok = ok and np.array_equal(hits["idx"][valid] - H1, ref["pdu_off"]) and \
            np.array_equal(g.unpack_bits(hits[valid]), ref["pdu_bits"])
        ok = ok and int(((hits["flags"] & g.HIT_CRC_OK) != 0).sum()) == co.crc_pass_count(ref["pdu_bits"])
        st = ctx.get_state()
        ok = ok and st[1] == ref["state"][1] and np.float32(st[0]).view(np.uint32) == np.float32(ref["state"][0]).view(np.uint32)
        ctx.close()
        out["parity"][tag] = {"ok": bool(ok), "hits": int(len(hits))}
        out["ok"] = out["ok"] and bool(ok)

    n_par = 1 << 21
    d2 = synth.synth_numpy(n_par, sps=2, n_bursts=int(n_par * 5e-3), seed=4001)
    parity("sps2_c64_dense", 2, d2["iq"], 0.01)
    parity("sps2_f32_lowthr", 2, co.mag2(d2["iq"]), 0.002)
    d8 = synth.synth_numpy(n_par, sps=8, n_bursts=int(n_par * 1.25e-3), seed=4002)
    parity("sps8_c64", 8, d8["iq"], 0.01)
    d4 = synth.synth_numpy(n_par, sps=4, n_bursts=int(n_par * 2.5e-3), seed=4003)
    parity("sps4_c64", 4, d4["iq"], 0.01)
    d6 = synth.synth_numpy(n_par // 2, sps=6, n_bursts=1000, seed=4004)
    parity("sps6_generic", 6, d6["iq"], 0.01)

    # ---- golden vectors of the unmodified reference (tests/golden), whole-buffer chunking, |IQ|^2 and IQ input ----
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import golden_io as G
    gold_ok, gold_n = True, 0
    for cname in G.case_names():
        c = G.Case(cname)
        whole = [ch for ch in c.chunkings if len(ch.sizes) == 1]
        if not whole:
            continue
        gd = whole[0]
        H1 = 8 * c.sps - 1
        for buf in ([c.x] + ([c.iq] if c.iq is not None else [])):
            ctx = g.Context(c.sps, max(len(c.x), 1 << 16))
            hits = ctx.burst_work(buf, threshold=c.threshold)
            valid = (hits["flags"] & g.HIT_DEMOD_VALID) != 0
            ok = (np.array_equal(hits["idx"] - H1, gd.tag_off) and G.snr_close(hits["snr"], gd.tag_snr) and
                  np.array_equal(hits["idx"][valid] - H1, gd.pdu_off) and
                  np.array_equal(g.unpack_bits(hits[valid]), gd.pdu_bits) and
                  np.array_equal(((hits["flags"][valid] & g.HIT_CRC_OK) != 0).astype(np.uint8), gd.pdu_crc) and
                  ctx.get_state()[1] == gd.state_eob[-1])
            ctx.close()
            gold_n += 1
            if not ok:
                gold_ok = False
                out["parity"].setdefault("golden_failed", []).append(cname)
    out["parity"]["golden"] = {"ok": gold_ok, "cases": gold_n}
    out["ok"] = out["ok"] and gold_ok

    # ---- timing: the bench workload (2 Msps, 5e-3 bursts/sample) and BASELINE configs[2] (8 Msps, 1.25e-3),
    # 2^log2n samples resident in HBM ----
    for key, sps, dens in (("timing", 2, 5e-3), ("timing_sps8", 8, 1.25e-3)):
        out[key] = time_one(g, synth, sps, dens, log2n, iters)
    return out


def time_one(g, synth, sps, dens, log2n, iters):
    import numpy as np
    n = 1 << log2n
    seg = 1 << min(log2n, 24)
    d = synth.synth_numpy(seg, sps=sps, n_bursts=int(seg * dens), seed=1090)
    host = np.ascontiguousarray(d["iq"])
    rt = None
    if os.environ.get("ADSB_VARIANT_HOSTMEM") == "1":
        # dry run of this script against the host emulation (tests/emu): "device" memory is host memory
        big = np.tile(host, n // seg)
        dev = C.c_void_p(big.ctypes.data)
    else:
        for nm in ("libcudart.so.12", "libcudart.so", "/usr/local/cuda/lib64/libcudart.so.12", "/usr/local/cuda/lib64/libcudart.so"):
            try:
                rt = C.CDLL(nm)
                break
            except OSError:
                continue
        if rt is None:
            return "libcudart not found"
        rt.cudaMalloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t]
        rt.cudaMemcpy.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
        rt.cudaFree.argtypes = [C.c_void_p]
        dev = C.c_void_p()
        if rt.cudaMalloc(C.byref(dev), n * 8 + 256) != 0:
            return "cudaMalloc failed"
        for r in range(n // seg):
            assert rt.cudaMemcpy(C.c_void_p(dev.value + r * seg * 8), host.ctypes.data_as(C.c_void_p), seg * 8, 1) == 0
        assert rt.cudaDeviceSynchronize() == 0
    ctx = g.Context(sps, n)
    ctx.profile_enable(True)
    scan, res = [], []
    for it in range(iters + 3):
        ctx.burst_work_dev(dev.value, 1, n, n, 8 * sps - 1, 0.01)
        a, b = ctx.profile_last()
        if it >= 3:
            scan.append(a)
            res.append(b)
    cnt, _ = ctx.burst_finish(fetch_hits=False)
    scan.sort()
    res.sort()
    t = {"log2n": log2n, "iters": iters, "scan_ms_median": scan[len(scan) // 2], "scan_ms_min": scan[0],
         "resolve_ms_median": res[len(res) // 2], "gbs_median": n * 8 / max(scan[len(scan) // 2], 1e-9) / 1e6,
         "hits": cnt["n_hits"], "candidates": cnt["n_candidates"], "crc_ok": cnt["n_crc_ok"]}
    ctx.close()
    if rt is not None:
        rt.cudaFree(dev)
    return t


def run(names, log2n, iters, write_best=True):
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    results = []
    for name in names:
        # "variant@N": the same library with the L2 prefetch distance set to N tiles (ADSB_PREFETCH_TILES)
        # "variant%T": the same library with T threads per scan CTA (ADSB_SCAN_THREADS, needs an "nt" build)
        vname, _, pf = name.partition("@")
        vname, _, nt = vname.partition("%")
        env = dict(os.environ, ADSB_B200_LIB=lib_of(vname))
        if pf:
            env["ADSB_PREFETCH_TILES"] = pf
        if nt:
            env["ADSB_SCAN_THREADS"] = nt
        t0 = time.time()
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "one", name, str(log2n), str(iters)], env=env,
                           stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
        line = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else ""
        try:
            res = json.loads(line)
        except Exception:
            res = {"variant": name, "ok": False, "error": (r.stderr or r.stdout)[-800:]}
        res["wall_s"] = round(time.time() - t0, 1)
        results.append(res)
        print(json.dumps(res), flush=True)
        with open(os.path.join(ROOT, "gpurun_out", "variants.jsonl"), "a") as f:
            f.write(json.dumps(res) + "\n")
    good = [r for r in results if r.get("ok") and isinstance(r.get("timing"), dict)]
    if good and write_best:
        best = min(good, key=lambda r: r["timing"]["scan_ms_median"])
        with open(os.path.join(ROOT, "gpurun_out", "best_variant.txt"), "w") as f:
            f.write(best["variant"] + "\n")
        print("best:", best["variant"], best["timing"]["scan_ms_median"], "ms", file=sys.stderr)


if __name__ == "__main__":
    mode = sys.argv[1] if len(sys.argv) > 1 else "run"
    if mode == "build":
        build()
    elif mode == "one":
        print(json.dumps(check_one(sys.argv[2], int(sys.argv[3]), int(sys.argv[4]))))
    else:
        import argparse
        ap = argparse.ArgumentParser()
        ap.add_argument("mode")
        ap.add_argument("--log2n", type=int, default=28)
        ap.add_argument("--iters", type=int, default=30)
        ap.add_argument("--variants", default=",".join(VARIANTS))
        ap.add_argument("--no-best", action="store_true", help="do not (over)write gpurun_out/best_variant.txt")
        a = ap.parse_args()
        run(a.variants.split(","), a.log2n, a.iters, not a.no_best)
