#!/usr/bin/env python
"""Benchmark of the ADS-B burst path (BASELINE.json): Msamples/s and CRC-passing bursts/s on synthetic
2 Msps complex64 IQ with injected DF17 squitters, whole-buffer semantics.

    python bench.py --gpus 1 --steps 200 --warmup 5            # this framework, one JSON line
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...   # one long stream, N shards
    python bench.py --impl reference ...                        # the reference algorithm on the host cores

A step is one pass of the fused |IQ|^2 + framer + demod + CRC path over the rank's resident
2^28-sample shard (BASELINE config 2: 10k bursts per stream-second = 5e-3 bursts/sample); with N > 1 the
step also exchanges the shard halos and the blanking carry over NCCL (gr_adsb_b200/sharding.py).
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SPS = 2
THRESHOLD = 0.01                 # grc/adsb_framer.block.yml:19 default
DENSITY = 5e-3                   # bursts per sample = 10k bursts per second at 2 Msps
SEED = 1090
METRIC = "Msamples/s through fused |IQ|^2+framer+demod(+CRC), 2 Msps synthetic IQ with DF17 squitters"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (burst copy)"
    except Exception:
        return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"


def ncu_traffic():
    """dram bytes per launch of the scan kernel from the committed ncu capture, if any."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "scan_kernel_traffic.json")))
    except Exception:
        return None


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler(threading.Thread):
    def __init__(self, index):
        threading.Thread.__init__(self, daemon=True)
        self.index = index
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self._halt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:      # pragma: no cover
            log("clock sampling unavailable:", e)

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
            getattr(nv, "nvmlDeviceGetCurrentClocksThrottleReasons")
        while not self._halt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = get_reasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.004)

    def stop(self):
        self._halt.set()
        self.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


# ------------------------------------------------------------------------------------------------ CPU legs
_REF_SAMPLE = None      # bounded sample of the reference arm; inherited by the forked workers (no pickling)


def _ref_worker(_):
    return _cpu_worker((_REF_SAMPLE, None))


def _cpu_worker(args):
    iq, chunk = args
    from oracle import adsb_oracle as ao
    t0 = time.perf_counter()
    x = ao.mag2(iq)
    r = ao.run_stream(x, SPS * 1e6, THRESHOLD, chunk)
    crc = sum(ao.crc_pass(p[2]) for p in r["pdus"])
    return time.perf_counter() - t0, len(r["tags"]), len(r["pdus"]), crc


def cpu_baseline_once(iq, chunk=None):
    """NumPy port of framer.py + demod.py (+ decoder.py's parity check) on one core."""
    dt, tags, pdus, crc = _cpu_worker((iq, chunk))
    return {"value": len(iq) / dt / 1e6, "unit": "Msamples/s", "cores": 1, "kind": "port",
            "sample": "first %d samples of the rank-0 shard (complex64 -> |IQ|^2 -> framer+demod+CRC, %s), %.1f s"
                      % (len(iq), "one work() call = whole-buffer semantics, as on the GPU" if chunk is None else "work() chunks of %d" % chunk, dt),
            "tags": tags, "pdus": pdus, "crc_pass": crc, "crc_pass_bursts_per_s": crc / dt}


def run_reference_arm(args):
    """`--impl reference`: the reference's CPU algorithm (oracle port; the reference itself is Python and
    does not travel to the GPU box) on all host cores, one independent shard per core per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    from gr_adsb_b200 import synth
    cores = os.cpu_count() or 1
    try:
        cores = len(os.sched_getaffinity(0))
    except Exception:
        pass
    # calibrate with all cores busy (they slow each other down a lot), then size the per-step sample so that
    # the K + W steps take ~100 s in total
    global _REF_SAMPLE
    n_cal = 1 << 16
    _REF_SAMPLE = synth.synth_numpy(n_cal, sps=SPS, n_bursts=int(n_cal * DENSITY), seed=SEED)["iq"]
    with mp.get_context("fork").Pool(cores) as pool:
        pool.map(_ref_worker, range(cores), chunksize=1)
        t0 = time.perf_counter()
        pool.map(_ref_worker, range(cores), chunksize=1)
        per_core_rate = n_cal / (time.perf_counter() - t0)
    per_worker = int(max(1 << 13, min(1 << 21, per_core_rate * 100.0 / (args.steps + args.warmup))))
    _REF_SAMPLE = synth.synth_numpy(per_worker, sps=SPS, n_bursts=int(per_worker * DENSITY), seed=SEED + 1)["iq"]
    log("reference arm: %d cores, %d samples per core per step" % (cores, per_worker))
    with mp.get_context("fork").Pool(cores) as pool:       # every core chews the same bounded sample
        for _ in range(args.warmup):
            pool.map(_ref_worker, range(cores), chunksize=1)
        t0 = time.perf_counter()
        crc = 0
        for _ in range(args.steps):
            out = pool.map(_ref_worker, range(cores), chunksize=1)
            crc += sum(o[3] for o in out)
        dt = time.perf_counter() - t0
    total = per_worker * cores * args.steps
    val = total / dt / 1e6
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "Msamples/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(per_worker * cores, 1),
            "crc_pass_bursts_per_s": crc / dt,
            "cpu_baseline": {"value": val, "unit": "Msamples/s", "cores": cores, "kind": "port",
                             "sample": "%d samples per core per step, %d cores, one work() call per sample = whole-buffer semantics (oracle/adsb_oracle.py: "
                                       "NumPy restatement of framer.py+demod.py+CRC; the Python reference cannot travel to the GPU box)"
                                       % (per_worker, cores)},
            "e2e": {"value": val, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def workload_config(shard_items, world):
    return {"workload": "BASELINE configs[1]: 2 Msps complex64 IQ, 10k injected DF17 bursts per stream-second "
                        "(5e-3 bursts/sample, amplitudes -10..0 dB over Exp(1e-3) noise), threshold 0.01, whole-buffer semantics",
            "samples_per_gpu": shard_items, "sps": SPS, "threshold": THRESHOLD, "bursts_per_sample": DENSITY,
            "sharding": ("1 shard" if world == 1 else "%d contiguous shards, NCCL halo + blanking-carry exchange" % world),
            "l2_policy": "inputs (%.2f GiB per GPU) larger than the 126 MB L2; no flush needed" % (shard_items * 8 / 2**30)}


# ------------------------------------------------------------------------------------------------ GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--log2n", type=int, default=int(os.environ.get("ADSB_BENCH_LOG2N", "28")))
    ap.add_argument("--e2e-steps", type=int, default=4)
    ap.add_argument("--no-cpu", action="store_true", help="skip the bounded CPU baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--sps", type=int, default=SPS, help="samples per symbol (secondary workloads; default 2)")
    ap.add_argument("--density", type=float, default=DENSITY, help="bursts per sample (default 5e-3 = 10k/s at 2 Msps)")
    ap.add_argument("--sc16", action="store_true", help="feed interleaved int16 IQ (4 B/sample; SURVEY.md 8f-4) instead of complex64")
    ap.add_argument("--cu8", action="store_true", help="feed interleaved uint8 IQ (2 B/sample, RTL-SDR wire format) instead of complex64")
    ap.add_argument("--fec", action="store_true", help="enable the conservative error correction (decoder.py error_corr)")
    ap.add_argument("--f32", action="store_true", help="feed float32 |IQ|^2 instead of complex64 (BASELINE configs[0] layout)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        return run_reference_arm(args)
    globals()["SPS"] = args.sps
    globals()["DENSITY"] = args.density

    import torch
    import torch.distributed as dist
    import gr_adsb_b200 as g
    from gr_adsb_b200 import sharding, synth

    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")      # keep stdout to the one JSON line
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        log("warning: --gpus %d but WORLD_SIZE %d (launch with torch.distributed.run for N > 1)" % (args.gpus, world))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    L = 1 << args.log2n
    geo = sharding.shard_geometry(rank, world, L, SPS)
    buf = torch.zeros((geo["n_in"], 2), dtype=torch.float32, device=dev)
    shard = buf[geo["left"]: geo["left"] + L]
    _, truth = synth.synth_torch(L, SPS, int(L * DENSITY), SEED + rank, dev, out=shard, return_truth=False)
    is_c64 = 0 if args.f32 else (2 if args.sc16 else (3 if args.cu8 else 1))
    item_bytes = {0: 4, 1: 8, 2: 4, 3: 2}[is_c64]
    fmt_name = ("float32 |IQ|^2", "complex64 IQ", "int16 IQ", "uint8 IQ")[is_c64]
    if args.cu8:
        work_buf = torch.clamp(torch.round(buf * 128.0 + 127.5), 0, 255).to(torch.uint8).contiguous()
        shard = None
        del buf
        buf = work_buf
    if args.sc16:
        work_buf = torch.clamp(torch.round(buf * 32768.0), -32768, 32767).to(torch.int16).contiguous()
        shard = None
        del buf
        buf = work_buf
    if args.f32:
        work_buf = (buf[:, 0] * buf[:, 0] + buf[:, 1] * buf[:, 1]).contiguous()
        shard = None
        del buf
        buf = work_buf
    ctx = g.Context(SPS, geo["n_in"], device=local)
    if args.fec:
        ctx.set_error_correction(g.FEC_CONSERVATIVE)
    stream = torch.cuda.Stream(device=dev)      # a real (non-NULL) stream: kernels and timing events share it
    torch.cuda.set_stream(stream)
    sptr = stream.cuda_stream
    assert sptr != 0
    blank = geo["blank"]
    NONE = sharding.NONE
    rounds = [0, 0]          # carry rounds, re-resolves
    info_dev = torch.zeros(3, dtype=torch.int64, device=dev)
    gathered = torch.zeros(3 * world, dtype=torch.int64, device=dev)

    def step():
        if world > 1:
            sharding.exchange_halos(buf, geo, dist)
        ctx.burst_work_dev(buf.data_ptr(), is_c64, geo["n_in"], geo["n_out"], geo["lead"], THRESHOLD,
                           base=geo["base"], eob_in=-1, halo_left=geo["halo_left"], stream=sptr)
        if world == 1:
            return
        # blanking carry (framer.py:121-123,165) without a host round trip: every rank resolved assuming no
        # incoming state; export (first candidate, last hit) -> all-gather -> plan on the device -> the resolve
        # kernels re-run conditionally; a second round only checks that nothing is left to fix (sticky counter,
        # read after the timed loop)
        for final in (False, True):
            ctx.burst_export_carry_dev(g.CARRY_USED_ON_DEVICE, info_dev.data_ptr(), stream=sptr)
            dist.all_gather_into_tensor(gathered, info_dev)
            ctx.burst_carry_round_dev(gathered.data_ptr(), world, rank, final_check=final, stream=sptr)
            rounds[0] += 1

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    sync_all()
    rounds[0] = rounds[1] = 0
    sampler = ClockSampler(local)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(args.steps):
        step()
    ev1.record(stream)
    sync_all()
    clocks = sampler.stop()
    ms = ev0.elapsed_time(ev1)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    counters, _ = ctx.burst_finish(fetch_hits=False)
    if world > 1 and counters["carry_unstable"]:
        raise RuntimeError("blanking carry did not settle in one round on rank %d (%d steps): results would differ "
                           "from the whole-buffer oracle" % (rank, counters["carry_unstable"]))
    stats = torch.tensor([counters["n_hits"], counters["n_crc_ok"] + counters["n_fec_fixed"], counters["n_demod"],
                          counters["n_candidates"]],
                         dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(stats)
    n_hits, n_crc, n_demod, n_cand = (int(v) for v in stats.tolist())
    ms_per_step = ms / args.steps
    value = world * L / (ms_per_step * 1e-3) / 1e6

    # ---- per-kernel durations (CUDA events on the launch stream) for the roofline of the scan kernel
    ctx.profile_enable(True)
    scan_ms, res_ms = [], []
    for _ in range(min(args.steps, 50)):
        ctx.burst_work_dev(buf.data_ptr(), is_c64, geo["n_in"], geo["n_out"], geo["lead"], THRESHOLD,
                           base=geo["base"], eob_in=-1, halo_left=geo["halo_left"], stream=sptr)
        a, b = ctx.profile_last()
        scan_ms.append(a)
        res_ms.append(b)
    ctx.profile_enable(False)
    scan_avg = float(np.mean(scan_ms))
    peak, peak_src = measured_peak()
    alg_bytes = float(item_bytes) * geo["n_out"]   # 8 B per complex64 (4 B float32 / int16 IQ, 2 B uint8 IQ) sample, read once
    achieved = alg_bytes / (scan_avg * 1e-3) / 1e9
    traffic = ncu_traffic()
    roofline = {"bound": "hbm", "kernel": "scan_slice_kernel<sps=%d,%s>" % (SPS, ("float32", "complex64", "sc16", "cu8")[is_c64]), "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src,
                "traffic": (traffic or {}).get("dram_bytes_per_launch"),
                "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": scan_avg,
                "resolve_kernels_ms": float(np.mean(res_ms))}

    line = {"metric": METRIC, "value": value, "unit": "Msamples/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(L, world),
            "clocks": clocks, "gpu_launches": (5 if world == 1 else 5 + 2 + 2 + 3) * args.steps,
            "hits_per_step": n_hits, "candidates_per_step": n_cand, "crc_pass_per_step": n_crc,
            "crc_pass_bursts_per_s": n_crc / (ms_per_step * 1e-3),
            "roofline": roofline}
    if args.sps != 2 or args.density != 5e-3 or args.f32 or args.fec or args.sc16 or args.cu8:
        line["config"]["workload"] = ("secondary workload (not the headline): %d Msps %s, %.3g bursts/sample, threshold %.3g, "
                                      "whole-buffer semantics" % (SPS, fmt_name, DENSITY, THRESHOLD))
    if args.fec:
        line["error_correction"] = {"mode": "Conservative (decoder.py:738-763)", "fixed_per_step_rank0": counters["n_fec_fixed"]}
    if world > 1:
        line["carry"] = {"allgather_rounds_per_step": rounds[0] / args.steps, "host_syncs_per_step": 0,
                         "unsettled_steps": counters["carry_unstable"]}

    # ---- end to end through the C ABI with HOST buffers (H2D of the samples + D2H of the hits timed)
    if rank == 0 and not args.no_e2e and is_c64 in (1, 2, 3):
        try:
            src = buf[geo["left"]: geo["left"] + L]
            host = torch.empty(src.shape, dtype=src.dtype, pin_memory=True)
            host.copy_(src)
            torch.cuda.synchronize()
            hits = np.zeros(L // (63 * SPS) + 2, g.HIT_DTYPE)
            n_e2e = 0
            for i in range(args.e2e_steps + 1):
                if i == 1:
                    t0 = time.perf_counter()
                ctx.reset_state()
                n_e2e = ctx.burst_work_ptr(host.data_ptr(), is_c64, L, L, 8 * SPS - 1, THRESHOLD, 0, hits)
            dt = (time.perf_counter() - t0) / args.e2e_steps
            line["e2e"] = {"value": L / dt / 1e6, "unit": "Msamples/s", "h2d_bytes_per_step": L * item_bytes,
                           "d2h_bytes_per_step": int(n_e2e) * 32 + 96, "steps": args.e2e_steps,
                           "ms_per_step": dt * 1e3, "hits_per_step": int(n_e2e),
                           "api": "adsb_burst_work (pinned host %s in, adsb_hit[] out), 1 GPU"
                                  % fmt_name}
            del host
        except Exception as e:           # pragma: no cover
            line["e2e"] = {"error": repr(e)}

    # ---- bounded CPU baseline on the host cores (rank 0, N = 1)
    if rank == 0 and world == 1 and not args.no_cpu and is_c64 == 1:
        n_cpu = 1 << 25
        iq_cpu = shard[:n_cpu].cpu().numpy().view(np.complex64).reshape(-1)
        line["cpu_baseline"] = cpu_baseline_once(iq_cpu)
        line["cpu_baseline"]["host_cores_available"] = os.cpu_count()

    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    ctx.close()


if __name__ == "__main__":
    main()
