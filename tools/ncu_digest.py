#!/usr/bin/env python
"""Digest an ncu report of the scan kernel: headline metrics, instructions per barrier-delimited
phase, opcode mix, hottest source lines.  Usage: tools/ncu_digest.py report.ncu-rep <mangled-kernel-substring>"""
import collections
import csv
import io
import os
import re
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "scan_slice_kernelILi2ELb1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.environ.get("ADSB_B200_LIB", os.path.join(ROOT, "gr_adsb_b200", "csrc", "libadsb_b200.so"))   # the build the report was taken with


def run(cmd):
    return subprocess.run(cmd, shell=True, capture_output=True, text=True).stdout


raw = list(csv.reader(io.StringIO(run("ncu -i %s --page raw --csv" % rep))))
hdr, units, vals = raw[0], raw[1], raw[2]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.avg", "launch__grid_size",
        "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        # the L1 / shared-memory data pipe (LDS, STS, LDG, SHFL wavefronts) and the ALU pipe: what bounds this kernel
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "memory_l1_wavefronts_shared",
        "memory_l1_wavefronts_shared_ideal", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"]
M = {}
for i, h_ in enumerate(hdr):
    if h_ in want:
        M[h_] = (vals[i], units[i])
        print("%-70s %s %s" % (h_, vals[i], units[i]))

os.makedirs("/tmp/cub", exist_ok=True)
run("cd /tmp/cub && rm -f *.cubin && cuobjdump -xelf all %s >/dev/null 2>&1 && nvdisasm -g -c *.cubin > all.sass" % lib)
lines = open("/tmp/cub/all.sass").read().split("\n")
start = next(i for i, l in enumerate(lines) if l.startswith(".text.") and kern in l and l.endswith(":"))
end = next(i for i in range(start + 1, len(lines)) if lines[i].startswith("//--------------------- .text.") or lines[i].startswith("//--------------------- .nv"))
cur, insts = None, []
for l in lines[start:end]:
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        insts.append((cur, m.group(2)))
rows = list(csv.reader(io.StringIO(run("ncu -i %s --page source --csv" % rep))))
h2 = rows[1]
ci = {h_: i for i, h_ in enumerate(h2)}
data = rows[2:]
assert len(data) == len(insts), (len(data), len(insts))
tot = sum(int(r[ci["Instructions Executed"]]) for r in data)
tots = sum(int(r[ci["# Samples"]]) for r in data)
grid = float(M.get("launch__grid_size", ("1", ""))[0])
print("total warp inst %d  (%.0f per CTA), stall samples %d" % (tot, tot / grid, tots))
seg, segs, segsamp, segstart = 0, collections.Counter(), collections.Counter(), {}
op = collections.Counter()
agg, samp = collections.Counter(), collections.Counter()
stalls = collections.Counter()
stall_cols = [c for c in h2 if c.startswith("stall_") and "Not Issued" not in c]
for r, (loc, txt) in zip(data, insts):
    n, s = int(r[ci["Instructions Executed"]]), int(r[ci["# Samples"]])
    if "BAR.SYNC" in txt or "BAR.RED" in txt:
        seg += 1
        segstart[seg] = loc
    segs[seg] += n
    segsamp[seg] += s
    t = txt.split()
    o = t[1] if t[0].startswith("@") else t[0]
    op[o.split(".")[0]] += n
    agg[loc] += n
    samp[loc] += s
    for c in stall_cols:
        try:
            stalls[c] += int(r[ci[c]])
        except ValueError:
            pass
for k in sorted(segs):
    print("  after barrier %2d (%s): %5.1f%% inst (%6.0f/CTA)  %5.1f%% stall samples" % (
        k, segstart.get(k), 100.0 * segs[k] / tot, segs[k] / grid, 100.0 * segsamp[k] / max(tots, 1)))
print("opcodes:", [(k, round(100.0 * v / tot, 1)) for k, v in op.most_common(22)])
ts = sum(stalls.values())
print("stalls:", [(k, round(100.0 * v / max(ts, 1), 1)) for k, v in stalls.most_common(10)])
src = {}
for f in ("adsb_scan.cuh", "adsb_device.cuh"):
    src[f] = open(os.path.join(ROOT, "gr_adsb_b200", "csrc", f)).read().split("\n")
for k, v in agg.most_common(int(os.environ.get("TOP", "28"))):
    f, ln = k if k else ("?", 0)
    text = src[f][ln - 1].strip()[:88] if f in src else ""
    print("%5.2f%% inst %5.2f%% stall  %s:%d  %s" % (100.0 * v / tot, 100.0 * samp[k] / max(tots, 1), f, ln, text))
