#!/usr/bin/env python
"""How much does the emulated gpu suite (tests/emu) actually see?  Apply small source mutations to the kernels, one at
a time, rebuild the host emulation and run the parity tests: every mutant should be killed (a failing test or an abort
of the emulation).  Works on a scratch copy of the repository; the tree it is started from is not touched.

    python tools/mutation_check.py            # prints one line per mutant and a summary
"""
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# (file, original text, mutated text, what it breaks)
MUTANTS = [
    ("adsb_scan.cuh", "    // (ranks 49, 50 of 100) are ranks 63, 64 of the 128, the two middle ones\n    constexpr int kLowPads = (128 - kNoise) / 2;\n#pragma unroll\n    for (int r = 0; r < 16; r++) {",
     "    // (ranks 49, 50 of 100) are ranks 63, 64 of the 128, the two middle ones\n    constexpr int kLowPads = (128 - kNoise) / 2 - 1;\n#pragma unroll\n    for (int r = 0; r < 16; r++) {", "median padding off by one (quad medians)"),
    ("adsb_scan.cuh", "C |= (acc & ~nxt) >> ((L + 1) >> 1);", "C |= (acc & ~nxt) >> (L >> 1);", "pulse centre rounding (framer.py:113)"),
    ("adsb_scan.cuh", "Q = ~Q << (LB - G::DB);", "Q = ~Q << (LB - 1 - G::DB);", "pre-filter bit alignment"),
    ("adsb_scan.cuh", "int off = base + before + excl;", "int off = base + before + excl + (lane == 5);", "compaction offset of one lane"),
    ("adsb_scan.cuh", "const unsigned sel = (lane & 16) ? 0x7632u : 0x5410u;", "const unsigned sel = (lane & 16) ? 0x7632u : 0x5401u;", "bit packing byte order"),
    ("adsb_scan.cuh", "if (p1_pairs) {\n                if (u & 1) {", "if (p1_pairs) {\n                if (!(u & 1)) {", "bitmap pair store parity"),
    ("adsb_scan.cuh", "if (cnt == 1 && lmask[ci] == 0) {", "if (cnt >= 1 && lmask[ci] == 0) {", "register centre used for multi-candidate threads"),
    ("adsb_device.cuh", "lo_mid = __shfl_sync(kFull, m, lane & 16);", "lo_mid = __shfl_sync(kFull, m, (lane & 16) | 8);", "pair medians (4 / 8 Msps): lower middle rank taken from the upper half"),
    ("adsb_device.cuh", "cex(k[2], k[4]); cex(k[3], k[5]);\n    cex(k[1], k[2]); cex(k[3], k[4]); cex(k[5], k[6]);\n#pragma unroll", "cex(k[2], k[4]);\n    cex(k[1], k[2]); cex(k[3], k[4]); cex(k[5], k[6]);\n#pragma unroll", "pair medians: one comparator of the 19-comparator network missing"),
    ("adsb_device.cuh", "const unsigned o = __shfl_xor_sync(kFull, k[7 - r], 15);       // position e <-> 127 - e", "const unsigned o = __shfl_xor_sync(kFull, k[r], 15);", "pair medians: mirror exchange register order"),
    ("adsb_device.cuh", "cex(k[4], k[8]); cex(k[6], k[10]); cex(k[5], k[9]); cex(k[7], k[11]);", "cex(k[4], k[8]); cex(k[6], k[10]); cex(k[5], k[9]);", "quad medians: one comparator of the 63-comparator network missing"),
    ("adsb_device.cuh", "const unsigned o = __shfl_xor_sync(kFull, k[15 - r], 7);       // position e <-> 127 - e", "const unsigned o = __shfl_xor_sync(kFull, k[15 - r], 6);", "quad medians: final exchange with the wrong lane"),
    ("adsb_device.cuh", "const unsigned oh = __shfl_xor_sync(kFull, k[15 - r], lm);", "const unsigned oh = __shfl_xor_sync(kFull, k[r], lm);", "quad medians: mirror exchange register order"),
    ("adsb_scan.cuh", "const unsigned mo = __shfl_xor_sync(kFull, m, 4);", "const unsigned mo = __shfl_xor_sync(kFull, m, 2);", "quad medians: second middle rank from the wrong lane"),
    ("adsb_scan.cuh", "if ((anyn >> (lane & 24)) & 0xffu) med", "if ((anyn >> (lane & 16)) & 0xffu) med", "quad medians: NaN vote read from another candidate's lanes"),
    ("adsb_scan.cuh", "const int n_quads = (SPS != 0 && fast) ? ((n_cand + 3) >> 2) : 0;", "const int n_quads = (SPS != 0 && fast) ? (n_cand >> 2) : 0;", "quad medians: remainder candidates get no median task"),
    ("adsb_scan.cuh", "if (gb.sync_or(have_long)) {", "if (gb.sync_or(false)) {", "long-run pass never taken"),
    ("adsb_resolve.cuh", "if (cj > last + D) {", "if (cj >= last + D) {", "blanking boundary (framer.py:121)"),
    ("adsb_scan.cuh", "    gb.sync();\n    slice_tile<SPS, FMT, NT>", "    slice_tile<SPS, FMT, NT>", "barrier between compaction and slicing removed"),
]


def main():
    tmp = tempfile.mkdtemp(prefix="adsb_mut_")
    work = os.path.join(tmp, "repo")
    shutil.copytree(ROOT, work, ignore=shutil.ignore_patterns(".git", "gpurun_out", "*.so", "__pycache__", ".hypothesis", ".pytest_cache"))
    env = dict(os.environ, ADSB_EMU="1")
    env.pop("ADSB_B200_LIB", None)
    tests = ["tests/test_gpu_parity.py", "tests/test_gpu_z_adversarial.py", "tests/test_gpu_u8_ingest.py"]
    killed = 0
    for fname, a, b, what in MUTANTS:
        path = os.path.join(work, "gr_adsb_b200", "csrc", fname)
        src = open(path).read()
        if src.count(a) != 1:
            print("SKIP   %-70s (pattern occurs %d times)" % (what, src.count(a)))
            continue
        open(path, "w").write(src.replace(a, b))
        r = subprocess.run([sys.executable, "-m", "pytest"] + tests + ["-m", "gpu", "-x", "-q", "-p", "no:cacheprovider"],
                           cwd=work, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        open(path, "w").write(src)
        dead = r.returncode != 0
        killed += dead
        last = [l for l in r.stdout.strip().splitlines() if l.strip()][-1][:90] if r.stdout.strip() else ""
        print("%s %-70s %s" % ("KILLED " if dead else "SURVIVED", what, last), flush=True)
    print("%d of %d mutants killed" % (killed, len(MUTANTS)))
    shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    main()
