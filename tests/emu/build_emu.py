"""TEST INFRASTRUCTURE ONLY: compile gr_adsb_b200/csrc (the CUDA sources, unchanged) with g++ against the
fiber-based stand-in for the CUDA runtime in tests/emu, into tests/emu/libadsb_emu.so.  The gr_adsb_b200
package never loads this library; only tests do (tests/conftest.py, ADSB_EMU=1)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "gr_adsb_b200", "csrc")
# extra -D flags (kernel variants under test), e.g. ADSB_EMU_DEFS="-DADSB_OPT_COMPACT_W0=1"
DEFS = os.environ.get("ADSB_EMU_DEFS", "").split()
LIB = os.path.join(HERE, "libadsb_emu%s.so" % ("_" + "".join(c if c.isalnum() else "_" for c in "".join(DEFS)) if DEFS else ""))

FLAGS = ["-std=c++17", "-O1", "-g", "-fPIC", "-shared", "-DADSB_HOST_EMU", "-ffp-contract=off",
         "-Wall", "-Wno-unused-variable", "-Wno-unknown-pragmas", "-I" + os.path.join(HERE, "include")]


def build(force=False):
    srcs = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cu", ".cuh"))]
    srcs += [os.path.join(HERE, "cuda_emu.cpp"), os.path.join(HERE, "include", "cuda_runtime.h"),
             os.path.join(ROOT, "include", "adsb_b200.h")]
    if not force and os.path.isfile(LIB) and all(os.path.getmtime(LIB) >= os.path.getmtime(s) for s in srcs):
        return LIB
    cmd = [os.environ.get("CXX", "g++")] + FLAGS + DEFS + ["-x", "c++", os.path.join(CSRC, "adsb_api.cu"),
                                                    os.path.join(HERE, "cuda_emu.cpp"), "-o", LIB]
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force=True))
