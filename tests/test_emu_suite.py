"""The `gpu` parity tests once more on the GPU-less machine, against the host emulation of the CUDA sources
(tests/emu: the same .cu/.cuh files compiled with g++, every CUDA thread a fiber, warp collectives and barriers as
rendezvous, shuffled thread schedule).  This catches logic errors in the kernels and in the C ABI's host code
before a GPU run; the parity claim itself rests on the `-m gpu` run on the B200.  Nothing in gr_adsb_b200 loads
the emulated library: only this subprocess does, through ADSB_EMU=1 (tests/conftest.py)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("seed", [1])
def test_gpu_suite_under_host_emulation(seed):
    env = dict(os.environ, ADSB_EMU="1", EMU_SEED=str(seed))
    env.pop("ADSB_B200_LIB", None)
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests"), "-m", "gpu", "-x", "-q",
                        "-p", "no:cacheprovider"], cwd=ROOT, env=env, stdout=subprocess.PIPE,
                       stderr=subprocess.STDOUT, text=True, timeout=1500)
    tail = "\n".join(r.stdout.splitlines()[-25:])
    assert r.returncode == 0, tail
    assert " passed" in tail and "failed" not in tail, tail
