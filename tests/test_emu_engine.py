"""The host emulation (tests/emu) is itself test infrastructure: check its collectives and barriers against their
CUDA definitions, so that a pass of the emulated parity suite means what it says."""
import os
import subprocess

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu")


def test_emulated_collectives_match_cuda_semantics(tmp_path):
    exe = str(tmp_path / "emu_selftest")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-DADSB_HOST_EMU", "-I" + os.path.join(HERE, "include"),
                           os.path.join(HERE, "selftest.cpp"), os.path.join(HERE, "cuda_emu.cpp"), "-o", exe])
    for seed in ("0", "1", "12345"):
        r = subprocess.run([exe], env=dict(os.environ, EMU_SEED=seed), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        assert r.returncode == 0 and "emu selftest ok" in r.stdout, r.stdout[-2000:]
