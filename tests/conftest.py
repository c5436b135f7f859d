import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

# ADSB_EMU=1: run the `gpu` tests on a GPU-less machine against the host emulation of the CUDA sources
# (tests/emu: same .cu/.cuh files, every CUDA thread a fiber).  A logic check of the kernels and of the C ABI's
# host code before spending GPU time; it is not a product path and proves nothing about the real device.
EMU = os.environ.get("ADSB_EMU") == "1"
if EMU:
    from emu import build_emu
    os.environ["ADSB_B200_LIB"] = build_emu.build()


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _have_gpu():
    """The product needs only the CUDA runtime (no torch): ask the library itself."""
    try:
        from gr_adsb_b200 import _native
        return int(_native.lib().adsb_device_count()) > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if EMU:
        skip = pytest.mark.skip(reason="ADSB_EMU=1 runs only the gpu-marked tests (against the emulation)")
        for it in items:
            if "gpu" not in it.keywords:
                it.add_marker(skip)
        return
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)
