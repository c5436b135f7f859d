"""CPU-side checks of the drop-in boundary: the shared library loads, exports exactly what
include/adsb_b200.h declares, and the product refuses to run without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

import gr_adsb_b200 as g
from gr_adsb_b200 import _native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "adsb_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(adsb_[a-z_0-9]+)\s*\(", src)))


def test_header_symbols_exported():
    _native.build()
    lib = ctypes.CDLL(_native.LIB_PATH)
    names = _declared()
    assert len(names) >= 14
    for n in names:
        assert hasattr(lib, n), "missing export " + n
    assert sorted(_native.EXPORTS) == names
    assert lib.adsb_abi_version() == 1


def test_hit_record_layout():
    assert g.HIT_DTYPE.itemsize == 32
    assert g.HIT_DTYPE.fields["idx"][1] == 0 and g.HIT_DTYPE.fields["snr"][1] == 8
    assert g.HIT_DTYPE.fields["flags"][1] == 12 and g.HIT_DTYPE.fields["bits"][1] == 16
    assert ctypes.sizeof(_native.Counters) == 80


def test_strerror_and_bad_args():
    L = g.lib()
    assert L.adsb_strerror(0) == b"ok"
    assert b"device" in L.adsb_strerror(-5)
    h = ctypes.c_void_p()
    assert L.adsb_create(ctypes.byref(h), 0, 3, 1024) == -1      # odd sps (framer.py:137 needs sps//2 taps)
    assert L.adsb_create(ctypes.byref(h), 0, 2, 0) == -1
    assert L.adsb_reset_state(None) == -1


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(g.AdsbError) as e:
        g.Context(2, 1 << 16)
    assert e.value.code == -5
    with pytest.raises(g.AdsbError):
        g.framer(2e6, 0.01)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "gr_adsb_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "oracle/" not in txt, f


def test_synth_crc_roundtrip():
    from gr_adsb_b200 import synth
    rng = np.random.default_rng(0)
    m = synth.df17_messages(rng, 50)
    assert np.all(synth.crc24_syndrome(m) == 0)
    m[:, 40] ^= 1
    assert np.all(synth.crc24_syndrome(m) != 0)
    assert synth.bits_to_hex(synth.hex_to_bits("8D4840D6202CC371C32CE0576098")) == "8D4840D6202CC371C32CE0576098"


def test_emulation_library_is_refused_outside_test_mode(tmp_path):
    """The host-emulation build of the kernels (tests/emu) is test infrastructure: the package must not load it as
    its compute path unless the test switch ADSB_EMU=1 is set."""
    import subprocess
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from emu import build_emu
    lib = build_emu.build()
    code = "import gr_adsb_b200 as g; g.Context(2, 1 << 16)"
    env = dict(os.environ, ADSB_B200_LIB=lib)
    env.pop("ADSB_EMU", None)
    r = subprocess.run([sys.executable, "-c", code], cwd=ROOT, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode != 0 and "host-emulation TEST library" in r.stdout
