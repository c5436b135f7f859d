"""The GRC block definitions and the `gnuradio.adsb` import shim name what the package really exports (the contract files
in gr-adsb: grc/adsb_framer.block.yml:5-9, grc/adsb_demod.block.yml:5-7, python/adsb/__init__.py:24-26)."""
import inspect
import os
import re
import subprocess
import sys

import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_grc_definitions_match_the_blocks():
    import gr_adsb_b200 as g
    d = yaml.safe_load(open(os.path.join(ROOT, "grc", "adsb_b200_burst.block.yml")))
    assert d["file_format"] == 1 and d["templates"]["imports"] == "import gr_adsb_b200"
    make = d["templates"]["make"]
    assert make.startswith("gr_adsb_b200.burst(")
    used = set(re.findall(r"\$\{(\w+)\}", make))
    assert used == {p["id"] for p in d["parameters"]}
    sig = inspect.signature(g.burst.__init__)
    for kw in re.findall(r"(\w+)=\$\{", make):
        assert kw in sig.parameters
    assert d["templates"]["callbacks"] == ["set_threshold(${threshold})"] and hasattr(g.burst, "set_threshold")
    assert [o["label"] for o in d["outputs"]] == ["demodulated", "out"]          # as grc/adsb_demod.block.yml
    u = yaml.safe_load(open(os.path.join(ROOT, "grc", "adsb_b200_unbatch.block.yml")))
    assert u["templates"]["make"] == "gr_adsb_b200.pdu_unbatch()" and callable(g.pdu_unbatch)
    assert u["inputs"][0]["label"] == "batch" and u["outputs"][0]["label"] == "demodulated"


def test_gnuradio_adsb_shim_exports_the_reference_names():
    code = ("import sys; sys.path.insert(0, %r); sys.path.insert(0, %r); import gnuradio.adsb as adsb; "
            "import gr_adsb_b200 as g; assert adsb.framer is g.framer and adsb.demod is g.demod and adsb.burst is g.burst; "
            "print('ok')" % (ROOT, os.path.join(ROOT, "gnuradio_shim")))
    r = subprocess.run([sys.executable, "-c", code], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 0 and r.stdout.strip() == "ok", r.stderr
