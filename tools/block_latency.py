#!/usr/bin/env python
"""Per-work()-call cost of the drop-in blocks at GNU-Radio-like buffer sizes (host buffers, PCIe both ways).
Usage: python tools/block_latency.py  ->  one line per chunk size: Msamples/s through framer+demod."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import gr_adsb_b200 as g
from gr_adsb_b200 import synth
from harness import run_blocks

n = 1 << 22
d = synth.synth_numpy(n, sps=2, n_bursts=400, seed=3)          # BASELINE configs[0] density (100 bursts/s)
x = synth.mag2(d["iq"])
for chunk in (8192, 65536, 1 << 20):
    fr, dm = g.framer(2e6, 0.01), g.demod(2e6)
    sizes = [chunk] * (n // chunk)
    run_blocks(fr, dm, x[: chunk * 4], sizes[:4])              # warm up
    fr, dm = g.framer(2e6, 0.01), g.demod(2e6)
    t0 = time.perf_counter()
    r = run_blocks(fr, dm, x, sizes)
    dt = time.perf_counter() - t0
    print("work() chunk %8d: %7.1f Msamples/s, %6.1f us per framer+demod call pair, %d tags, %d PDUs"
          % (chunk, n / dt / 1e6, dt / len(sizes) * 1e6, len(r["tags"]), len(r["pdus"])))
