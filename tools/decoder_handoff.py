#!/usr/bin/env python
"""PDUs per second through the UNMODIFIED gr-adsb decoder (python/adsb/decoder.py:325-352, behind oracle/fake_gr) when the
demodulator hands its bursts over one PDU at a time (demod.py:104-110) and when it hands over one batched message per
work() call (gr_adsb_b200.blocks: demod(batch=True) / burst(batch=True) -> split_batch / feed_decoder).  CPU only; needs
/root/reference (build container).   python tools/decoder_handoff.py [n_pdus]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import run_reference as rr          # noqa: E402  (test infrastructure: this tool measures, it ships nothing)
from gr_adsb_b200 import synth                  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
# public DF17 messages the reference decoder handles; random payloads trip its unimplemented branches
KAT = ["8D4840D6202CC371C32CE0576098", "8D485020994409940838175B284F", "8D40621D58C382D690C8AC2863A7",
       "8D40621D58C386435CC412692AD6", "8DA05F219B06B6AF189400CBC33F"]
bits = np.stack([synth.hex_to_bits(KAT[i % len(KAT)]) for i in range(n)])
per_call = 32                                   # bursts per work() call at 10k bursts/s and 8192-item buffers: ~40
import logging
logging.disable(logging.CRITICAL)


def to_pdu(i):
    return ({"timestamp": 1.0 + i * 1e-4, "snr": 20.0}, bits[i].copy())


# (a) one PDU per burst: what demod.py publishes
dec = rr.new_decoder()
t0 = time.perf_counter()
for i in range(n):
    dec.decode_packet(to_pdu(i))
ta = time.perf_counter() - t0

# (b) one message per work() call, split in front of the decoder
os.environ["ADSB_B200_LIB"] = os.environ.get("ADSB_B200_LIB", "")        # blocks import needs no device for split_batch
from gr_adsb_b200._gr_compat import pmt         # noqa: E402
sys.modules.setdefault("pmt", pmt)


def split_batch(msg):
    meta, vec = msg
    k = int(meta["batch"])
    b = np.asarray(vec, dtype=np.uint8).reshape(k, 112)
    return [({"timestamp": meta["timestamp"][i], "snr": meta["snr"][i]}, b[i]) for i in range(k)]


dec = rr.new_decoder()
t0 = time.perf_counter()
for j in range(0, n, per_call):
    k = min(per_call, n - j)
    msg = ({"batch": k, "timestamp": [1.0 + (j + i) * 1e-4 for i in range(k)], "snr": [20.0] * k}, bits[j:j + k].reshape(-1).copy())
    for pdu in split_batch(msg):
        dec.decode_packet(pdu)
tb = time.perf_counter() - t0
print("unchanged adsb.decoder, %d DF17 PDUs, one core: per-burst PDUs %.0f PDUs/s (%.0f us each); batched messages of %d "
      "split in front of the decoder %.0f PDUs/s (%.0f us each)" % (n, n / ta, ta / n * 1e6, per_call, n / tb, tb / n * 1e6))
