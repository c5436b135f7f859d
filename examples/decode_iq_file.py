#!/usr/bin/env python
"""Detect and demodulate ADS-B bursts in a recorded IQ file on a B200 (whole-buffer semantics).

    python examples/decode_iq_file.py capture.fc32 --fs 2e6 --threshold 0.01 [--fec] [--max 20]

The file holds interleaved float32 I/Q (GNU Radio file_sink of a complex stream; what
examples/adsb_rx.py's USRP source produces); `--format sc16` / `--format cu8` read interleaved int16 / uint8 I/Q
instead (USRP wire format / rtl_sdr output), widened on the device.  For every burst whose parity checks out it prints the
sample offset, SNR and the 112-bit message as hex -- the input the reference adsb.decoder works on.
`--synth N` writes a synthetic N-sample capture first (handy where no recording is at hand).
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gr_adsb_b200 as g                      # noqa: E402
from gr_adsb_b200 import synth                # noqa: E402


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("path")
    ap.add_argument("--fs", type=float, default=2e6)
    ap.add_argument("--threshold", type=float, default=0.01)
    ap.add_argument("--fec", action="store_true", help='decoder error_corr="Conservative" on the GPU')
    ap.add_argument("--max", type=int, default=20, help="messages to print")
    ap.add_argument("--format", default="fc32", choices=["fc32", "sc16", "cu8"], help="sample format of the file")
    ap.add_argument("--synth", type=int, default=0, help="first write a synthetic capture of this many samples")
    a = ap.parse_args()
    sps = int(a.fs // 1e6)
    if a.synth:
        d = synth.synth_numpy(a.synth, sps=sps, n_bursts=max(1, a.synth // 2000), seed=1, nonoverlap=True)
        ri = d["iq"].view(np.float32).reshape(-1, 2)
        if a.format == "sc16":
            np.clip(np.round(ri * 32768.0), -32768, 32767).astype(np.int16).tofile(a.path)
        elif a.format == "cu8":
            np.clip(np.round(ri * 128.0 + 127.5), 0, 255).astype(np.uint8).tofile(a.path)
        else:
            d["iq"].tofile(a.path)
    if a.format == "fc32":
        iq = np.fromfile(a.path, dtype=np.complex64)
    else:
        iq = np.fromfile(a.path, dtype=np.int16 if a.format == "sc16" else np.uint8).reshape(-1, 2)
    ctx = g.Context(sps, len(iq))
    if a.fec:
        ctx.set_error_correction(g.FEC_CONSERVATIVE)
    hits = ctx.burst_work(iq, threshold=a.threshold)
    c = ctx.counters()
    print("%d samples: %d bursts tagged, %d demodulated, %d pass parity (+%d repaired)" % (
        len(iq), c["n_hits"], c["n_demod"], c["n_crc_ok"], c["n_fec_fixed"]))
    ok = hits[(hits["flags"] & g.HIT_CRC_OK) != 0]
    for h, bits in list(zip(ok, g.unpack_bits(ok)))[: a.max]:
        print("  offset %10d  t=%.6f s  snr %5.1f dB  %s%s" % (
            h["idx"] - (8 * sps - 1), (h["idx"] - (8 * sps - 1)) / a.fs, h["snr"], synth.bits_to_hex(bits),
            "  (repaired)" if h["flags"] & g.HIT_FEC_FIXED else ""))


if __name__ == "__main__":
    main()
