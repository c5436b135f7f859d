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
    ap.add_argument("--chunk", type=int, default=1 << 26, help="samples per device chunk (three are resident)")
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
    # Stream the file through the device: the reads go straight into the context's pinned staging buffer
    # (adsb_stream_acquire / adsb_stream_commit), chunk k is scanned while chunk k+1 crosses PCIe, and the result is
    # what ONE framer + demod call over the whole file would give -- for recordings of any size, not just those that fit
    # in host or device memory.
    fmt = {"fc32": g.FMT_C64, "sc16": g.FMT_SC16, "cu8": g.FMT_CU8}[a.format]
    item_bytes = {"fc32": 8, "sc16": 4, "cu8": 2}[a.format]
    n_items = os.path.getsize(a.path) // item_bytes
    chunk = max(1 << 14, min(a.chunk, n_items))
    ctx = g.Context(sps, chunk)
    if a.fec:
        ctx.set_error_correction(g.FEC_CONSERVATIVE)
    ctx.stream_open(fmt, a.threshold, chunk)
    parts = []
    with open(a.path, "rb", buffering=0) as f:
        left = n_items
        while left:
            dst = ctx.stream_buffer()                       # free part of the pinned staging buffer
            want = min(left, len(dst))
            got = f.readinto(memoryview(dst[:want]).cast("B")) // item_bytes
            if got == 0:
                break
            ctx.stream_commit(got)
            left -= got
            parts.append(ctx.stream_pop())
    ctx.stream_flush()
    parts.append(ctx.stream_pop())
    hits = np.concatenate(parts)
    st = ctx.stream_stats()
    n_demod = int(((hits["flags"] & g.HIT_DEMOD_VALID) != 0).sum())
    n_fixed = int(((hits["flags"] & g.HIT_FEC_FIXED) != 0).sum())
    print("%d samples in %d chunks: %d bursts tagged, %d demodulated, %d pass parity (%d of them repaired)" % (
        st["items"], st["chunks_done"], len(hits), n_demod, st["crc_ok"], n_fixed))
    ok = hits[(hits["flags"] & g.HIT_CRC_OK) != 0]
    for h, bits in list(zip(ok, g.unpack_bits(ok)))[: a.max]:
        print("  offset %10d  t=%.6f s  snr %5.1f dB  %s%s" % (
            h["idx"] - (8 * sps - 1), (h["idx"] - (8 * sps - 1)) / a.fs, h["snr"], synth.bits_to_hex(bits),
            "  (repaired)" if h["flags"] & g.HIT_FEC_FIXED else ""))


if __name__ == "__main__":
    main()
