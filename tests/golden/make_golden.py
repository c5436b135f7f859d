#!/usr/bin/env python
"""Generate the golden vectors in this directory by running the UNMODIFIED reference.

Runs only where `/root/reference` exists (the build container):

    python tests/golden/make_golden.py

Each `<case>.npz` stores the exact input stream, the parameters, and -- for every chunking of
the stream into work() calls -- what reference framer.py + demod.py produced through
`oracle/run_reference.run_stream`: tag offsets + SNR, PDU offsets + SNR + 112 bits, and the
framer's (prev_in0, prev_eob_idx) after every call.  The reference publishes no vectors of
its own (qa_framer.py / qa_demod.py are empty), so these files are what "parity" means.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from gr_adsb_b200 import synth          # noqa: E402
from oracle import run_reference as rr  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))

PUBLIC_DF17 = ["8D4840D6202CC371C32CE0576098", "8D485020994409940838175B284F",
               "8D40621D58C382D690C8AC2863A7", "8D40621D58C386435CC412692AD6",
               "8DA05F219B06B6AF189400CBC33F", "8D4840D6202CC371C32CE0576099"]


def chunk_list(total, chunk):
    return rr.chunk_sizes(total, chunk)


def emit(name, x, fs, threshold, chunkings, iq=None, thresholds=None, note=""):
    rec = {"x": np.ascontiguousarray(x, np.float32), "fs": np.float64(fs),
           "threshold": np.float64(threshold), "n_chunkings": np.int64(len(chunkings)),
           "note": np.array(note)}
    if iq is not None:
        rec["iq"] = np.ascontiguousarray(iq, np.complex64)
    dec = rr.new_decoder()
    for ci, ch in enumerate(chunkings):
        sizes = chunk_list(len(x), ch)
        thr_list = None
        if thresholds is not None:
            thr_list = [thresholds[k % len(thresholds)] for k in range(len(sizes))]
            rec["c%d_thresholds" % ci] = np.array(thr_list, np.float64)
        r = rr.run_stream(x, fs, threshold, sizes, thresholds=thr_list)
        rec["c%d_sizes" % ci] = np.array(sizes, np.int64)
        rec["c%d_tag_off" % ci] = np.array([t[0] for t in r["tags"]], np.int64)
        rec["c%d_tag_snr" % ci] = np.array([t[1] for t in r["tags"]], np.float64)
        rec["c%d_pdu_off" % ci] = np.array([p[0] for p in r["pdus"]], np.int64)
        rec["c%d_pdu_snr" % ci] = np.array([p[1] for p in r["pdus"]], np.float64)
        bits = np.array([p[2] for p in r["pdus"]], np.uint8).reshape(-1, 112)
        rec["c%d_pdu_bits" % ci] = np.packbits(bits, axis=1)
        rec["c%d_pdu_crc" % ci] = np.array([rr.crc_pass(b, dec) for b in bits], np.uint8)
        rec["c%d_state_in0" % ci] = np.array([s[0] for s in r["state"]], np.float32)
        rec["c%d_state_eob" % ci] = np.array([s[1] for s in r["state"]], np.int64)
        print("%-22s chunk=%-14s tags=%-5d pdus=%-5d crc_ok=%d" % (
            name, str(ch) if not isinstance(ch, list) else "list[%d]" % len(ch),
            len(r["tags"]), len(r["pdus"]), int(rec["c%d_pdu_crc" % ci].sum())))
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **rec)


def main():
    # 1. BASELINE config 1 in miniature: float32 |IQ|^2, 2 Msps, clean non-overlapping DF17 bursts
    d = synth.synth_numpy(1 << 17, sps=2, n_bursts=24, seed=1090, nonoverlap=True,
                          bits=np.concatenate((np.array([synth.hex_to_bits(h) for h in PUBLIC_DF17]),
                                               synth.df17_messages(np.random.default_rng(5), 18))))
    emit("sps2_clean", synth.mag2(d["iq"]), 2e6, 0.01, [None, 8192, 4096, 65536, 1000],
         note="config-1 style; first 6 bursts are public DF17 messages (last one has a flipped bit)")

    # 2. 8 Msps (4 samples per half symbol)
    d = synth.synth_numpy(1 << 18, sps=8, n_bursts=20, seed=8008, nonoverlap=True)
    emit("sps8_clean", synth.mag2(d["iq"]), 8e6, 0.01, [None, 8192, 65536, 30000])

    # 3. 4 Msps with GR-like 8192 chunks: exhibits the stale prev_eob_idx state
    d = synth.synth_numpy(1 << 17, sps=4, n_bursts=30, seed=4004)
    emit("sps4_stale_state", synth.mag2(d["iq"]), 4e6, 0.01, [None, 8192, 2048])

    # 4. noisy: 19 % of samples above threshold, random (overlapping) bursts
    d = synth.synth_numpy(1 << 16, sps=2, n_bursts=60, seed=66, noise_var=6e-3)
    emit("sps2_noisy", synth.mag2(d["iq"]), 2e6, 0.01, [None, 8192, 1500])

    # 5. threshold far below the noise floor: long above-threshold runs
    d = synth.synth_numpy(1 << 16, sps=2, n_bursts=60, seed=67, noise_var=1e-3)
    emit("sps2_lowthr", synth.mag2(d["iq"]), 2e6, 0.0005, [None, 8192])
    d = synth.synth_numpy(1 << 16, sps=8, n_bursts=20, seed=68, noise_var=1e-3)
    emit("sps8_lowthr", synth.mag2(d["iq"]), 8e6, 0.0002, [None, 8192])

    # 6. dense overlapping bursts (BASELINE config 4 density: 0.05 bursts / sample)
    d = synth.synth_numpy(1 << 16, sps=2, n_bursts=3277, seed=404)
    emit("sps2_dense", synth.mag2(d["iq"]), 2e6, 0.01, [None, 8192])
    # config-2 density (5e-3 bursts / sample), complex64 input kept for the fused path
    d = synth.synth_numpy(1 << 16, sps=2, n_bursts=328, seed=202)
    emit("sps2_cfg2_iq", synth.mag2(d["iq"]), 2e6, 0.01, [None, 8192], iq=d["iq"])
    d = synth.synth_numpy(1 << 17, sps=8, n_bursts=82, seed=303)
    emit("sps8_cfg3_iq", synth.mag2(d["iq"]), 8e6, 0.01, [None, 16384], iq=d["iq"])

    # 7. edge cases
    n = 1 << 14
    emit("edge_zeros", np.zeros(n, np.float32), 2e6, 0.01, [None, 4096])
    emit("edge_thr0_allabove", np.abs(synth.mag2(synth.synth_numpy(n, 2, 5, 1)["iq"])), 2e6, 0.0, [None, 4096])
    x = np.full(n, 0.5, np.float32)
    x[5000:5003] = 0.0
    emit("edge_constant_high", x, 2e6, 0.01, [None, 4096])
    # burst right at the start of the stream and one whose first pulse is window index 0 of call 2
    bits = synth.df17_messages(np.random.default_rng(9), 3)
    d = synth.synth_numpy(n, sps=2, seed=70, bits=bits, starts=[0, 3000, 9000], noise_var=1e-4)
    emit("edge_start_and_p0", synth.mag2(d["iq"]), 2e6, 0.01, [None, [3015, n - 3015], [3016, n - 3016], 4096],
         note="start 0 -> truncated noise window of 15 zeros (snr=inf); call 2 of chunking 1 sees p=0 (snr=nan)")
    # nan / inf / negative samples in a float32 stream
    d = synth.synth_numpy(n, sps=2, n_bursts=20, seed=71)
    x = synth.mag2(d["iq"]).copy()
    rng = np.random.default_rng(72)
    x[rng.integers(0, n, 40)] = np.nan
    x[rng.integers(0, n, 40)] = np.inf
    x[rng.integers(0, n, 40)] = -1.0
    emit("edge_nonfinite", x, 2e6, 0.01, [None, 4096])
    # quantised input (8-bit ADC style): many exact ties and zeros
    d = synth.synth_numpy(1 << 15, sps=2, n_bursts=100, seed=73, noise_var=2e-3)
    q = np.round(d["iq"].view(np.float32) * 16.0) / 16.0
    x = synth.mag2(q.view(np.complex64))
    emit("edge_quantised", x, 2e6, 0.01, [None, 8192], iq=q.view(np.complex64))
    # threshold changed between calls (framer.set_threshold, framer.py:68-69)
    d = synth.synth_numpy(1 << 15, sps=2, n_bursts=60, seed=74)
    emit("sps2_set_threshold", synth.mag2(d["iq"]), 2e6, 0.01, [4096, 2500],
         thresholds=[0.01, 0.05, 0.002, 0.2])
    # ragged / tiny calls
    d = synth.synth_numpy(6000, sps=2, n_bursts=10, seed=75)
    sizes = [1, 2, 3, 500, 16, 15, 17, 240, 241, 1000, 1]
    sizes.append(6000 - sum(sizes))
    emit("sps2_ragged", synth.mag2(d["iq"]), 2e6, 0.01, [sizes, 300])

    # 8. blanking-chain stress: bare preambles (4 pulses, no data) every 16..200 samples, so most
    #    candidates fall inside the 63-symbol blanking interval of an earlier one (framer.py:121,165)
    for sps, seed in ((2, 81), (8, 82)):
        rng = np.random.default_rng(seed)
        n = (1 << 15) * (sps // 2)
        x = rng.exponential(1e-4, n).astype(np.float32)
        pos = 200
        h = sps // 2
        while pos < n - 64 * sps:
            a = np.float32(rng.uniform(0.1, 1.0))
            for hs in (0, 2, 7, 9):
                x[pos + hs * h: pos + (hs + 1) * h] += a
            pos += int(rng.integers(16, 200)) * h
        emit("chain_stress_sps%d" % sps, x, sps * 1e6, 0.01, [None, 8192, 1024])
    # dense overlapping bursts with a 40 dB amplitude spread: the strongest still capture
    d = synth.synth_numpy(1 << 16, sps=2, n_bursts=3277, seed=405, amp_db=(-40.0, 0.0))
    emit("sps2_dense_wide", synth.mag2(d["iq"]), 2e6, 0.01, [None, 8192])


def emit_fec():
    """Vectors for the "Conservative" error correction (decoder.py:304-323,738-780): 112-bit messages with no
    error, 1-bit, 2-contiguous-bit, 2 scattered bits, wrong DF, 3 random flips -- and what the unmodified
    decoder makes of them (parity verdict, bits after correction)."""
    dec = rr.new_decoder(error_corr="Conservative")
    rng = np.random.default_rng(1090)
    base = np.concatenate((np.array([synth.hex_to_bits(h) for h in PUBLIC_DF17]), synth.df17_messages(rng, 594)))
    msgs, kinds = [], []
    for k, m in enumerate(base):
        b = m.copy()
        kind = k % 6
        if kind == 1:
            b[rng.integers(0, 112)] ^= 1
        elif kind == 2:
            i = rng.integers(0, 111)
            b[i] ^= 1
            b[i + 1] ^= 1
        elif kind == 3:
            b[rng.integers(0, 112)] ^= 1
            b[rng.integers(0, 112)] ^= 1
        elif kind == 4:
            b[:5] = (0, 1, 0, 1, 1)
        elif kind == 5:
            b[rng.integers(5, 112, 3)] ^= 1
        msgs.append(b)
        kinds.append(kind)
    msgs = np.array(msgs, np.uint8)
    out = [rr.conservative_fix(b, dec) for b in msgs]
    passed = np.array([o[0] for o in out], np.uint8)
    after = np.array([o[1] for o in out], np.uint8)
    np.savez_compressed(os.path.join(OUT, "fec_vectors.npz"), bits_in=np.packbits(msgs, axis=1),
                        kind=np.array(kinds, np.uint8), passed=passed, bits_out=np.packbits(after, axis=1))
    print("fec_vectors: %d messages, %d pass after correction, %d changed" % (
        len(msgs), int(passed.sum()), int(np.any(after != msgs, axis=1).sum())))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "fec":
        emit_fec()
    else:
        main()
        emit_fec()
