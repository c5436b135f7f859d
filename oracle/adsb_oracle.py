"""CPU oracle for the ADS-B burst-detect + PPM-demod path — TEST INFRASTRUCTURE ONLY.

A NumPy restatement of the reference algorithm (gr-adsb `python/adsb/framer.py:72-182`
and `python/adsb/demod.py:57-136`), without GNU Radio.  It exists so the CUDA path can be
checked on the GPU box, where `/root/reference` does not exist.  Only `tests/`,
`__graft_entry__.smoke()` and the CPU-baseline legs of `bench.py` may import it; the product
package never does.

PINNING: the reference ships no golden vectors and its own tests are empty stubs
(qa_framer.py:34-37, qa_demod.py:34-37), so parity is pinned by executing the unmodified
reference here (`oracle/run_reference.py`) on seeded inputs and committing its outputs
under `tests/golden/` (`tests/golden/make_golden.py`).  `tests/test_oracle_golden.py` checks
this restatement against every one of those fixtures; `tests/test_oracle_vs_reference.py`
additionally fuzzes it against the live reference wherever `/root/reference` is present.

The structure deliberately keeps the reference's cost profile (vectorised threshold/edge
search, then an interpreter loop over pulses, `np.median` per detection), because this port
is also what `bench.py` times as the CPU baseline ("kind": "port").
"""
import numpy as np

SYMBOL_RATE = 1e6            # framer.py:27, demod.py:28
PREAMBLE_SYMBOLS = 8         # framer.py:28
SHORT_MSG_BITS = 56          # framer.py:29
NOISE_WINDOW = 100           # framer.py:31
MSG_BITS = 112               # demod.py:29
TEMPLATE = np.array([1, 0, 1, 0, 0, 0, 0, 1, 0, 1, 0, 0, 0, 0, 0, 0])   # framer.py:50


def sps_of(fs):
    """framer.py:44-45 / demod.py:42-43."""
    assert fs % SYMBOL_RATE == 0, "integer samples per symbol required"
    return int(fs // SYMBOL_RATE)


class FramerState(object):
    """The two scalars the reference framer carries between work() calls (framer.py:54,57)."""

    def __init__(self):
        self.prev_in0 = 0
        self.prev_eob_idx = -1


def framer_work(st, in0, n_out, sps, threshold):
    """One framer.work() call (framer.py:72-182) on window `in0` (len n_out + 8*sps - 1).

    Returns [(p, snr), ...] in window coordinates, ascending p; the GR tag offset is
    nitems_written - (8*sps - 1) + p (framer.py:170).  `st` is updated in place.
    """
    N = int(n_out)
    half = sps // 2
    # framer.py:83-87 -- threshold bitmap over [prev sample] + first N samples
    ext = np.empty(N + 1, dtype=np.float32)
    ext[0] = st.prev_in0
    ext[1:] = in0[:N]
    above = (ext >= threshold).astype(np.int64)
    st.prev_in0 = in0[N - 1]
    # framer.py:91-93 -- edges
    step = np.diff(above)
    rises = np.flatnonzero(step == 1)
    falls = np.flatnonzero(step == -1)
    hits = []
    if rises.size and falls.size:
        # framer.py:98-110 -- drop a leading orphan fall / a trailing orphan rise
        if falls[0] < rises[0]:
            falls = falls[1:]
        if rises.size == falls.size + 1:
            rises = rises[:-1]
        # framer.py:113 -- truncated integer mean of the two edges
        centres = (rises + falls) // 2 if rises.size == falls.size else np.empty(0, dtype=np.int64)
        span = PREAMBLE_SYMBOLS * sps
        for p in centres:
            p = int(p)
            if p <= st.prev_eob_idx:          # framer.py:121 -- still inside the previous burst
                continue
            st.prev_eob_idx = -1              # framer.py:123
            taps = in0[p:p + span:half]       # framer.py:137
            marks = (taps > in0[p] / 2).astype(np.int64)          # framer.py:140-141
            if np.count_nonzero(marks == TEMPLATE) != TEMPLATE.size:   # framer.py:144-147
                continue
            lo = 0 if p < NOISE_WINDOW else p - NOISE_WINDOW      # framer.py:156-159
            snr = 10.0 * np.log10(in0[p] / np.median(in0[lo:p])) + 1.6
            st.prev_eob_idx = p + (PREAMBLE_SYMBOLS + SHORT_MSG_BITS - 1) * sps   # framer.py:165
            hits.append((p, snr))
        if st.prev_eob_idx >= N:              # framer.py:177-179
            st.prev_eob_idx -= N
    return hits


def demod_slice(in0, sob_idx, sps):
    """demod.py:87-95 -- 112 single-sample PPM decisions starting at `sob_idx`."""
    half = sps // 2
    k = sob_idx + sps * np.arange(MSG_BITS)
    return (in0[k] > in0[k + half]).astype(np.uint8)


def demod_work(in0, n_read, tags, sps):
    """One demod.work() call (demod.py:57-136).  `tags` = [(offset, snr)] (absolute offsets,
    any order); only those inside [n_read, n_read+len(in0)) are considered (demod.py:67).
    Returns [(offset, snr, bits)] for bursts fully inside the buffer (demod.py:82); the rest are
    dropped as straddlers (demod.py:130-133)."""
    out = []
    n = len(in0)
    for off, snr in sorted((t for t in tags if n_read <= t[0] < n_read + n), key=lambda t: t[0]):
        sob = off + PREAMBLE_SYMBOLS * sps - n_read                       # demod.py:75,79
        eob = off + (PREAMBLE_SYMBOLS + MSG_BITS - 1) * sps + sps / 2 - n_read   # demod.py:76,80
        if eob < n:                                                        # demod.py:82
            out.append((int(off), snr, demod_slice(in0, int(sob), sps)))
    return out


def chunk_sizes(total, chunk):
    if chunk is None:
        return [total] if total > 0 else []
    if isinstance(chunk, (list, tuple, np.ndarray)):
        sizes = [int(c) for c in chunk]
        assert sum(sizes) == total
        return sizes
    sizes = [chunk] * (total // chunk)
    if total % chunk:
        sizes.append(total % chunk)
    return sizes


def run_stream(x, fs, threshold, chunk=None, thresholds=None):
    """Lock-step framer -> demod over consecutive work() calls, GR-style history of 8*sps-1
    zeros at stream start.  Same return layout as `oracle/run_reference.run_stream`."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    sps = sps_of(fs)
    H = PREAMBLE_SYMBOLS * sps
    st = FramerState()
    hist = np.zeros(H - 1, dtype=np.float32)
    pos = 0
    tags, pdus, state = [], [], []
    for k, n in enumerate(chunk_sizes(len(x), chunk)):
        thr = threshold if thresholds is None else thresholds[k]
        in0 = np.concatenate((hist, x[pos:pos + n]))
        with np.errstate(all="ignore"):
            hits = framer_work(st, in0, n, sps, thr)
        new = [(pos - (H - 1) + p, float(s)) for p, s in hits]
        tags.extend(new)
        state.append((float(st.prev_in0), int(st.prev_eob_idx)))
        pdus.extend((o, float(s), b) for o, s, b in demod_work(in0[H - 1:], pos, new, sps))
        hist = in0[len(in0) - (H - 1):].copy()
        pos += n
    return {"tags": tags, "pdus": pdus, "state": state}


def mag2(iq):
    """complex64 -> float32 |IQ|^2, multiply-then-add in float32 (gr-blocks
    complex_to_mag_squared, instantiated adsb_rx.py:180)."""
    v = np.ascontiguousarray(iq, dtype=np.complex64).view(np.float32).reshape(-1, 2)
    return (v[:, 0] * v[:, 0] + v[:, 1] * v[:, 1]).astype(np.float32)


# --- CRC-24 (decoder.py:269, 693-715, 667-685) --------------------------
CRC_POLY_BITS = np.array([1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 0, 1, 0, 0, 0, 0, 0, 0, 1, 0, 0, 1])


def compute_crc(data_bits):
    """Long division of data*x^24 by the Mode S generator (decoder.py:693-715) -> 24 bits."""
    d = np.concatenate((np.asarray(data_bits, dtype=np.int64), np.zeros(24, dtype=np.int64)))
    for i in range(len(data_bits)):
        if d[i]:
            d[i:i + 25] ^= CRC_POLY_BITS
    return d[len(data_bits):].astype(np.uint8)


def crc_pass(bits112):
    """decoder.check_parity() for the "Extended Squitter Only" filter (decoder.py:667-685):
    DF in {17,18,19} and PI == CRC(bits[0:88])."""
    b = np.asarray(bits112, dtype=np.uint8)
    df = int("".join(str(int(v)) for v in b[:5]), 2)
    if df not in (17, 18, 19):
        return 0
    return int(np.array_equal(compute_crc(b[:88]), b[88:112]))


# --- Conservative error correction (decoder.py:304-323, 716-763, 765-780) -------------
def residual25(bits):
    """decoder.compute_crc_2 (decoder.py:716-736): long division that stops one step early and returns the
    25-bit window bits[len-25:len] as an integer -- the quantity the reference's syndrome table is keyed by."""
    d = np.asarray(bits, dtype=np.int64).copy()
    n = len(d) - 25
    for i in range(n):
        if d[i]:
            d[i:i + 25] ^= CRC_POLY_BITS
    v = 0
    for b in d[n:n + 25]:
        v = (v << 1) | int(b)
    return v


def fec_table(payload_length=112):
    """decoder.compute_crc_syndromes_for_contiguous_bursts (decoder.py:304-323) for bursts of 1 and 2 bits:
    {residual25(error pattern): [bit positions]}."""
    lut = {}
    for burst in (1, 2):
        for i in range(payload_length - burst + 1):
            z = np.zeros(payload_length, dtype=np.int64)
            z[i:i + burst] = 1
            key = residual25(z)
            assert key not in lut, "FEC syndrome collision"         # decoder.py:319-321
            lut[key] = list(range(i, i + burst))
    return lut


_FEC112 = None


def conservative_fix(bits112):
    """What decoder.decode_packet does up to the parity verdict with msg_filter="Extended Squitter Only",
    error_corr="Conservative" (decoder.py:341-344): returns (parity_passed, fixed, bits_after)."""
    global _FEC112
    b = np.array(bits112, dtype=np.uint8).copy()
    if crc_pass(b):
        return 1, 0, b
    df = int("".join(str(int(v)) for v in b[:5]), 2)
    if df not in (17, 18, 19):            # check_parity never set payload_length (decoder.py:667-670,739-740)
        return 0, 0, b
    if _FEC112 is None:
        _FEC112 = fec_table(112)
    pos = _FEC112.get(residual25(b))
    if pos is None:
        return 0, 0, b
    for p in pos:
        b[p] ^= 1
    ok = int(residual25(b) == 0)          # decoder.py:756-758
    return ok, ok, b
