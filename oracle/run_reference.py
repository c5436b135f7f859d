"""Drive the UNMODIFIED reference blocks — TEST INFRASTRUCTURE ONLY.

Imports `/root/reference/python/adsb/{framer,demod,decoder}.py` exactly as they are,
behind the stand-in `gnuradio` / `pmt` / `colorama` modules of `oracle/fake_gr/`, and
plays the role of the GNU Radio scheduler: history, item counters, tag routing
framer -> demod, PDU collection.

Only usable where `/root/reference` exists (the build container).  It is how the
golden vectors under `tests/golden/` were produced (`tests/golden/make_golden.py`)
and how `oracle/adsb_oracle.py` / `oracle/adsb_oracle.c` were pinned; it never runs
on the GPU box and nothing in the product imports it.
"""
import importlib.util
import os
import sys

import numpy as np

REFERENCE_DIR = os.environ.get("ADSB_REFERENCE_DIR", "/root/reference/python/adsb")
_FAKE_GR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "fake_gr")

_mods = {}


def available():
    return os.path.isfile(os.path.join(REFERENCE_DIR, "framer.py"))


def _load(name):
    """Load reference module `name` (framer/demod/decoder) from source, unmodified."""
    if name in _mods:
        return _mods[name]
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_DIR)
    if _FAKE_GR not in sys.path:
        sys.path.insert(0, _FAKE_GR)
    if not hasattr(np, "NaN"):          # decoder.py uses np.NaN (removed in NumPy 2)
        np.NaN = np.nan
    spec = importlib.util.spec_from_file_location(
        "_adsb_reference_" + name, os.path.join(REFERENCE_DIR, name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    _mods[name] = mod
    return mod


def new_framer(fs, threshold):
    return _load("framer").framer(fs, threshold)


def new_demod(fs):
    d = _load("demod").demod(fs)
    d.start_timestamp = 0.0   # wall clock at construction (demod.py:46); pinned so timestamp == offset/fs
    return d


def new_decoder(msg_filter="Extended Squitter Only", error_corr="None"):
    return _load("decoder").decoder(msg_filter, error_corr, "None")


def chunk_sizes(total, chunk):
    """Split `total` items into work() call sizes. `chunk` = int, None (whole) or a list."""
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
    """Feed float32 |IQ|^2 stream `x` through reference framer -> demod in lock step.

    One framer.work() + one demod.work() per chunk (SURVEY.md A.9 policy: the demod only
    sees tags whose offset lies inside the chunk it is handed, so late tags are dropped).
    `thresholds`, if given, is one threshold per call (exercises set_threshold()).

    Returns dict with
      tags : list of (offset:int, snr:float)        every tag the framer attached, call order
      pdus : list of (offset:int, snr:float, bits:uint8[112])   what the demod published
      state: list of (prev_in0, prev_eob_idx) after every framer call
    """
    x = np.ascontiguousarray(x, dtype=np.float32)
    fr = new_framer(fs, threshold)
    dm = new_demod(fs)
    H = fr.N_hist
    hist = np.zeros(H - 1, dtype=np.float32)
    pos = 0
    tags, pdus, state = [], [], []
    port = "demodulated"
    for k, n in enumerate(chunk_sizes(len(x), chunk)):
        if thresholds is not None:
            fr.set_threshold(thresholds[k])
        in0 = np.concatenate((hist, x[pos:pos + n]))
        out0 = np.empty(n, dtype=np.float32)
        fr._nitems_written = pos
        fr._nitems_read = pos
        ntag0 = len(fr._tags_out)
        with np.errstate(all="ignore"):
            r = fr.work([in0], [out0])
        assert r == n
        new_tags = fr._tags_out[ntag0:]
        for t in new_tags:
            assert t.key == "burst" and t.value[0] == "SOB" and t.srcid == "framer"
            tags.append((int(t.offset), float(t.value[1])))
        state.append((float(fr.prev_in0), int(fr.prev_eob_idx)))

        # demod sees the framer's output buffer and the tags attached so far
        dm._nitems_written = pos
        dm._nitems_read = pos
        dm._tags_in = list(fr._tags_out)
        dout = np.empty(n, dtype=np.float32)
        nmsg0 = len(dm._msgs[port])
        with np.errstate(all="ignore"):
            r = dm.work([out0], [dout])
        assert r == n and np.array_equal(dout.view(np.uint32), out0.view(np.uint32))
        for meta, vec in dm._msgs[port][nmsg0:]:
            off = int(round(meta["timestamp"] * fs))
            pdus.append((off, float(meta["snr"]), np.array(vec, dtype=np.uint8)))
        hist = in0[len(in0) - (H - 1):].copy()
        pos += n
    return {"tags": tags, "pdus": pdus, "state": state}


def crc_pass(bits112, dec=None):
    """reference decoder.check_parity() verdict for one 112-bit vector (decoder.py:568-690)."""
    dec = dec or new_decoder()
    dec.reset()
    dec.bits = np.array(bits112, dtype=np.uint8)
    dec.snr = 0.0
    dec.datetime = ""
    import logging
    logging.disable(logging.CRITICAL)
    try:
        dec.decode_header()
        return int(dec.check_parity())
    finally:
        logging.disable(logging.NOTSET)


def compute_crc(bits, dec=None):
    """reference decoder.compute_crc(bits, crc_poly) (decoder.py:693-715) -> 24 bits."""
    dec = dec or new_decoder()
    return np.array(dec.compute_crc(np.array(bits, dtype=int), dec.crc_poly), dtype=np.uint8)


def conservative_fix(bits112, dec=None):
    """The reference decoder's own steps of decode_packet (decoder.py:337-344) with error_corr="Conservative":
    returns (parity_passed, bits_after)."""
    dec = dec or new_decoder(error_corr="Conservative")
    import logging
    logging.disable(logging.CRITICAL)
    try:
        dec.reset()
        dec.bits = np.array(bits112, dtype=np.uint8).copy()
        dec.snr = 0.0
        dec.datetime = ""
        dec.decode_header()
        passed = dec.check_parity()
        if passed == 0:
            passed = dec.correct_errors()
        return int(bool(passed)), np.array(dec.bits, dtype=np.uint8)
    finally:
        logging.disable(logging.NOTSET)
