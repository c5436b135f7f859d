"""A lock-step stand-in for the GNU Radio scheduler, for machines without GNU Radio (examples, bench.py, tests).

Feeds a stream through `[framer, demod]` or `[burst]` the way a flowgraph would when both blocks are scheduled once
per buffer: the first block gets `history() - 1` items of look-back prepended (zeros at stream start), item counters
advance by the items produced, the framer's tags reach the demod of the same call, PDUs are collected from the
"demodulated" port.  Only works with the stand-in `gr` of `_gr_compat` (with GNU Radio installed, build a real
`gr.top_block` -- the blocks are ordinary `gr.sync_block`s there).
"""
import numpy as np

from ._gr_compat import HAVE_GNURADIO

PORT = "demodulated"


def run_chain(blocks, x, chunk, collect=False):
    """x: float32 |IQ|^2 (or complex64 for `burst(complex_input=True)`); chunk: items per work() call (int or list).
    -> dict(n_tags, n_pdus[, tags, pdus])"""
    if HAVE_GNURADIO:                                  # pragma: no cover
        raise RuntimeError("flowgraph.run_chain drives the GNU-Radio-less stand-in blocks; use gr.top_block instead")
    first, last = blocks[0], blocks[-1]
    x = np.ascontiguousarray(x)
    H = first.history()
    hist = np.zeros(H - 1, dtype=x.dtype)
    n_total = len(x)
    sizes = chunk if isinstance(chunk, (list, tuple)) else None
    pos = k = n_tags = 0
    tags, pdus = [], []
    if hasattr(last, "start_timestamp"):
        last.start_timestamp = 0.0
    while pos < n_total:
        n = sizes[k] if sizes is not None else min(int(chunk), n_total - pos)
        in0 = np.concatenate((hist, x[pos:pos + n]))
        out0 = np.empty(n, dtype=np.float32)
        first._nitems_written = first._nitems_read = pos
        first._tags_out = []
        assert first.work([in0], [out0]) == n
        n_tags += len(first._tags_out)
        if collect:
            tags.extend((int(t.offset), t.value[1]) for t in first._tags_out)
        if last is not first:
            last._nitems_written = last._nitems_read = pos
            last._tags_in = first._tags_out
            dout = np.empty(n, dtype=np.float32)
            assert last.work([out0], [dout]) == n
        hist = in0[len(in0) - (H - 1):]
        pos += n
        k += 1
    msgs = last._msgs.get(PORT, [])
    if collect:
        for m in msgs:
            for meta, vec in _pdus_of(m):
                pdus.append((int(round(meta["timestamp"] * first.fs)), float(meta["snr"]), np.array(vec)))
    n_pdus = sum(len(_pdus_of(m)) for m in msgs)
    out = {"n_tags": n_tags, "n_pdus": n_pdus}
    if collect:
        out["tags"], out["pdus"] = tags, pdus
    return out


def _pdus_of(msg):
    """A published message is one PDU `(meta, u8[112])` or a batch `(meta{'batch': n, ...}, u8[n,112])`."""
    meta, vec = msg
    if isinstance(meta, dict) and "batch" in meta:
        bits = np.asarray(vec, dtype=np.uint8).reshape(meta["batch"], 112)
        return [({"timestamp": meta["timestamp"][i], "snr": meta["snr"][i]}, bits[i]) for i in range(meta["batch"])]
    return [(meta, vec)]
