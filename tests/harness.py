"""Plays GNU Radio scheduler for a (framer, demod) block pair built on gr_adsb_b200._gr_compat:
history, item counters, tag routing, PDU collection -- lock step, one framer.work() + one
demod.work() per chunk.  Mirrors oracle/run_reference.run_stream."""
import numpy as np


def run_blocks(fr, dm, x, sizes, thresholds=None):
    x = np.ascontiguousarray(x, dtype=np.float32)
    H = fr.N_hist
    hist = np.zeros(H - 1, dtype=np.float32)
    pos = 0
    tags, pdus, state = [], [], []
    dm.start_timestamp = 0.0
    port = "demodulated"
    for k, n in enumerate(sizes):
        if thresholds is not None:
            fr.set_threshold(thresholds[k])
        in0 = np.concatenate((hist, x[pos:pos + n]))
        out0 = np.empty(n, dtype=np.float32)
        fr._nitems_written = pos
        fr._nitems_read = pos
        nt0 = len(fr._tags_out)
        assert fr.work([in0], [out0]) == n
        for t in fr._tags_out[nt0:]:
            assert t.key == "burst" and t.value[0] == "SOB" and t.srcid == "framer"
            assert type(t.value[1]) is float
            tags.append((int(t.offset), t.value[1]))
        state.append(fr._ctx.get_state())
        dm._nitems_written = pos
        dm._nitems_read = pos
        dm._tags_in = list(fr._tags_out)
        dout = np.empty(n, dtype=np.float32)
        nm0 = len(dm._msgs[port])
        assert dm.work([out0], [dout]) == n
        assert np.array_equal(dout.view(np.uint32), x[pos:pos + n].view(np.uint32))
        for meta, vec in dm._msgs[port][nm0:]:
            assert vec.dtype == np.uint8 and vec.shape == (112,)
            pdus.append((int(round(meta["timestamp"] * fr.fs)), float(meta["snr"]), np.array(vec)))
        hist = in0[len(in0) - (H - 1):].copy()
        pos += n
    return {"tags": tags, "pdus": pdus, "state": state}
