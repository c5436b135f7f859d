"""`adsb.framer` and `adsb.demod` with gr-adsb's GNU Radio block surface, computed on a B200.

Same constructor signatures, attribute names, history, tag key/value/srcid, message port and PDU
layout as the reference blocks (python/adsb/framer.py:33-182, python/adsb/demod.py:31-136; GRC
make templates grc/adsb_framer.block.yml:7-9, grc/adsb_demod.block.yml:7), so they drop into the
stock adsb_rx flowgraph ahead of the unchanged Python adsb.decoder.  All signal processing happens
in the CUDA library behind `include/adsb_b200.h`; there is no NumPy fallback.
"""
import datetime

import numpy as np

from . import _native
from ._gr_compat import gr, pmt

SYMBOL_RATE = 1e6          # symbols/second                       (framer.py:27, demod.py:28)
NUM_PREAMBLE_BITS = 8      #                                       (framer.py:28)
MIN_NUM_BITS = 56          #                                       (framer.py:29)
NUM_PREAMBLE_PULSES = NUM_PREAMBLE_BITS * 2                      # (framer.py:30)
NUM_NOISE_SAMPLES = 100    #                                       (framer.py:31)
MAX_NUM_BITS = 112         #                                       (demod.py:29)

_DEFAULT_MAX_ITEMS = 1 << 20


class _DeviceBlock(object):
    """Owns the native context; grows it (keeping the framer state) if GNU Radio hands larger buffers."""

    def _ctx_for(self, n_items):
        ctx = getattr(self, "_ctx", None)
        if ctx is None or n_items > ctx.max_items:
            state = ctx.get_state() if ctx is not None else None
            cap = max(_DEFAULT_MAX_ITEMS, 2 * int(n_items))
            new = _native.Context(self.sps, cap, device=getattr(self, "device", 0))
            if state is not None:
                new.set_state(*state)
                ctx.close()
            self._ctx = new
        return self._ctx

    def close(self):
        """Release the device buffers now (otherwise when the block is garbage collected)."""
        ctx = getattr(self, "_ctx", None)
        if ctx is not None:
            ctx.close()
            self._ctx = None


class framer(gr.sync_block, _DeviceBlock):
    """ADS-B Framer: burst detection, emits "burst" stream tags ("SOB", snr)."""

    def __init__(self, fs, threshold, device=0):
        gr.sync_block.__init__(self, name="ADS-B Framer", in_sig=[np.float32], out_sig=[np.float32])
        self.fs = fs
        assert self.fs % SYMBOL_RATE == 0, \
            "ADS-B Framer is designed to operate on an integer number of samples per symbol, not %f sps" % (self.fs / SYMBOL_RATE)
        self.sps = int(fs // SYMBOL_RATE)
        assert self.sps >= 2 and self.sps % 2 == 0, "the half-symbol taps need an even number of samples per symbol"
        self.threshold = threshold
        self.device = device
        self.preamble_pulses = [1, 0, 1, 0, 0, 0, 0, 1, 0, 1, 0, 0, 0, 0, 0, 0]
        self.N_hist = NUM_PREAMBLE_BITS * self.sps
        self.set_history(self.N_hist)
        self.set_tag_propagation_policy(gr.TPP_ONE_TO_ONE)
        self._ctx_for(_DEFAULT_MAX_ITEMS)

    # the reference keeps these as plain attributes (framer.py:54,57); here they live in the context
    @property
    def prev_in0(self):
        return self._ctx.get_state()[0]

    @property
    def prev_eob_idx(self):
        return self._ctx.get_state()[1]

    def set_threshold(self, threshold):
        self.threshold = threshold

    def work(self, input_items, output_items):
        in0 = input_items[0]
        out0 = output_items[0]
        N = len(out0)
        ctx = self._ctx_for(len(in0))
        idx, snr = ctx.framer_work(in0, N, self.threshold)
        if len(idx):
            origin = self.nitems_written(0) - (self.N_hist - 1)
            key, src = pmt.to_pmt("burst"), pmt.to_pmt("framer")
            for p, s in zip(idx.tolist(), snr.tolist()):
                self.add_item_tag(0, origin + p, key, pmt.to_pmt(("SOB", float(s))), src)
        out0[:] = in0[self.N_hist - 1:]
        return N


class demod(gr.sync_block, _DeviceBlock):
    """ADS-B demodulator: 112-bit PPM slice of every tagged burst, published as PDUs on "demodulated"."""

    def __init__(self, fs, device=0):
        gr.sync_block.__init__(self, name="demod", in_sig=[np.float32], out_sig=[np.float32])
        self.fs = fs
        assert self.fs % SYMBOL_RATE == 0, \
            "ADS-B Demodulator is designed to operate on an integer number of samples per symbol, not %f sps" % (self.fs / SYMBOL_RATE)
        self.sps = int(fs // SYMBOL_RATE)
        assert self.sps >= 2 and self.sps % 2 == 0
        self.device = device
        self.start_timestamp = (datetime.datetime.now(datetime.timezone.utc)
                                - datetime.datetime(1970, 1, 1, tzinfo=datetime.timezone.utc)).total_seconds()
        self.bits = []
        self.straddled_packet = 0
        self.set_tag_propagation_policy(gr.TPP_ONE_TO_ONE)
        self.message_port_register_out(pmt.to_pmt("demodulated"))
        self._ctx_for(_DEFAULT_MAX_ITEMS)

    def work(self, input_items, output_items):
        in0 = input_items[0]
        out0 = output_items[0]
        if self.straddled_packet == 1:
            self.straddled_packet = 0
        first = self.nitems_read(0)
        tags = self.get_tags_in_range(0, first, first + len(in0), pmt.to_pmt("burst"))
        if tags:
            written = self.nitems_written(0)
            sob = np.array([t.offset + NUM_PREAMBLE_BITS * self.sps - written for t in tags], dtype=np.int64)
            bits, valid = self._ctx_for(len(in0)).demod_work(in0, sob)
            port = pmt.to_pmt("demodulated")
            for k, t in enumerate(tags):
                if not valid[k]:
                    self.straddled_packet = 1     # burst not fully inside this buffer: dropped (demod.py:130-133)
                    continue
                snr = pmt.to_python(t.value)[1]
                self.bits = bits[k].copy()
                meta = pmt.to_pmt({"timestamp": self.start_timestamp + t.offset / self.fs, "snr": snr})
                self.message_port_pub(port, pmt.cons(meta, pmt.to_pmt(self.bits)))
        out0[:] = in0
        return len(output_items[0])


class burst(gr.sync_block, _DeviceBlock):
    """framer + demod in ONE block and one device round trip per work() call.

    Emits exactly what the chain `framer -> demod` emits when the scheduler runs the two in lock step:
    the same "burst" tags on the output stream and the same PDUs on "demodulated" (bursts that do not
    end inside the buffer are dropped, demod.py:82,130-133).  With `complex_input=True` the block also
    absorbs `blocks.complex_to_mag_squared` (examples/adsb_rx.py:180) and takes the SDR's complex64
    stream directly; the output stream is then |IQ|^2 of the input, as the removed block produced.
    """

    def __init__(self, fs, threshold, complex_input=False, device=0):
        in_t = np.complex64 if complex_input else np.float32
        gr.sync_block.__init__(self, name="ADS-B Burst (framer+demod)", in_sig=[in_t], out_sig=[np.float32])
        self.fs = fs
        assert self.fs % SYMBOL_RATE == 0, "integer number of samples per symbol required"
        self.sps = int(fs // SYMBOL_RATE)
        assert self.sps >= 2 and self.sps % 2 == 0
        self.threshold = threshold
        self.complex_input = bool(complex_input)
        self.device = device
        self.N_hist = NUM_PREAMBLE_BITS * self.sps
        self.set_history(self.N_hist)
        self.set_tag_propagation_policy(gr.TPP_ONE_TO_ONE)
        self.message_port_register_out(pmt.to_pmt("demodulated"))
        self.start_timestamp = (datetime.datetime.now(datetime.timezone.utc)
                                - datetime.datetime(1970, 1, 1, tzinfo=datetime.timezone.utc)).total_seconds()
        self._ctx_for(_DEFAULT_MAX_ITEMS)

    @property
    def prev_in0(self):
        return self._ctx.get_state()[0]

    @property
    def prev_eob_idx(self):
        return self._ctx.get_state()[1]

    def set_threshold(self, threshold):
        self.threshold = threshold

    def work(self, input_items, output_items):
        in0 = input_items[0]
        out0 = output_items[0]
        N = len(out0)
        ctx = self._ctx_for(len(in0))
        hits = ctx.burst_work(in0, n_out=N, lead=0, threshold=self.threshold)
        if len(hits):
            origin = self.nitems_written(0) - (self.N_hist - 1)
            key, src, port = pmt.to_pmt("burst"), pmt.to_pmt("framer"), pmt.to_pmt("demodulated")
            bits = _native.unpack_bits(hits)
            for k in range(len(hits)):
                off = origin + int(hits["idx"][k])
                snr = float(hits["snr"][k])
                self.add_item_tag(0, off, key, pmt.to_pmt(("SOB", snr)), src)
                if hits["flags"][k] & _native.HIT_DEMOD_VALID:
                    meta = pmt.to_pmt({"timestamp": self.start_timestamp + off / self.fs, "snr": snr})
                    self.message_port_pub(port, pmt.cons(meta, pmt.to_pmt(bits[k].copy())))
        if self.complex_input:
            # the |IQ|^2 stream the removed complex_to_mag_squared block produced, from the device copy of in0
            if out0.dtype == np.float32 and out0.flags["C_CONTIGUOUS"]:
                ctx.burst_mag_out(out0)
            else:
                out0[:] = ctx.burst_mag_out(np.empty(N, np.float32))
        else:
            out0[:] = in0[self.N_hist - 1:]           # framer.py:181
        return N
