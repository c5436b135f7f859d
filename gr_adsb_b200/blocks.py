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
NUM_TAPS = 2 * MAX_NUM_BITS   # half-symbol samples demod.py:87-95 compares: in0[sob + j*sps/2], j = 0..223


class _BurstCache(object):
    """What the framer's device pass already demodulated, for the demod block behind it.

    The scan kernel slices the 112 PPM bits of every burst it detects (it has the samples in shared memory anyway), so
    when `framer -> demod` are wired as in the stock flowgraph (examples/adsb_rx.py:180-196) the demod's work() needs no
    device call of its own: it looks the burst up here by its absolute tag offset.  An entry is only used if the 224
    half-symbol samples the reference demod would compare (demod.py:87-95) are bit-identical in the demod's own input
    buffer, so anything placed between the two blocks that alters the samples simply turns this into a miss (and the
    demod then runs its own device call).  Shared by all blocks of the process; bounded."""

    def __init__(self, limit=8192):
        self.limit = limit
        self.d = {}

    def put(self, fs, offset, taps, bits):
        if len(self.d) >= self.limit:
            for k in list(self.d)[: self.limit // 2]:        # dicts keep insertion order: drop the oldest half
                self.d.pop(k, None)
        self.d[(fs, offset)] = (taps, bits)

    def take(self, fs, offset):
        return self.d.pop((fs, offset), None)


_CACHE = _BurstCache()


class _DeviceBlock(object):
    """Owns the native context; grows it (keeping the framer state) if GNU Radio hands larger buffers."""

    def _fused(self, in0, fmt, N):
        """adsb_burst_work over this work() call's buffer -> structured hit records (a view of a reused array).
        Everything the call needs is converted once and kept: a work() call of GNU Radio size costs tens of
        microseconds in all, so per-call allocations and conversions would show."""
        ctx = self._ctx_for(len(in0))
        if not in0.flags["C_CONTIGUOUS"]:
            in0 = np.ascontiguousarray(in0)
        cap = N // (63 * self.sps) + 2
        st = getattr(self, "_fs", None)
        if st is None or st[0] < cap:
            import ctypes
            hits = np.zeros(max(cap, 256), _native.HIT_DTYPE)
            nref = ctypes.c_size_t(0)
            st = self._fs = (len(hits), hits, hits.ctypes.data, ctypes.byref(nref), [None, None])
        thr = self.threshold
        tc = st[4]
        if tc[0] != thr:                    # threshold as the reference compares it: rounded to float32 (framer.py:84)
            tc[0], tc[1] = thr, float(np.float32(thr))
        n = ctx.burst_work_raw(in0.ctypes.data, fmt, len(in0), 0, N, tc[1], st[2], st[0], st[3])
        return ctx, st[1][:n]

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

    def __init__(self, fs, threshold, device=0, share_bits=True):
        gr.sync_block.__init__(self, name="ADS-B Framer", in_sig=[np.float32], out_sig=[np.float32])
        self.share_bits = bool(share_bits)     # leave the demodulated bits of detected bursts for the demod block (_BurstCache)
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
        if not self.share_bits:
            idx, snr = self._ctx_for(len(in0)).framer_work(in0, N, self.threshold)
            hits = None
        else:
            # same detection, and the 112 bits of every burst that lies inside this buffer ride along for the demod block
            if in0.dtype != np.float32:
                in0 = in0.astype(np.float32)
            _, hits = self._fused(in0, _native.FMT_F32, N)
            idx, snr = hits["idx"], hits["snr"]
        if len(idx):
            origin = self.nitems_written(0) - (self.N_hist - 1)
            key, src = pmt.to_pmt("burst"), pmt.to_pmt("framer")
            for p, s in zip(idx.tolist(), snr.tolist()):
                self.add_item_tag(0, origin + p, key, pmt.to_pmt(("SOB", float(s))), src)
            if hits is not None:
                h = self.sps // 2
                ok = (hits["flags"] & _native.HIT_DEMOD_VALID) != 0
                if ok.any():
                    bits = _native.unpack_bits(hits[ok])
                    for p, b in zip(idx[ok].tolist(), bits):
                        sob = p + NUM_PREAMBLE_BITS * self.sps
                        _CACHE.put(self.fs, origin + p, in0[sob: sob + NUM_TAPS * h: h].copy(), b)
        out0[:] = in0[self.N_hist - 1:]
        return N


class demod(gr.sync_block, _DeviceBlock):
    """ADS-B demodulator: 112-bit PPM slice of every tagged burst, published as PDUs on "demodulated".

    `batch=True` publishes ONE message per work() call instead of one per burst:
    `(meta{"batch": n, "timestamp": [n floats], "snr": [n floats]}, u8vector(n*112))`; `pdu_unbatch` (a message block)
    or `feed_decoder()` turn it back into the per-burst PDUs the unchanged adsb.decoder consumes (decoder.py:325-335)."""

    def __init__(self, fs, device=0, batch=False):
        gr.sync_block.__init__(self, name="demod", in_sig=[np.float32], out_sig=[np.float32])
        self.fs = fs
        assert self.fs % SYMBOL_RATE == 0, \
            "ADS-B Demodulator is designed to operate on an integer number of samples per symbol, not %f sps" % (self.fs / SYMBOL_RATE)
        self.sps = int(fs // SYMBOL_RATE)
        assert self.sps >= 2 and self.sps % 2 == 0
        self.device = device
        self.batch = bool(batch)
        self.start_timestamp = (datetime.datetime.now(datetime.timezone.utc)
                                - datetime.datetime(1970, 1, 1, tzinfo=datetime.timezone.utc)).total_seconds()
        self.bits = []
        self.straddled_packet = 0
        self.cache_hits = self.cache_misses = 0
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
            sps, h, n = self.sps, self.sps // 2, len(in0)
            out_bits, out_meta, miss = [], [], []
            for t in tags:
                sob = t.offset + NUM_PREAMBLE_BITS * sps - written                  # demod.py:75,79
                if not sob + (MAX_NUM_BITS - 1) * sps + h < n:                      # demod.py:76,80,82
                    self.straddled_packet = 1     # burst not fully inside this buffer: dropped (demod.py:130-133)
                    _CACHE.take(self.fs, t.offset)
                    continue
                ent = _CACHE.take(self.fs, t.offset)
                bits = None
                if ent is not None and sob >= 0 and np.array_equal(in0[sob: sob + NUM_TAPS * h: h], ent[0]):
                    bits = ent[1]
                    self.cache_hits += 1
                else:
                    miss.append((len(out_bits), sob))
                    self.cache_misses += 1
                out_bits.append(bits)
                out_meta.append(t)
            if miss:
                got, valid = self._ctx_for(n).demod_work(in0, np.array([m[1] for m in miss], dtype=np.int64))
                assert valid.all()
                for (slot, _), b in zip(miss, got):
                    out_bits[slot] = b.copy()
            port = pmt.to_pmt("demodulated")
            if self.batch:
                if out_bits:
                    meta = {"batch": len(out_bits),
                            "timestamp": [self.start_timestamp + t.offset / self.fs for t in out_meta],
                            "snr": [pmt.to_python(t.value)[1] for t in out_meta]}
                    self.bits = np.stack(out_bits)
                    self.message_port_pub(port, pmt.cons(pmt.to_pmt(meta), pmt.to_pmt(self.bits.reshape(-1))))
            else:
                for t, b in zip(out_meta, out_bits):
                    snr = pmt.to_python(t.value)[1]
                    self.bits = b
                    meta = pmt.to_pmt({"timestamp": self.start_timestamp + t.offset / self.fs, "snr": snr})
                    self.message_port_pub(port, pmt.cons(meta, pmt.to_pmt(self.bits)))
        out0[:] = in0
        return len(output_items[0])


def split_batch(msg):
    """A batched "demodulated" message -> the list of per-burst PDUs `(meta{"timestamp","snr"}, u8[112])` in the wire
    format adsb.decoder.decode_packet expects (demod.py:104-110, decoder.py:330-335).  Per-burst PDUs pass through."""
    meta = pmt.to_python(pmt.car(msg))
    vec = pmt.to_python(pmt.cdr(msg))
    if not (isinstance(meta, dict) and "batch" in meta):
        return [msg]
    n = int(meta["batch"])
    bits = np.asarray(vec, dtype=np.uint8).reshape(n, MAX_NUM_BITS)
    return [pmt.cons(pmt.to_pmt({"timestamp": meta["timestamp"][i], "snr": meta["snr"][i]}), pmt.to_pmt(bits[i]))
            for i in range(n)]


def feed_decoder(decoder, msg):
    """Hand a (batched or per-burst) "demodulated" message to an UNMODIFIED adsb.decoder: one decode_packet() call per
    burst (decoder.py:325-352)."""
    for pdu in split_batch(msg):
        decoder.decode_packet(pdu)


class pdu_unbatch(gr.basic_block):
    """Message block between a batching demod / burst block and the unchanged adsb.decoder: in port "batch", out port
    "demodulated", one per-burst PDU out for every burst of a batch in."""

    def __init__(self):
        gr.basic_block.__init__(self, name="ADS-B PDU unbatch", in_sig=None, out_sig=None)
        self.message_port_register_in(pmt.to_pmt("batch"))
        self.message_port_register_out(pmt.to_pmt("demodulated"))
        self.set_msg_handler(pmt.to_pmt("batch"), self.handle)

    def handle(self, msg):
        port = pmt.to_pmt("demodulated")
        for pdu in split_batch(msg):
            self.message_port_pub(port, pdu)


class burst(gr.sync_block, _DeviceBlock):
    """framer + demod in ONE block and one device round trip per work() call.

    Emits exactly what the chain `framer -> demod` emits when the scheduler runs the two in lock step:
    the same "burst" tags on the output stream and the same PDUs on "demodulated" (bursts that do not
    end inside the buffer are dropped, demod.py:82,130-133).  With `complex_input=True` the block also
    absorbs `blocks.complex_to_mag_squared` (examples/adsb_rx.py:180) and takes the SDR's complex64
    stream directly; the output stream is then |IQ|^2 of the input, as the removed block produced.
    """

    def __init__(self, fs, threshold, complex_input=False, device=0, batch=False):
        self.batch = bool(batch)     # one "demodulated" message per work() call (see demod / split_batch)
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
        want = np.complex64 if self.complex_input else np.float32
        if in0.dtype != want:
            in0 = in0.astype(want)
        ctx, hits = self._fused(in0, _native.FMT_C64 if self.complex_input else _native.FMT_F32, N)
        if len(hits):
            origin = self.nitems_written(0) - (self.N_hist - 1)
            key, src, port = pmt.to_pmt("burst"), pmt.to_pmt("framer"), pmt.to_pmt("demodulated")
            bits = _native.unpack_bits(hits)
            offs = (origin + hits["idx"]).tolist()
            snrs = hits["snr"].tolist()
            valid = ((hits["flags"] & _native.HIT_DEMOD_VALID) != 0).tolist()
            for off, snr in zip(offs, snrs):
                self.add_item_tag(0, off, key, pmt.to_pmt(("SOB", float(snr))), src)
            if self.batch:
                sel = [k for k in range(len(hits)) if valid[k]]
                if sel:
                    meta = {"batch": len(sel), "timestamp": [self.start_timestamp + offs[k] / self.fs for k in sel],
                            "snr": [float(snrs[k]) for k in sel]}
                    self.message_port_pub(port, pmt.cons(pmt.to_pmt(meta), pmt.to_pmt(bits[sel].reshape(-1))))
            else:
                for k in range(len(hits)):
                    if valid[k]:
                        meta = pmt.to_pmt({"timestamp": self.start_timestamp + offs[k] / self.fs, "snr": float(snrs[k])})
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
