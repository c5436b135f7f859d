"""ctypes binding of the C ABI in `include/adsb_b200.h` (built into `csrc/libadsb_b200.so`).

There is no CPU fallback: if the shared library or a B200 is missing, creating a context raises.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
BUILT_LIB = os.path.join(CSRC, "libadsb_b200.so")                 # what build() produces
LIB_PATH = os.environ.get("ADSB_B200_LIB", BUILT_LIB)
HEADER = os.path.join(os.path.dirname(_HERE), "include", "adsb_b200.h")

ADSB_OK = 0
E_CAPACITY = -4
HIT_DEMOD_VALID = 0x1
HIT_CRC_OK = 0x2
HIT_LATE_TAG = 0x4
HIT_FEC_FIXED = 0x8
FEC_NONE, FEC_CONSERVATIVE = 0, 1
FMT_F32, FMT_C64, FMT_SC16, FMT_CU8 = 0, 1, 2, 3
CARRY_USED_ON_DEVICE = -(1 << 63)
E_PEER = -8
SHARD_CTRL_BYTES, IPC_HANDLE_BYTES = 8192, 64
SHARD_PH_SIGNAL, SHARD_PH_SCAN, SHARD_PH_EXPORT, SHARD_PH_PLAN, SHARD_PH_CHECK, SHARD_PH_ALL = 1, 2, 4, 8, 16, 15

HIT_DTYPE = np.dtype([("idx", "<i8"), ("snr", "<f4"), ("flags", "<u4"), ("bits", "u1", (16,))], align=True)
assert HIT_DTYPE.itemsize == 32


class Counters(C.Structure):
    _fields_ = [("n_samples", C.c_uint64), ("n_candidates", C.c_uint64), ("n_hits", C.c_uint64),
                ("n_demod", C.c_uint64), ("n_crc_ok", C.c_uint64), ("last_pulse", C.c_int64),
                ("last_accept", C.c_int64), ("first_candidate", C.c_int64),
                ("carry_unstable", C.c_uint32), ("n_fec_fixed", C.c_uint32),
                ("any_rise", C.c_uint32), ("any_fall", C.c_uint32)]

    def as_dict(self):
        return {k: int(getattr(self, k)) for k, _ in self._fields_}


NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-pthread", "-shared"]


def build(force=False, verbose=False):
    """Compile csrc/*.cu for sm_100a into csrc/libadsb_b200.so (in-tree, so it travels with the repo)."""
    srcs = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cu", ".cuh"))]
    srcs.append(HEADER)
    if not force and os.path.isfile(BUILT_LIB) and all(os.path.getmtime(BUILT_LIB) >= os.path.getmtime(s) for s in srcs):
        return BUILT_LIB
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + \
          ["-o", BUILT_LIB, os.path.join(CSRC, "adsb_api.cu")]
    subprocess.check_call(cmd)
    return BUILT_LIB


_lib = None

EXPORTS = ["adsb_strerror", "adsb_abi_version", "adsb_device_count", "adsb_create", "adsb_destroy",
           "adsb_last_error", "adsb_reset_state", "adsb_get_state", "adsb_set_state", "adsb_get_counters",
           "adsb_framer_work", "adsb_demod_work", "adsb_burst_work", "adsb_burst_work_dev", "adsb_burst_finish",
           "adsb_profile_enable", "adsb_profile_last", "adsb_burst_reresolve_dev",
           "adsb_burst_export_carry_dev", "adsb_burst_carry_round_dev", "adsb_set_error_correction",
           "adsb_burst_mag_out", "adsb_shard_alloc", "adsb_shard_attach", "adsb_shard_attach_ptrs", "adsb_shard_detach",
           "adsb_shard_write", "adsb_shard_step", "adsb_latency_stats", "adsb_set_latency_path", "adsb_latency_trace", "adsb_stream_open", "adsb_stream_push",
           "adsb_stream_acquire", "adsb_stream_commit", "adsb_stream_pop", "adsb_stream_flush", "adsb_stream_close",
           "adsb_stream_stats"]


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if "ADSB_B200_LIB" not in os.environ:
        try:
            build()           # no-op when the in-tree library is newer than its sources
        except FileNotFoundError as e:
            # no nvcc on this machine (a GPU box that received the prebuilt library with the snapshot): use what is there
            if not os.path.isfile(LIB_PATH):
                raise RuntimeError("%s is not built and there is no nvcc here (%s): run "
                                   "`python -c 'import __graft_entry__ as g; g.build()'` where the CUDA toolkit is" % (LIB_PATH, e))
        # a compile error (CalledProcessError) propagates: running on a stale library after an edit would test old kernels
    if not os.path.isfile(LIB_PATH):
        raise RuntimeError("%s not built: run `python -c 'import __graft_entry__ as g; g.build()'`" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    if hasattr(L, "adsb_host_emulation_marker") and os.environ.get("ADSB_EMU") != "1":
        # tests/emu builds the CUDA sources for the host to check kernel logic without a GPU; that library is test
        # infrastructure and must never serve as the product path
        raise RuntimeError("%s is the host-emulation TEST library (tests/emu); it only loads in ADSB_EMU=1 test runs"
                           % LIB_PATH)
    vp, sz, i64, f32, i32 = C.c_void_p, C.c_size_t, C.c_int64, C.c_float, C.c_int
    L.adsb_strerror.restype = C.c_char_p
    L.adsb_strerror.argtypes = [i32]
    L.adsb_abi_version.restype = i32
    L.adsb_device_count.restype = i32
    L.adsb_create.restype = i32
    L.adsb_create.argtypes = [C.POINTER(vp), i32, i32, sz]
    L.adsb_destroy.restype = None
    L.adsb_destroy.argtypes = [vp]
    L.adsb_last_error.restype = C.c_char_p
    L.adsb_last_error.argtypes = [vp]
    L.adsb_reset_state.restype = i32
    L.adsb_reset_state.argtypes = [vp]
    L.adsb_get_state.restype = i32
    L.adsb_get_state.argtypes = [vp, C.POINTER(f32), C.POINTER(i64)]
    L.adsb_set_state.restype = i32
    L.adsb_set_state.argtypes = [vp, f32, i64]
    L.adsb_set_error_correction.restype = i32
    L.adsb_set_error_correction.argtypes = [vp, i32]
    L.adsb_get_counters.restype = i32
    L.adsb_get_counters.argtypes = [vp, C.POINTER(Counters)]
    L.adsb_framer_work.restype = i32
    L.adsb_framer_work.argtypes = [vp, vp, sz, sz, f32, vp, vp, sz, C.POINTER(sz)]
    L.adsb_demod_work.restype = i32
    L.adsb_demod_work.argtypes = [vp, vp, sz, vp, sz, vp, vp]
    L.adsb_burst_work.restype = i32
    L.adsb_burst_work.argtypes = [vp, vp, i32, sz, sz, sz, f32, i64, vp, sz, C.POINTER(sz)]
    L.adsb_burst_mag_out.restype = i32
    L.adsb_burst_mag_out.argtypes = [vp, vp, sz]
    L.adsb_burst_work_dev.restype = i32
    L.adsb_burst_work_dev.argtypes = [vp, vp, i32, sz, sz, sz, sz, f32, i64, i64, vp, C.POINTER(vp)]
    L.adsb_burst_finish.restype = i32
    L.adsb_burst_finish.argtypes = [vp, C.POINTER(Counters), vp, sz, C.POINTER(sz)]
    L.adsb_burst_reresolve_dev.restype = i32
    L.adsb_burst_reresolve_dev.argtypes = [vp, i64, vp]
    L.adsb_burst_export_carry_dev.restype = i32
    L.adsb_burst_export_carry_dev.argtypes = [vp, i64, vp, vp]
    L.adsb_burst_carry_round_dev.restype = i32
    L.adsb_burst_carry_round_dev.argtypes = [vp, vp, i32, i32, i32, vp]
    L.adsb_shard_alloc.restype = i32
    L.adsb_shard_alloc.argtypes = [vp, sz, C.POINTER(vp), vp]
    L.adsb_shard_attach.restype = i32
    L.adsb_shard_attach.argtypes = [vp, i32, i32, vp]
    L.adsb_shard_attach_ptrs.restype = i32
    L.adsb_shard_attach_ptrs.argtypes = [vp, i32, i32, C.POINTER(vp)]
    L.adsb_shard_detach.restype = i32
    L.adsb_shard_detach.argtypes = [vp]
    L.adsb_shard_write.restype = i32
    L.adsb_shard_write.argtypes = [vp, sz, vp, sz, vp]
    L.adsb_shard_step.restype = i32
    L.adsb_shard_step.argtypes = [vp, i32, sz, f32, i32, i32, vp]
    L.adsb_latency_stats.restype = i32
    L.adsb_latency_stats.argtypes = [vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    L.adsb_stream_open.restype = i32
    L.adsb_stream_open.argtypes = [vp, i32, f32, sz]
    L.adsb_stream_push.restype = i32
    L.adsb_stream_push.argtypes = [vp, vp, sz]
    L.adsb_stream_acquire.restype = i32
    L.adsb_stream_acquire.argtypes = [vp, C.POINTER(vp), C.POINTER(sz)]
    L.adsb_stream_commit.restype = i32
    L.adsb_stream_commit.argtypes = [vp, sz]
    L.adsb_stream_pop.restype = i32
    L.adsb_stream_pop.argtypes = [vp, vp, sz, C.POINTER(sz)]
    L.adsb_stream_flush.restype = i32
    L.adsb_stream_flush.argtypes = [vp]
    L.adsb_stream_close.restype = i32
    L.adsb_stream_close.argtypes = [vp]
    L.adsb_stream_stats.restype = i32
    L.adsb_stream_stats.argtypes = [vp] + [C.POINTER(C.c_uint64)] * 4
    L.adsb_latency_trace.restype = i32
    L.adsb_latency_trace.argtypes = [vp, C.POINTER(C.c_uint64), C.POINTER(C.c_double)]
    L.adsb_set_latency_path.restype = i32
    L.adsb_set_latency_path.argtypes = [vp, i32]
    L.adsb_profile_enable.restype = i32
    L.adsb_profile_enable.argtypes = [vp, i32]
    L.adsb_profile_last.restype = i32
    L.adsb_profile_last.argtypes = [vp, C.POINTER(f32), C.POINTER(f32)]
    _lib = L
    return L


class AdsbError(RuntimeError):
    def __init__(self, code, detail=""):
        self.code = code
        msg = lib().adsb_strerror(code).decode()
        RuntimeError.__init__(self, "adsb_b200: %s (%d)%s" % (msg, code, (": " + detail) if detail else ""))


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class Context(object):
    """One `adsb_ctx`: device buffers + the framer state of one block instance."""

    def __init__(self, sps, max_items, device=0):
        self._h = C.c_void_p()
        self.sps = int(sps)
        self.hist = 8 * self.sps - 1
        self.max_items = int(max_items)
        self.device = int(device)
        rc = lib().adsb_create(C.byref(self._h), self.device, self.sps, self.max_items)
        if rc != ADSB_OK:
            self._h = C.c_void_p()
            raise AdsbError(rc, "adsb_create(device=%d, sps=%d, max_items=%d)" % (device, sps, max_items))

    def close(self):
        if self._h:
            lib().adsb_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != ADSB_OK:
            raise AdsbError(rc, lib().adsb_last_error(self._h).decode())

    # --- state ---------------------------------------------------------------------------
    def reset_state(self):
        self._check(lib().adsb_reset_state(self._h))

    def get_state(self):
        a, b = C.c_float(), C.c_int64()
        self._check(lib().adsb_get_state(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def set_state(self, prev_in0, prev_eob_idx):
        self._check(lib().adsb_set_state(self._h, float(prev_in0), int(prev_eob_idx)))

    def set_error_correction(self, mode):
        """FEC_NONE or FEC_CONSERVATIVE (decoder.py:765-780), for the fused entry points."""
        self._check(lib().adsb_set_error_correction(self._h, int(mode)))

    def counters(self):
        c = Counters()
        self._check(lib().adsb_get_counters(self._h, C.byref(c)))
        return c.as_dict()

    # --- block entry points -----------------------------------------------------------------
    def framer_work(self, in0, n_out, threshold):
        """-> (window indices int64[], snr float32[])"""
        in0 = np.ascontiguousarray(in0, dtype=np.float32)
        cap = int(n_out) // (63 * self.sps) + 2
        idx = np.empty(cap, np.int64)
        snr = np.empty(cap, np.float32)
        n = C.c_size_t(0)
        self._check(lib().adsb_framer_work(self._h, _ptr(in0), len(in0), int(n_out), float(np.float32(threshold)),
                                           _ptr(idx), _ptr(snr), cap, C.byref(n)))
        return idx[:n.value], snr[:n.value]

    def demod_work(self, in0, sob_idx):
        """-> (bits uint8[n_tags,112], valid uint8[n_tags])"""
        in0 = np.ascontiguousarray(in0, dtype=np.float32)
        sob = np.ascontiguousarray(sob_idx, dtype=np.int64)
        bits = np.zeros((len(sob), 112), np.uint8)
        valid = np.zeros(len(sob), np.uint8)
        self._check(lib().adsb_demod_work(self._h, _ptr(in0), len(in0), _ptr(sob), len(sob), _ptr(bits), _ptr(valid)))
        return bits, valid

    def burst_work(self, buf, n_out=None, lead=None, threshold=0.01, base=0, cap=None):
        """Fused path over a host buffer (float32 |IQ|^2, complex64 IQ, int16[n,2] sc16 IQ or uint8[n,2] cu8 IQ)
        -> structured hits."""
        buf = np.ascontiguousarray(buf)
        if buf.dtype == np.complex64:
            c64 = FMT_C64
        elif buf.dtype == np.float32:
            c64 = FMT_F32
        elif buf.dtype == np.int16 and buf.ndim == 2 and buf.shape[1] == 2:
            c64 = FMT_SC16                       # interleaved int16 I/Q, value = int * 2**-15
        elif buf.dtype == np.uint8 and buf.ndim == 2 and buf.shape[1] == 2:
            c64 = FMT_CU8                        # interleaved uint8 I/Q, value = (u - 127.5) / 128 (RTL-SDR)
        else:
            raise TypeError("burst_work takes float32 |IQ|^2, complex64 IQ, int16[n,2] IQ or uint8[n,2] IQ")
        n_in = len(buf)
        if lead is None:
            lead = self.hist if n_out is None else int(n_out) + self.hist - n_in
        if n_out is None:
            n_out = n_in + lead - self.hist
        cap = cap or (int(n_out) // (63 * self.sps) + 2)
        hits = np.zeros(cap, HIT_DTYPE)
        n = C.c_size_t(0)
        self._check(lib().adsb_burst_work(self._h, _ptr(buf), c64, n_in, int(lead), int(n_out),
                                          float(np.float32(threshold)), int(base), _ptr(hits), cap, C.byref(n)))
        return hits[:n.value]

    def burst_work_raw(self, host_ptr, fmt, n_in, lead, n_out, thr32, hits_ptr, cap, n_ref):
        """The bare call for per-work() use (blocks.py): everything pre-converted by the caller -- host pointer (int),
        threshold already rounded to float32, pointer / capacity of a reusable adsb_hit array, a ctypes size_t to
        receive the count.  Returns the number of hits."""
        rc = lib().adsb_burst_work(self._h, host_ptr, fmt, n_in, lead, n_out, thr32, 0, hits_ptr, cap, n_ref)
        if rc != ADSB_OK:
            self._check(rc)
        return n_ref._obj.value

    def burst_work_ptr(self, host_ptr, c64, n_in, n_out, lead, threshold, base, hits):
        """Same, raw host pointer + caller-owned hits array (bench.py's end-to-end leg)."""
        n = C.c_size_t(0)
        self._check(lib().adsb_burst_work(self._h, C.c_void_p(host_ptr), int(c64), int(n_in), int(lead), int(n_out),
                                          float(np.float32(threshold)), int(base), _ptr(hits), len(hits), C.byref(n)))
        return n.value

    def burst_work_dev(self, dev_ptr, c64, n_in, n_out, lead, threshold, base=0, eob_in=-1, halo_left=0, stream=0):
        """Device-resident input; asynchronous.  Returns the device pointer of the hit records."""
        out = C.c_void_p()
        self._check(lib().adsb_burst_work_dev(self._h, C.c_void_p(dev_ptr), int(c64), int(n_in), int(lead), int(n_out),
                                              int(halo_left), float(np.float32(threshold)), int(base), int(eob_in),
                                              C.c_void_p(stream), C.byref(out)))
        return out.value

    def burst_finish(self, fetch_hits=True, cap=None):
        """Counters of the latest resolve (dict) and, optionally, the hit records (structured array)."""
        c = Counters()
        n = C.c_size_t(0)
        if not fetch_hits:
            self._check(lib().adsb_burst_finish(self._h, C.byref(c), None, 0, C.byref(n)))
            return c.as_dict(), None
        cap = cap or 1 << 16
        hits = np.zeros(cap, HIT_DTYPE)
        rc = lib().adsb_burst_finish(self._h, C.byref(c), _ptr(hits), cap, C.byref(n))
        if rc == E_CAPACITY:
            raise AdsbError(rc, "burst_finish: pass cap >= %d" % n.value)
        self._check(rc)
        return c.as_dict(), hits[:n.value]

    def burst_mag_out(self, out):
        """Fill the float32 array `out` with |IQ|^2 of the last len(out) items of the latest burst_work input
        (computed on the device from the staged copy)."""
        assert out.dtype == np.float32 and out.flags["C_CONTIGUOUS"]
        self._check(lib().adsb_burst_mag_out(self._h, _ptr(out), len(out)))
        return out

    # --- multi-GPU carry (gr_adsb_b200/sharding.py) -------------------------------------------------
    def burst_reresolve_dev(self, eob_in, stream=0):
        self._check(lib().adsb_burst_reresolve_dev(self._h, int(eob_in), C.c_void_p(stream)))

    def burst_export_carry_dev(self, used_global, dev_ptr, stream=0):
        self._check(lib().adsb_burst_export_carry_dev(self._h, int(used_global), C.c_void_p(dev_ptr), C.c_void_p(stream)))

    def burst_carry_round_dev(self, gathered_ptr, world, rank, final_check=False, stream=0):
        self._check(lib().adsb_burst_carry_round_dev(self._h, C.c_void_p(gathered_ptr), int(world), int(rank),
                                                      1 if final_check else 0, C.c_void_p(stream)))

    # --- multi-GPU peer mode (include/adsb_b200.h: adsb_shard_*) -----------------------------------------
    def shard_alloc(self, sample_bytes):
        """-> (device pointer of the sample area, 64-byte CUDA IPC handle of the allocation)"""
        ptr = C.c_void_p()
        handle = C.create_string_buffer(IPC_HANDLE_BYTES)
        self._check(lib().adsb_shard_alloc(self._h, int(sample_bytes), C.byref(ptr), handle))
        return ptr.value, handle.raw

    def shard_attach(self, rank, world, handles):
        """handles: the `world` IPC handles (bytes, rank order) exchanged by any transport"""
        blob = b"".join(handles) if not isinstance(handles, (bytes, bytearray)) else bytes(handles)
        assert len(blob) == world * IPC_HANDLE_BYTES
        self._check(lib().adsb_shard_attach(self._h, int(rank), int(world), C.c_char_p(blob)))

    def shard_attach_ptrs(self, rank, world, sample_ptrs):
        """Ranks living in one process: `sample_ptrs[r]` = what rank r's shard_alloc returned."""
        arr = (C.c_void_p * world)(*[int(p) - SHARD_CTRL_BYTES for p in sample_ptrs])
        self._check(lib().adsb_shard_attach_ptrs(self._h, int(rank), int(world), arr))

    def shard_detach(self):
        self._check(lib().adsb_shard_detach(self._h))

    def shard_write(self, src_ptr, nbytes, byte_offset=0, stream=None):
        self._check(lib().adsb_shard_write(self._h, int(byte_offset), C.c_void_p(int(src_ptr)), int(nbytes),
                                           C.c_void_p(stream) if stream else None))

    def shard_step(self, fmt, shard_items, threshold, phases=SHARD_PH_ALL, use_graph=False, stream=0):
        self._check(lib().adsb_shard_step(self._h, int(fmt), int(shard_items), float(np.float32(threshold)), int(phases),
                                          1 if use_graph else 0, C.c_void_p(stream)))

    # --- streaming session (include/adsb_b200.h: adsb_stream_*) ---------------------------------------------
    _FMT_DTYPES = {FMT_F32: (np.float32, 1), FMT_C64: (np.complex64, 1), FMT_SC16: (np.int16, 2), FMT_CU8: (np.uint8, 2)}

    def stream_open(self, fmt, threshold, chunk_items):
        self._check(lib().adsb_stream_open(self._h, int(fmt), float(np.float32(threshold)), int(chunk_items)))
        self._stream_fmt = int(fmt)

    def stream_push(self, samples):
        """Append samples (any length): float32 |IQ|^2, complex64, int16[n,2] or uint8[n,2] as the session's format says."""
        dt, width = self._FMT_DTYPES[self._stream_fmt]
        a = np.ascontiguousarray(samples, dtype=dt)
        n = a.size // width
        self._check(lib().adsb_stream_push(self._h, _ptr(a), n))

    def stream_buffer(self):
        """The free part of the pinned staging buffer as a NumPy array to fill in place (np.fromfile / readinto);
        follow with stream_commit(items_written)."""
        p, cap = C.c_void_p(), C.c_size_t()
        self._check(lib().adsb_stream_acquire(self._h, C.byref(p), C.byref(cap)))
        dt, width = self._FMT_DTYPES[self._stream_fmt]
        n = cap.value * width
        buf = (C.c_char * (n * np.dtype(dt).itemsize)).from_address(p.value)
        arr = np.frombuffer(buf, dtype=dt, count=n)
        return arr.reshape(-1, 2) if width == 2 else arr

    def stream_commit(self, n_items):
        self._check(lib().adsb_stream_commit(self._h, int(n_items)))

    def stream_pop(self, cap=1 << 16):
        """Hits completed so far (structured array, possibly empty), in stream order."""
        out = []
        while True:
            hits = np.zeros(cap, HIT_DTYPE)
            n = C.c_size_t(0)
            self._check(lib().adsb_stream_pop(self._h, _ptr(hits), cap, C.byref(n)))
            out.append(hits[:n.value])
            if n.value < cap:
                break
        return np.concatenate(out) if len(out) > 1 else out[0]

    def stream_flush(self):
        self._check(lib().adsb_stream_flush(self._h))

    def stream_close(self):
        self._check(lib().adsb_stream_close(self._h))

    def stream_stats(self):
        v = [C.c_uint64() for _ in range(4)]
        self._check(lib().adsb_stream_stats(self._h, *[C.byref(x) for x in v]))
        return dict(zip(("items", "chunks_done", "hits", "crc_ok"), (x.value for x in v)))

    def latency_stats(self):
        """(calls served by the one-kernel latency path, how many of them fell back to the resolve kernels)"""
        a, b = C.c_uint64(), C.c_uint64()
        self._check(lib().adsb_latency_stats(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def set_latency_path(self, on):
        self._check(lib().adsb_set_latency_path(self._h, 1 if on else 0))

    # --- measurement hook ------------------------------------------------------------------------------
    def profile_enable(self, on=True):
        self._check(lib().adsb_profile_enable(self._h, 1 if on else 0))

    def profile_last(self):
        """(scan_ms, resolve_ms) of the latest burst_work_dev, from CUDA events on its launch stream."""
        a, b = C.c_float(), C.c_float()
        self._check(lib().adsb_profile_last(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value



def unpack_bits(hits):
    """hits['bits'] (packed, MSB first) -> uint8[n,112]"""
    if len(hits) == 0:
        return np.zeros((0, 112), np.uint8)
    return np.unpackbits(np.ascontiguousarray(hits["bits"]), axis=1)[:, :112]
