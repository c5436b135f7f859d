/*
 * adsb_b200.h -- C ABI of the B200-native ADS-B burst-detect + PPM-demod path.
 *
 * gr-adsb has no FFI of its own: its hot path is two pure-Python GNU Radio blocks.  This
 * header is the boundary a native replacement for that path binds to; each entry point names
 * the reference code it stands in for (paths relative to the gr-adsb tree):
 *
 *   adsb_framer_work      python/adsb/framer.py:72-182   framer.work()
 *   adsb_demod_work       python/adsb/demod.py:57-136    demod.work() (the PPM slice, :82-95)
 *   adsb_burst_work*      examples/adsb_rx.py:180-196    complex_to_mag_squared -> framer -> demod
 *                                                        (+ the parity check of decoder.py:667-685)
 *   adsb_reset_state / adsb_get_state / adsb_set_state
 *                         python/adsb/framer.py:54,57    prev_in0 / prev_eob_idx carried between calls
 *
 * Conventions: plain pointers and sizes only; every function returns ADSB_OK (0) or a negative
 * ADSB_E_* code and never throws; nothing is written past the capacities the caller passes.
 * A context is owned by one block instance and is not thread safe (GNU Radio calls a block's
 * work() from one thread).  Host entry points take HOST pointers that are only read/written
 * during the call (GNU Radio buffers are only valid for the duration of work()).
 * "Window coordinates" below are indices into the in0 array handed to framer.work():
 * window index i is absolute item  nitems_written(0) - (8*sps - 1) + i  (framer.py:170).
 */
#ifndef ADSB_B200_H
#define ADSB_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ADSB_B200_ABI_VERSION 1

#define ADSB_FMT_F32 0
#define ADSB_FMT_C64 1
#define ADSB_FMT_SC16 2
#define ADSB_FMT_CU8 3

enum {
    ADSB_OK = 0,
    ADSB_E_INVAL = -1,     /* bad argument (odd sps, null pointer, sizes that do not fit max_items ...) */
    ADSB_E_CUDA = -2,      /* a CUDA runtime call failed; adsb_last_error() has the text            */
    ADSB_E_NOMEM = -3,     /* host or device allocation failed                                       */
    ADSB_E_CAPACITY = -4,  /* more hits than the caller's `cap`; n_hits holds the number required      */
    ADSB_E_NODEVICE = -5,  /* no CUDA device / not an sm_100 part                                     */
    ADSB_E_HALO = -6,      /* a pulse run crosses the left edge of a halo'd shard (multi-GPU only)    */
    ADSB_E_INTERNAL = -7,
    ADSB_E_PEER = -8       /* multi-GPU peer mode: IPC mapping failed or a neighbouring rank never reached the step */
};

/* adsb_hit.flags */
#define ADSB_HIT_DEMOD_VALID 0x1u /* burst lies fully inside the buffer: bits[] valid (demod.py:82)            */
#define ADSB_HIT_CRC_OK      0x2u /* DF in {17,18,19} and PI == CRC-24(bits[0:88]) (decoder.py:667-685)         */
#define ADSB_HIT_FEC_FIXED   0x8u /* parity restored by the conservative error correction (then CRC_OK is set too)  */
#define ADSB_HIT_LATE_TAG    0x4u /* window index < 8*sps-1: tag lands on an item of the previous call (A.9) */

/* One detected burst.  `bits` holds the 112 demodulated bits packed MSB first (numpy.packbits
 * order: message bit k is bit 7-(k%8) of bits[k/8]); bits[14..15] are zero. */
typedef struct adsb_hit {
    int64_t idx;       /* pulse-centre index p: window coordinates + the `base` given to the call     */
    float snr;         /* 10*log10(x[p]/median(x[p-100:p])) + 1.6 dB (framer.py:156-159); nan/inf kept */
    uint32_t flags;    /* ADSB_HIT_*                                                                   */
    uint8_t bits[16];
} adsb_hit;

/* Per-call counters (all in items of the call's window). */
typedef struct adsb_counters {
    uint64_t n_samples;     /* items scanned (n_out)                                       */
    uint64_t n_candidates;  /* pulse centres whose 16 half-symbol taps match the preamble  */
    uint64_t n_hits;        /* candidates surviving the 63-symbol blanking -> "burst" tags */
    uint64_t n_demod;       /* hits with ADSB_HIT_DEMOD_VALID                              */
    uint64_t n_crc_ok;      /* hits with ADSB_HIT_CRC_OK                                   */
    int64_t last_pulse;     /* centre of the last complete pulse, -1 if none               */
    int64_t last_accept;    /* centre of the last hit, window coordinates, -1 if none       */
    int64_t first_candidate;/* centre of the first candidate, window coordinates, -1 if none */
    uint32_t carry_unstable; /* device-side carry rounds whose final check still wanted a fix-up  */
    uint32_t n_fec_fixed;    /* hits with ADSB_HIT_FEC_FIXED (not included in n_crc_ok)            */
    uint32_t any_rise;      /* framer.py:95 guard inputs                                   */
    uint32_t any_fall;
} adsb_counters;

typedef struct adsb_ctx adsb_ctx;

const char *adsb_strerror(int code);
int adsb_abi_version(void);
int adsb_device_count(void);

/* Create a context on CUDA device `device` for `sps` samples per symbol (fs / 1e6; even, 2..32)
 * able to take up to `max_items` input items per call.  Device buffers are sized once here. */
int adsb_create(adsb_ctx **ctx, int device, int sps, size_t max_items);
void adsb_destroy(adsb_ctx *ctx);
const char *adsb_last_error(const adsb_ctx *ctx);

/* framer.py:54,57 -- the state carried between work() calls. reset: prev_in0 = 0, prev_eob_idx = -1 */
int adsb_reset_state(adsb_ctx *ctx);
int adsb_get_state(const adsb_ctx *ctx, float *prev_in0, int64_t *prev_eob_idx);
int adsb_set_state(adsb_ctx *ctx, float prev_in0, int64_t prev_eob_idx);
int adsb_get_counters(const adsb_ctx *ctx, adsb_counters *out);

/* decoder.py:765-780 error_corr: ADSB_FEC_NONE (default, "None") or ADSB_FEC_CONSERVATIVE ("Conservative":
 * decoder.py:304-323,738-763 -- hits that fail the parity check with DF 17..19 whose 25-bit residual matches a
 * 1-bit or 2-contiguous-bit error pattern get those bits flipped, ADSB_HIT_CRC_OK | ADSB_HIT_FEC_FIXED set).
 * Applies to the fused entry points (adsb_burst_work*); the block API leaves correction to adsb.decoder. */
#define ADSB_FEC_NONE 0
#define ADSB_FEC_CONSERVATIVE 1
int adsb_set_error_correction(adsb_ctx *ctx, int mode);

/* framer.work(): in0[n_in] is float32 |IQ|^2 with n_in == n_out + 8*sps - 1 (GNU Radio history).
 * Writes the pulse-centre window index and SNR of every "burst" tag, ascending; the caller adds
 * nitems_written(0) - (8*sps-1) to get the tag offset (framer.py:168-174) and copies
 * in0[8*sps-1:] to out0 (framer.py:181).  Updates the carried state exactly as the reference
 * does (including the stale prev_eob_idx behaviour). */
int adsb_framer_work(adsb_ctx *ctx, const float *in0, size_t n_in, size_t n_out, float threshold,
                     int64_t *hit_idx, float *hit_snr, size_t cap, size_t *n_hits);

/* demod.work() slicing: for each tag t, sob_idx[t] = tag.offset + 8*sps - nitems_written(0)
 * (demod.py:75,79).  valid[t] = 1 and bits[t*112 .. t*112+111] (one bit per byte) are written
 * when the burst ends inside in0[0:n] (demod.py:82); otherwise valid[t] = 0 (straddler,
 * demod.py:130-133) and its bits are left untouched.  Stateless. */
int adsb_demod_work(adsb_ctx *ctx, const float *in0, size_t n, const int64_t *sob_idx, size_t n_tags,
                    uint8_t *bits, uint8_t *valid);

/* Fused path over a HOST buffer: |IQ|^2 where needed, framer, demod and parity check in one pass on the
 * GPU.  `is_c64` selects the item format: ADSB_FMT_F32 (0) float32 |IQ|^2, ADSB_FMT_C64 (1) interleaved re,im
 * float32 pairs, ADSB_FMT_SC16 (2) interleaved re,im int16 with value = int * 2^-15 (UHD sc16 wire format /
 * blocks.interleaved_short_to_complex with scale 32768), ADSB_FMT_CU8 (3) interleaved re,im uint8 with value =
 * (u - 127.5) / 128 (RTL-SDR wire format); n_in counts items (complex samples).  The window is `lead` zero items followed by in[0:n_in]; n_out + 8*sps - 1 == lead + n_in.
 * lead = 8*sps-1 with a freshly reset context reproduces a flowgraph's first work() call over
 * the whole buffer ("whole-buffer mode"); lead = 0 continues a stream whose history the caller
 * prepends, exactly like adsb_framer_work.  hits[i].idx = base + window index. */
int adsb_burst_work(adsb_ctx *ctx, const void *in, int is_c64, size_t n_in, size_t lead, size_t n_out,
                    float threshold, int64_t base, adsb_hit *hits, size_t cap, size_t *n_hits);

/* out[0:n_out] = |IQ|^2 of the LAST n_out items handed to the most recent adsb_burst_work, computed on the
 * device from the copy still staged there (ADSB_FMT_F32 input: the items themselves).  This is the stream
 * blocks.complex_to_mag_squared (examples/adsb_rx.py:180,195-197) fed to the framer and the time sink; the fused
 * "burst" block with complex input writes it to its output port, as framer.py:181 does with its input. */
int adsb_burst_mag_out(adsb_ctx *ctx, float *out, size_t n_out);

/* Same, for a buffer already resident in device memory; runs asynchronously on CUDA stream
 * `stream` (a cudaStream_t, may be NULL = the context's own stream) and leaves the results on the
 * device (n_out + 8*sps - 1 <= lead + n_in here: items past that are a right halo): *d_hits is a device pointer to adsb_hit[] owned by the context (valid until the next
 * call); counts come back through adsb_burst_finish.  `eob_in` is the incoming prev_eob_idx in
 * window coordinates (-1 at stream start).  `halo_left` > 0 says the first `halo_left` items of d_in
 * are a left halo copied from the neighbouring shard: pulses are owned from window index halo_left
 * on, the window start is not the stream start, and the carried state is left alone (multi-GPU). */
int adsb_burst_work_dev(adsb_ctx *ctx, const void *d_in, int is_c64, size_t n_in, size_t lead, size_t n_out,
                        size_t halo_left, float threshold, int64_t base, int64_t eob_in, void *stream,
                        const adsb_hit **d_hits);

/* After adsb_burst_work_dev + stream sync: fetch counters, advance the carried state as
 * framer.py:121-123,165,177-179 would, optionally copy the hits to the host. */
int adsb_burst_finish(adsb_ctx *ctx, adsb_counters *counters, adsb_hit *hits, size_t cap, size_t *n_hits);

/* Multi-GPU carry fix-up: re-apply the blanking rule to the candidates of the last
 * adsb_burst_work_dev call with another incoming prev_eob_idx (window coordinates) -- only the two
 * tiny resolve kernels run again.  Follow with adsb_burst_finish (which may be called repeatedly;
 * each call re-reads the counters of the latest resolve). */
int adsb_burst_reresolve_dev(adsb_ctx *ctx, int64_t eob_in, void *stream);

/* Multi-GPU: after adsb_burst_work_dev / adsb_burst_reresolve_dev, write (first candidate, last hit,
 * used_global) in GLOBAL window coordinates (-(2^62) where absent) to the device array d_out3[3] on the
 * same stream, ready to be all-gathered without a host round trip (gr_adsb_b200/sharding.py). */
#define ADSB_CARRY_USED_ON_DEVICE INT64_MIN   /* used_global: take the value the device-side plan recorded */
int adsb_burst_export_carry_dev(adsb_ctx *ctx, int64_t used_global, int64_t *d_out3, void *stream);

/* Multi-GPU, no host round trip: d_gathered[world][3] is the all-gathered export of every shard.  Evaluates
 * the carry plan on the device for shard `rank`; unless `final_check`, follows with the resolve kernels in
 * conditional mode (they exit at once when this shard needs no fix-up).  With `final_check` nothing is
 * re-run; if any shard would still need one, adsb_counters.carry_unstable is bumped (read it after the
 * stream is idle: non-zero means more rounds were needed than were issued). */
int adsb_burst_carry_round_dev(adsb_ctx *ctx, const int64_t *d_gathered, int world, int rank, int final_check,
                               void *stream);

/* ---- Multi-GPU peer mode: one long stream cut into `world` equal contiguous shards, one per GPU/process --------------
 * (target semantics = ONE framer.work()/demod.work() call over the whole stream, framer.py:72-182,
 * demod.py:57-136.)  Nothing is copied between GPUs: every rank keeps its shard in a "symmetric" allocation that the
 * other ranks map (CUDA IPC over NVLink / NVSwitch); the scan kernel's edge tiles read the neighbouring shards in
 * place -- the 100-sample noise window and pulse-run look-back to the left (framer.py:95-113,156-159), preamble + 112
 * bits to the right (demod.py:75-95) -- and the one sequential quantity, the blanking state prev_eob_idx
 * (framer.py:121-123,165), travels as three int64 per rank written straight into every rank's control block.
 * Neither the host nor NCCL takes part in a step; the whole step is one CUDA graph launch.
 *
 *   adsb_shard_alloc        allocate [control block | samples] on the context's device, return the sample pointer and
 *                           (optionally) the 64-byte CUDA IPC handle to hand to the other ranks (any transport)
 *   adsb_shard_attach       map every rank's allocation from its IPC handle (handles[world][64], own entry ignored)
 *   adsb_shard_attach_ptrs  same with ready-made device pointers to the START of each rank's allocation (ranks that
 *                           live in one process, or memory mapped by other means); pointers stay owned by the caller
 *   adsb_shard_write        copy a step's samples into this rank's sample area (from pinned host memory: the ingest)
 *   adsb_shard_step         one step over this rank's `shard_items` samples of format `fmt` (ADSB_FMT_*), asynchronous
 *                           on `stream`; every rank calls it with the same arguments.  phases = ADSB_SHARD_PH_ALL runs
 *                           the whole step; the individual bits exist so that a single process driving several ranks
 *                           (tests) can interleave them, and EXPORT|PLAN (then CHECK) repeats the carry exchange when
 *                           adsb_counters.carry_unstable reports that one round was not enough (cascade through
 *                           several shards).  use_graph != 0 replays the step as a captured CUDA graph.
 *   results                 adsb_burst_finish (counters, hits with global window indices); the carried framer state of
 *                           the context is left alone.  Waits on other ranks are bounded: ADSB_E_PEER after 2 s.
 * Limits: shard_items >= 120*sps + 256; a pulse run or noise window reaching past the START of the left neighbour's
 * shard is reported as ADSB_E_HALO. */
#define ADSB_SHARD_CTRL_BYTES 8192
#define ADSB_IPC_HANDLE_BYTES 64
#define ADSB_SHARD_MAX_WORLD 32
#define ADSB_SHARD_PH_SIGNAL 1   /* tell the neighbours that this rank's samples of the step are resident          */
#define ADSB_SHARD_PH_SCAN   2   /* wait for the neighbours, scan + resolve assuming no incoming blanking state      */
#define ADSB_SHARD_PH_EXPORT 4   /* publish (first candidate, last hit, state assumed) to every rank                 */
#define ADSB_SHARD_PH_PLAN   8   /* wait for every rank's, re-resolve if the true incoming state demands, re-publish */
#define ADSB_SHARD_PH_CHECK  16  /* wait for every rank's, count the step as unstable if a fix-up is still needed     */
#define ADSB_SHARD_PH_ALL    15  /* one step.  The CHECK of a step is made by the next step's EXPORT (it then costs no second
                                    rendezvous), so adsb_counters.carry_unstable lags one step: after the last step of a run
                                    issue ADSB_SHARD_PH_CHECK once before reading it. */
int adsb_shard_alloc(adsb_ctx *ctx, size_t sample_bytes, void **d_samples, void *ipc_handle);
int adsb_shard_attach(adsb_ctx *ctx, int rank, int world, const void *ipc_handles);
int adsb_shard_attach_ptrs(adsb_ctx *ctx, int rank, int world, void *const *peer_bases);
int adsb_shard_detach(adsb_ctx *ctx);
/* Copy `bytes` from `src` (host or device memory, pinned host memory for a truly asynchronous copy) to byte_offset of
 * this rank's sample area: the ingest of a step.  stream == NULL: the context's stream, synchronised before returning. */
int adsb_shard_write(adsb_ctx *ctx, size_t byte_offset, const void *src, size_t bytes, void *stream);
int adsb_shard_step(adsb_ctx *ctx, int fmt, size_t shard_items, float threshold, int phases, int use_graph,
                    void *stream);

/* ---- Streaming session: one long recording / SDR stream pushed from the host in pieces ---------------------------------
 * (the ingest side of examples/adsb_rx.py:113-126,180.)  Result = ONE framer.work() + demod.work() call over
 * the whole stream (framer.py:72-182, demod.py:57-136), whatever the sizes of the pushes, for streams far larger than
 * device memory.  The context keeps two pinned staging buffers and three device buffers of `chunk_items`: while chunk
 * k+1 travels over PCIe, chunk k is scanned with its neighbours k-1 and k+1 read in place (noise window and pulse-run
 * look-back to the left, preamble + 112 bits to the right -- no overlap is copied twice) and the blanking state
 * (framer.py:121-123,165) is handed from chunk to chunk.  Sample formats as adsb_burst_work (int16 / uint8 IQ are
 * widened on the device: half / a quarter of the PCIe bytes).
 *   adsb_stream_open     start a session (ends any previous one); chunk_items <= max_items of adsb_create
 *   adsb_stream_push     append n_items samples from any host memory (copied into the pinned staging buffer)
 *   adsb_stream_acquire  ... or get the free part of the staging buffer to fill in place (fread, SDR driver), then
 *   adsb_stream_commit   say how many items were written; a full buffer is sent off, the call returns when the other
 *                        staging buffer is free again
 *   adsb_stream_pop      hits completed so far, in stream order (idx = stream index + 8*sps-1, as adsb_burst_work with
 *                        lead = 8*sps-1); never blocks
 *   adsb_stream_flush    end of stream: scan what is buffered with the tail semantics of the single call (the last
 *                        8*sps-1 items are look-ahead only; bursts that do not end inside the stream are not demodulated)
 *   adsb_stream_close    free the session's buffers (adsb_destroy does it too)
 * The session uses the context's device buffers: do not interleave other entry points of the same context. */
int adsb_stream_open(adsb_ctx *ctx, int fmt, float threshold, size_t chunk_items);
int adsb_stream_push(adsb_ctx *ctx, const void *samples, size_t n_items);
int adsb_stream_acquire(adsb_ctx *ctx, void **host_buf, size_t *cap_items);
int adsb_stream_commit(adsb_ctx *ctx, size_t n_items);
int adsb_stream_pop(adsb_ctx *ctx, adsb_hit *hits, size_t cap, size_t *n_hits);
int adsb_stream_flush(adsb_ctx *ctx);
int adsb_stream_close(adsb_ctx *ctx);
int adsb_stream_stats(const adsb_ctx *ctx, uint64_t *items, uint64_t *chunks_done, uint64_t *hits, uint64_t *crc_ok);

/* The latency path.  Calls of GNU Radio size (float32 or complex64 input, at most 64 tiles = 320 Ki items at 2 Msps,
 * no error correction) do not go through the copy / multi-kernel pipeline: adsb_framer_work, adsb_demod_work and
 * adsb_burst_work stage the caller's (pageable) buffer through pinned, device-mapped memory, launch ONE kernel that
 * scans, slices and resolves the call and writes hits and counters back into pinned memory, and spin on its
 * completion stamp -- same results, same state updates (framer.py:72-182, demod.py:57-136).  adsb_latency_stats:
 * calls served that way and how many of them had more candidates than the in-kernel resolve holds (finished by the
 * resolve kernels instead).  adsb_set_latency_path(ctx, 0) forces the throughput pipeline (A/B measurements; the
 * environment variable ADSB_NO_LATENCY_PATH=1 does the same for every new context). */
int adsb_latency_stats(const adsb_ctx *ctx, uint64_t *calls, uint64_t *fallbacks);
int adsb_set_latency_path(adsb_ctx *ctx, int on);
/* Diagnostics of the latest latency-path call: host3 = seconds spent staging the input, launching, and waiting for the
 * completion stamp; dev8 = device %globaltimer stamps of the resolving CTA (kernel entry, -, tile detected, sliced,
 * last-tile ticket, resolved), only filled by builds with -DADSB_SMALL_TIMING=1 (tools/latency_probe.py). */
int adsb_latency_trace(const adsb_ctx *ctx, uint64_t *dev8, double *host3);

/* Measurement hook for bench.py: when enabled, adsb_burst_work_dev records CUDA events on its launch
 * stream before the scan kernel, after it and after the resolve kernels; adsb_profile_last waits for
 * the last one and returns the two durations in milliseconds. */
int adsb_profile_enable(adsb_ctx *ctx, int on);
int adsb_profile_last(adsb_ctx *ctx, float *scan_ms, float *resolve_ms);

#ifdef __cplusplus
}
#endif
#endif /* ADSB_B200_H */
