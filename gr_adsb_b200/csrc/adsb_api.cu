// C ABI (include/adsb_b200.h) over the sm_100a kernels.  No torch types, no exceptions.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include <new>
#if !defined(ADSB_HOST_EMU)
#include <thread>
#endif

#if !defined(ADSB_HOST_EMU)
#include <nvtx3/nvToolsExt.h>          // header-only NVTX v3: ranges cost nothing unless a profiler is attached
#endif

#include "../../include/adsb_b200.h"
#include "adsb_device.cuh"
#include "adsb_resolve.cuh"
#include "adsb_scan.cuh"
#include "adsb_small.cuh"

using namespace adsb;

static_assert(sizeof(adsb_hit) == sizeof(Rec), "adsb_hit layout");
static_assert(kTilesPerChunk == kThreads, "K1 accumulates tile_chunk_tot[tile / kThreads]");
static_assert(ADSB_SHARD_MAX_WORLD == kMaxWorld, "header and kernels agree on the largest world");

namespace {

constexpr int kFirstHits = 64;          // hits fetched together with the header
constexpr size_t kPipeChunk = 1u << 22; // items per H2D chunk on the host fused path
constexpr int kEvRing = 4;

__global__ void init_header_kernel(Header *h)
{
    const unsigned unstable = h->carry_unstable;       // sticky across calls until read by adsb_burst_finish
    memset(h, 0, sizeof(Header));
    h->last_pulse = -1;
    h->last_accept = -1;
    h->first_cand = -1;
    h->carry_used = kCarryNone;
    h->carry_ein = kCarryNone;
    h->carry_unstable = unstable;
}

// multi-GPU blanking carry: (first candidate, last hit, incoming state assumed) in GLOBAL window coordinates,
// kNone where absent -- the triple gr_adsb_b200/sharding.py:carry_plan consumes
__global__ void export_carry_kernel(const Header *h, long long base, long long used_global, int used_from_dev,
                                    long long *out)
{
    out[0] = h->first_cand >= 0 ? base + h->first_cand : kCarryNone;
    out[1] = h->last_accept >= 0 ? base + h->last_accept : kCarryNone;
    out[2] = used_from_dev ? h->carry_used : used_global;
}

// |IQ|^2 stream of a staged input (what blocks.complex_to_mag_squared handed downstream, adsb_rx.py:180)
template <int FMT>
__global__ void mag_out_kernel(const void *in, long long first, long long n, float *out)
{
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        out[i] = load_x<FMT>(in, first + i);
}

struct HostStage {
    Header hdr;
    Rec first[kFirstHits];
};

}  // namespace

struct adsb_ctx {
    int device, sps, T, rh, hist;
    int nt;                      // threads per CTA of the scan kernel
    size_t max_items;
    float prev_in0;
    long long prev_eob;
    cudaStream_t stream, copy_stream;
    cudaEvent_t ev[kEvRing];
    void *d_in;
    size_t d_in_bytes;
    Rec *d_cand, *d_ord, *d_hits;
    unsigned cand_cap, hit_cap;
    uint2 *d_tile_tab;
    size_t tile_cap;
    unsigned *d_tile_chunk_tot, *d_cand_chunk_tot;   // per-chunk totals (ordered offsets without look-back)
    size_t tile_chunk_n, cand_chunk_n;
    Header *d_hdr;
    unsigned *d_crc;
    long long *d_sob;
    unsigned char *d_bits, *d_valid;
    size_t tag_cap;
    HostStage *h_stage;
    int num_sms;
    // bookkeeping of the call in flight (adsb_burst_work_dev -> adsb_burst_finish)
    struct {
        bool pending;
        long long n_out, lead, base, eob_in;
        int left_is_start;
        cudaStream_t stream;
    } last;
    adsb_counters counters;
    int profile;
    unsigned prefetch_tiles;     // L2 prefetch distance of the scan kernel, in tiles
    int fec_mode;                // ADSB_FEC_*
    unsigned *d_fec_t25, *d_fec_keys;
    unsigned short *d_fec_fix;
    unsigned rerun_count;
    cudaEvent_t pev[3];
    float *d_mag;                // |IQ|^2 output stream of adsb_burst_mag_out (allocated on first use)
    size_t d_mag_cap;
    int staged_fmt;              // format / item count of the input adsb_burst_work staged last (-1: none)
    size_t staged_n;
    bool small_configured[2];    // latency-path kernel attributes set (float32 / complex64 input)
    int small_fmt;
    int early_tile;              // ADSB_EARLY_TILE=0 in the environment: the final tile is scanned last, not second (A/B)
    int no_small;                // ADSB_NO_LATENCY_PATH=1 in the environment: always take the throughput pipeline (A/B, debugging)
    bool scan_configured[kNumFmt];  // per sample format: shared-memory attributes of the scan kernel set on this device
    // multi-GPU peer mode (adsb_shard_*): symmetric allocation [ShardCtrl | samples] of every rank, mapped here
    struct {
        int rank, world;         // world == 0: not attached
        void *own_base;          // this rank's allocation (adsb_shard_alloc); null when the caller brought its own
        size_t own_bytes;
        void *peer_base[kMaxWorld];   // every rank's allocation as seen from this device (own entry included)
        bool ipc_opened[kMaxWorld];
#if !defined(ADSB_HOST_EMU)
        cudaGraphExec_t exec;    // the captured step (null: none)
        cudaGraph_t graph;
#endif
        struct {
            int fmt, phases;
            size_t items;
            float thr;
            cudaStream_t stream;
        } key;                   // what `exec` was captured for
        unsigned long long steps_graph, steps_plain;
    } shard;
    // latency path (adsb_small.cuh): pinned, device-mapped staging of one small call
    struct {
        SmallOut *out;           // results + completion stamp
        void *in;                // the call's samples
        size_t in_bytes;
        float *mag;              // |IQ|^2 of the call's samples (complex input), n_in items
        size_t mag_items;
        long long *sob;          // demod: tag positions / results
        unsigned char *bits, *valid;
        size_t tag_cap;
        unsigned *d_done;        // device counter of small_demod_kernel
        unsigned long long seq;
        bool clean;              // the device header is as init_header_kernel leaves it
        bool last_was_small;     // adsb_burst_mag_out: the latest input sits in `in`, its |IQ|^2 in `mag`
        bool use_device_hits;    // the latest small call fell back to the resolve kernels: hits are in d_hits
        unsigned long long calls, fallbacks;
        double host_t[3];        // seconds: entry -> staged, -> launched, -> stamp seen (latest call)
    } small;
    // streaming session (adsb_stream_*): one long stream pushed from the host in pieces, whole-stream semantics
    struct {
        bool open, flushed;
        int fmt;
        float thr;
        size_t C, isz;               // chunk items, bytes per item
        void *d_buf[3];              // device chunks k % 3: a chunk's scan reads its left and right neighbour in place
        size_t d_items[3];
        void *h_pin[2];              // pinned staging, one chunk each
        int cur;                     // staging buffer being filled
        size_t fill;                 // items in it
        long long submitted;         // chunks handed to the copy engine
        long long processed;         // chunks scanned + resolved
        long long collected;         // chunks whose hits are in `out`
        unsigned long long total_items;
        long long e_global;          // blanking state after the latest collected chunk (global window index; kCarryNone: none)
        cudaEvent_t ev_copy[2], ev_h2d[3], ev_scan[3];
        Rec *out;                    // hits so far
        size_t out_n, out_cap, out_read;
        unsigned long long n_cand, n_crc, n_demod;
    } st;
    void *h_ring[2];             // adsb_burst_work from pageable memory: pinned staging, one H2D chunk each
    cudaEvent_t ev_ring[2];
    Rec *h_tmp;                  // host scratch of the two-block entry points
    size_t h_tmp_cap;
    unsigned char *h_bits;
    size_t h_bits_cap;
    char err[256];
};

namespace {

int fail(adsb_ctx *c, int code, const char *what, cudaError_t e = cudaSuccess)
{
    if (c) {
        if (e != cudaSuccess)
            snprintf(c->err, sizeof(c->err), "%s: %s", what, cudaGetErrorString(e));
        else
            snprintf(c->err, sizeof(c->err), "%s", what);
    }
    return code;
}

// Entry points run on the context's device and leave the caller's current device as they found it.
struct DeviceGuard {
    int prev;
    cudaError_t status;
    explicit DeviceGuard(int device) : prev(-1), status(cudaSuccess)
    {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != device) status = cudaSetDevice(device);
        else prev = -1;
    }
    ~DeviceGuard()
    {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

// one NVTX range per entry point (nsys / ncu --nvtx timelines show the C-ABI calls the blocks make)
struct NvtxRange {
#if !defined(ADSB_HOST_EMU)
    explicit NvtxRange(const char *name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
#else
    explicit NvtxRange(const char *) {}
#endif
};
#define ADSB_RANGE(name) NvtxRange nvtx_range_(name)

#define ON_DEVICE(ctx)                                                               \
    DeviceGuard guard_((ctx)->device);                                               \
    if (guard_.status != cudaSuccess) return fail((ctx), ADSB_E_CUDA, "cudaSetDevice", guard_.status)

#define CK(call)                                                          \
    do {                                                                  \
        cudaError_t e_ = (call);                                          \
        if (e_ != cudaSuccess) return fail(ctx, ADSB_E_CUDA, #call, e_);  \
    } while (0)

// decoder.compute_crc_2 (decoder.py:716-736) on a 112-bit pattern: reduce positions 0..86, return bits 87..111
unsigned residual25(const unsigned char *bits112)
{
    unsigned char d[kMsgBits];
    memcpy(d, bits112, kMsgBits);
    static const unsigned char poly[25] = {1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 0, 1, 0, 0, 0, 0, 0, 0, 1, 0, 0, 1};
    for (int i = 0; i < kMsgBits - 25; i++)
        if (d[i])
            for (int j = 0; j < 25; j++) d[i + j] ^= poly[j];
    unsigned v = 0;
    for (int i = kMsgBits - 25; i < kMsgBits; i++) v = (v << 1) | d[i];
    return v;
}

// tables of the conservative error correction (decoder.py:304-323): unit-vector residuals, and the residuals
// of every 1-bit / 2-contiguous-bit pattern sorted ascending with the bits to flip
void fec_tables(unsigned *t25, unsigned *keys, unsigned short *fix)
{
    unsigned char z[kMsgBits];
    for (int i = 0; i < kMsgBits; i++) {
        memset(z, 0, sizeof(z));
        z[i] = 1;
        t25[i] = residual25(z);
    }
    int n = 0;
    for (int burst = 1; burst <= 2; burst++)
        for (int i = 0; i + burst <= kMsgBits; i++) {
            unsigned k = 0;
            for (int q = 0; q < burst; q++) k ^= t25[i + q];
            keys[n] = k;
            fix[n] = (unsigned short)(i | (burst << 8));
            n++;
        }
    for (int a = 1; a < n; a++) {           // insertion sort by key (223 entries)
        const unsigned k = keys[a];
        const unsigned short f = fix[a];
        int b = a - 1;
        while (b >= 0 && keys[b] > k) {
            keys[b + 1] = keys[b];
            fix[b + 1] = fix[b];
            b--;
        }
        keys[b + 1] = k;
        fix[b + 1] = f;
    }
}

void crc_table(unsigned *t)
{
    // t[i] = x^(111-i) mod g, g = 0x1FFF409 (decoder.py:269)
    unsigned r = 1;
    for (int i = kMsgBits - 1; i >= 0; i--) {
        t[i] = r;
        r <<= 1;
        if (r & (1u << 24)) r ^= 0x1FFF409u;
    }
}

// `configured` = this context (one device, one sps, one CTA size) has already raised the kernel's shared-memory limits
template <int SPS, int FMT, int NT>
cudaError_t launch_scan_t(const ScanParams &P, unsigned n_blocks, cudaStream_t s, bool *configured)
{
    const size_t smem = tile_smem_bytes(Geo<SPS, NT>::T, P.rh, Geo<SPS, NT>::CW, NT / 32);
    if (!*configured) {
        cudaError_t e = cudaFuncSetAttribute(scan_slice_kernel<SPS, FMT, NT>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(scan_slice_kernel<SPS, FMT, NT>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                 cudaSharedmemCarveoutMaxShared);
        if (e != cudaSuccess) return e;
        *configured = true;
    }
    ADSB_LAUNCH((scan_slice_kernel<SPS, FMT, NT>), n_blocks, NT, smem, s, P);
    return cudaGetLastError();
}

template <int SPS, int NT>
cudaError_t launch_scan_f(const ScanParams &P, int fmt, unsigned n_blocks, cudaStream_t s, bool *configured)
{
    switch (fmt) {
    case kFmtC64: return launch_scan_t<SPS, kFmtC64, NT>(P, n_blocks, s, configured);
    case kFmtSC16: return launch_scan_t<SPS, kFmtSC16, NT>(P, n_blocks, s, configured);
    case kFmtCU8: return launch_scan_t<SPS, kFmtCU8, NT>(P, n_blocks, s, configured);
    default: return launch_scan_t<SPS, kFmtF32, NT>(P, n_blocks, s, configured);
    }
}

template <int NT>
cudaError_t launch_scan_s(int sps, const ScanParams &P, int fmt, unsigned n_blocks, cudaStream_t s, bool *configured)
{
    switch (sps) {
    case 2: return launch_scan_f<2, NT>(P, fmt, n_blocks, s, configured);
    case 4: return launch_scan_f<4, NT>(P, fmt, n_blocks, s, configured);
    case 8: return launch_scan_f<8, NT>(P, fmt, n_blocks, s, configured);
    default: return launch_scan_f<0, NT>(P, fmt, n_blocks, s, configured);
    }
}

cudaError_t launch_scan(adsb_ctx *c, const ScanParams &P0, int fmt, unsigned tile_first, unsigned n_blocks, cudaStream_t s)
{
    if (n_blocks == 0) return cudaSuccess;
    ScanParams P = P0;
    P.tile_first = tile_first;
    P.n_tiles_launch = n_blocks;
    // a launch that ends with the stream's final tile starts that tile early (see ScanParams::early_tile)
    P.early_tile = (c->early_tile && n_blocks >= 8 && (long long)tile_first + n_blocks == P.n_tiles_total) ? n_blocks - 1 : 0u;
    bool *configured = &c->scan_configured[fmt];
    return launch_scan_s<kThreads>(c->sps, P, fmt, n_blocks, s, configured);
}

long long tiles_for(const adsb_ctx *c, long long n_out, long long lead)
{
    if (n_out <= lead) return 0;
    return (n_out - lead + c->T - 1) / c->T;
}

// Validate the geometry shared by every fused entry point.
int check_geometry(adsb_ctx *ctx, size_t n_in, size_t lead, size_t n_out, size_t halo_left, bool shard = false)
{
    if (n_out == 0 || n_in == 0) return fail(ctx, ADSB_E_INVAL, "empty call");
    if (shard ? (lead + n_in < n_out + (size_t)ctx->hist) : (lead + n_in != n_out + (size_t)ctx->hist))
        return fail(ctx, ADSB_E_INVAL, "lead + n_in must equal n_out + 8*sps - 1");
    if (n_in > ctx->max_items) return fail(ctx, ADSB_E_INVAL, "n_in exceeds max_items given to adsb_create");
    if (lead != 0 && lead != (size_t)ctx->hist) return fail(ctx, ADSB_E_INVAL, "lead must be 0 or 8*sps - 1");
    if (halo_left != 0 && (halo_left < 256 || lead != 0)) return fail(ctx, ADSB_E_INVAL, "halo_left must be 0 or >= 256 with lead 0");
    return ADSB_OK;
}

ScanParams make_params(const adsb_ctx *c, const void *d_in, size_t n_in, size_t lead, size_t n_out, size_t halo_left,
                       float thr, long long base, int prev_bit, int do_demod)
{
    ScanParams P;
    memset(&P, 0, sizeof(P));
    P.in = d_in;
    P.n_arr = (long long)n_in;
    P.lead = (long long)lead;
    P.n_loc = (long long)(lead + n_in);
    P.win_hi = P.n_loc;        // no neighbours: in_left / in_right stay null
    P.own_lo = (long long)halo_left;
    P.own_hi = (long long)n_out;
    P.base = base;
    P.n_tiles_total = tiles_for(c, (long long)n_out, (long long)lead);
    P.thr = thr;
    P.sps = c->sps;
    P.rh = c->rh;
    P.left_is_start = halo_left == 0;
    P.prev_bit = prev_bit;
    P.do_demod = do_demod;
    P.hist = c->hist;
    P.cand = c->d_cand;
    P.cand_cap = c->cand_cap;
    P.tile_tab = c->d_tile_tab;
    P.tile_chunk_tot = c->d_tile_chunk_tot;
    P.hdr = c->d_hdr;
    P.crc_tab = c->d_crc;
    P.prefetch_tiles = c->prefetch_tiles;
    // tiles t with: t >= 1, (t+1)*T + rh <= n_arr, t*T + lead >= own_lo, (t+1)*T + lead <= own_hi
    {
        const long long T = c->T;
        long long lo = 1;
        const long long lo2 = (P.own_lo - P.lead + T - 1) / T;
        if (lo2 > lo) lo = lo2;
        long long hi = (P.n_arr - c->rh) / T - 1;
        const long long hi2 = (P.own_hi - P.lead) / T - 1;
        if (hi2 < hi) hi = hi2;
        if (hi >= lo) {
            P.fast_lo = (unsigned)lo;
            P.fast_span = (unsigned)(hi - lo);
        } else {
            P.fast_lo = 0;
            P.fast_span = 0xffffffffu;
        }
    }
    return P;
}

int launch_resolve(adsb_ctx *ctx, long long n_tiles, long long base, long long eob_in, cudaStream_t s, bool rerun = false,
                   bool dev_carry = false, bool reset_done = false)
{
    ResolveParams R;
    memset(&R, 0, sizeof(R));
    R.tile_tab = ctx->d_tile_tab;
    R.n_tiles = n_tiles;
    R.cand = ctx->d_cand;
    R.ord = ctx->d_ord;
    R.hits = ctx->d_hits;
    R.hit_cap = ctx->hit_cap;
    R.cand_cap = ctx->cand_cap;
    R.tile_chunk_tot = ctx->d_tile_chunk_tot;
    R.cand_chunk_tot = ctx->d_cand_chunk_tot;
    R.hdr = ctx->d_hdr;
    R.base = base;
    R.eob_in = eob_in;
    R.blank = 63 * ctx->sps;
    R.dev_carry = dev_carry ? 1 : 0;
    if (n_tiles <= 0) return ADSB_OK;
    const long long gchunks = (n_tiles + kTilesPerChunk - 1) / kTilesPerChunk;
    const unsigned cap_blocks = (unsigned)ctx->num_sms * 4u;
    const unsigned gblocks = (unsigned)(gchunks < cap_blocks ? gchunks : cap_blocks);
    const unsigned cblocks = cap_blocks;
    if (rerun) {
        if (!reset_done) ADSB_LAUNCH(reset_resolve_kernel, 1, 1, 0, s, ctx->d_hdr, R.dev_carry);
    } else {
        ADSB_LAUNCH(gather_kernel, gblocks, kThreads, 0, s, R);
    }
    ADSB_LAUNCH(chain_kernel, cap_blocks, kThreads, 0, s, R);
    ADSB_LAUNCH(compact_kernel, cblocks, kThreads, 0, s, R);
    if (ctx->fec_mode == ADSB_FEC_CONSERVATIVE) {
        FecParams F;
        F.hits = ctx->d_hits;
        F.hdr = ctx->d_hdr;
        F.hit_cap = ctx->hit_cap;
        F.t25 = ctx->d_fec_t25;
        F.keys = ctx->d_fec_keys;
        F.fix = ctx->d_fec_fix;
        ADSB_LAUNCH(fec_kernel, cap_blocks, kThreads, 0, s, F);
    }
    CK(cudaGetLastError());
    return ADSB_OK;
}

// zero the accepted-per-chunk totals chain_kernel accumulates (bounded by the candidates n_out items can hold)
int zero_cand_totals(adsb_ctx *ctx, long long n_out, cudaStream_t s)
{
    size_t n = (size_t)(n_out / 2 / kCandPerChunk) + 2;
    if (n > ctx->cand_chunk_n) n = ctx->cand_chunk_n;
    CK(cudaMemsetAsync(ctx->d_cand_chunk_tot, 0, n * sizeof(unsigned), s));
    return ADSB_OK;
}

int begin_call(adsb_ctx *ctx, long long n_tiles, long long n_out, cudaStream_t s)
{
    ctx->small.clean = false;               // a throughput-path call dirties the header (small_call sets it again itself)
    ctx->small.last_was_small = false;
    ADSB_LAUNCH(init_header_kernel, 1, 1, 0, s, ctx->d_hdr);
    CK(cudaGetLastError());
    size_t n = (size_t)(n_tiles / kTilesPerChunk) + 2;
    if (n > ctx->tile_chunk_n) n = ctx->tile_chunk_n;
    CK(cudaMemsetAsync(ctx->d_tile_chunk_tot, 0, n * sizeof(unsigned), s));
    return zero_cand_totals(ctx, n_out, s);
}

int ensure_input(adsb_ctx *ctx, size_t bytes)
{
    if (ctx->d_in_bytes >= bytes) return ADSB_OK;
    if (ctx->d_in) cudaFree(ctx->d_in);
    ctx->d_in = nullptr;
    ctx->d_in_bytes = 0;
    size_t want = ctx->max_items * 8 + 256;
    if (want < bytes) want = bytes;
    cudaError_t e = cudaMalloc(&ctx->d_in, want);
    if (e != cudaSuccess) {
        want = bytes;
        e = cudaMalloc(&ctx->d_in, want);
    }
    if (e != cudaSuccess) return fail(ctx, ADSB_E_NOMEM, "cudaMalloc(input staging)", e);
    ctx->d_in_bytes = want;
    return ADSB_OK;
}

// framer.py:121-123,165,177-179 -- advance prev_eob_idx from the call's summary
long long next_eob(const Header &h, long long eob_in, long long n_out, int sps)
{
    if (!((h.flags & kFlagRise) && (h.flags & kFlagFall))) return eob_in;   // framer.py:95 guard
    long long e = (h.last_accept >= 0) ? h.last_accept + 63LL * sps : eob_in;
    if (h.last_pulse >= 0 && h.last_pulse > e) e = -1;    // a later pulse reset it (framer.py:121-123)
    if (e >= n_out) e -= n_out;                            // framer.py:177-179
    return e;
}

int status_of(adsb_ctx *ctx, const Header &h)
{
    if (h.flags & kFlagPeerTimeout) return fail(ctx, ADSB_E_PEER, "a neighbouring rank did not reach the step (peer wait timed out)");
    if (h.flags & kFlagHalo) return fail(ctx, ADSB_E_HALO, "pulse run crosses the left halo edge");
    if (h.flags & (kFlagCandOverflow | kFlagHitOverflow)) return fail(ctx, ADSB_E_NOMEM, "candidate capacity exceeded (raise max_items)");
    return ADSB_OK;
}


// ---- latency path ------------------------------------------------------------------------------------------
double host_now()
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

template <int SPS, int FMT>
cudaError_t launch_small_t(const ScanParams &P, const SmallParams &Q, cudaStream_t s, bool *configured)
{
    const size_t smem = tile_smem_bytes(Geo<SPS, kThreads>::T, P.rh, Geo<SPS, kThreads>::CW, kWarps);
    if (!*configured) {
        cudaError_t e = cudaFuncSetAttribute(small_work_kernel<SPS, FMT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        *configured = true;
    }
    ADSB_LAUNCH((small_work_kernel<SPS, FMT>), (unsigned)Q.n_tiles, kThreads, smem, s, P, Q);
    return cudaGetLastError();
}

cudaError_t launch_small(adsb_ctx *c, const ScanParams &P, const SmallParams &Q, int fmt, cudaStream_t s)
{
    bool *cf = &c->small_configured[fmt == kFmtC64 ? 1 : 0];
    if (fmt == kFmtC64) {
        switch (c->sps) {
        case 2: return launch_small_t<2, kFmtC64>(P, Q, s, cf);
        case 4: return launch_small_t<4, kFmtC64>(P, Q, s, cf);
        case 8: return launch_small_t<8, kFmtC64>(P, Q, s, cf);
        default: return launch_small_t<0, kFmtC64>(P, Q, s, cf);
        }
    }
    switch (c->sps) {
    case 2: return launch_small_t<2, kFmtF32>(P, Q, s, cf);
    case 4: return launch_small_t<4, kFmtF32>(P, Q, s, cf);
    case 8: return launch_small_t<8, kFmtF32>(P, Q, s, cf);
    default: return launch_small_t<0, kFmtF32>(P, Q, s, cf);
    }
}

template <class T>
int pinned_grow(adsb_ctx *ctx, T **p, size_t *have, size_t want_items, const char *what)
{
    if (*have >= want_items) return ADSB_OK;
    if (*p) cudaFreeHost(*p);
    *p = nullptr;
    *have = 0;
    cudaError_t e = cudaHostAlloc((void **)p, want_items * sizeof(T), cudaHostAllocMapped);
    if (e != cudaSuccess) return fail(ctx, ADSB_E_NOMEM, what, e);
    *have = want_items;
    return ADSB_OK;
}

// spin on a completion stamp the kernel writes into pinned host memory (no cudaStreamSynchronize on this path); the
// stream is polled now and then so that a failed launch ends the wait
int wait_stamp(adsb_ctx *ctx, const unsigned long long *stamp, unsigned long long seq, cudaStream_t s)
{
    const volatile unsigned long long *v = stamp;
    for (unsigned long long spins = 0;; spins++) {
        if (*v == seq) return ADSB_OK;
        if ((spins & 0x3fff) == 0x3fff) {
            const cudaError_t q = cudaStreamQuery(s);
            if (q == cudaSuccess) return (*v == seq) ? ADSB_OK : fail(ctx, ADSB_E_INTERNAL, "latency path: kernel ended without its completion stamp");
            if (q != cudaErrorNotReady) return fail(ctx, ADSB_E_CUDA, "latency path kernel", q);
        }
    }
}

bool small_eligible(const adsb_ctx *ctx, int fmt, size_t n_in, size_t lead, size_t n_out)
{
    if (ctx->fec_mode != ADSB_FEC_NONE || ctx->profile || ctx->no_small) return false;
    if (fmt != kFmtF32 && fmt != kFmtC64) return false;
    return tiles_for(ctx, (long long)n_out, (long long)lead) <= kSmallTiles && tiles_for(ctx, (long long)n_out, (long long)lead) >= 1 &&
           n_in <= (size_t)kSmallTiles * ctx->T + ctx->rh + ctx->hist;
}

// One work() call on the latency path.  Same results and state updates as the throughput path (adsb_burst_work ->
// adsb_burst_finish); hits land in ctx->small.out->hits.
int small_call(adsb_ctx *ctx, const void *in, int fmt, size_t n_in, size_t lead, size_t n_out, float threshold, long long base,
               int do_demod, size_t *n_hits)
{
    int rc;
    const size_t isz = fmt == kFmtC64 ? 8 : 4;
    cudaStream_t s = ctx->stream;
    if (!ctx->small.out) {
        cudaError_t e = cudaHostAlloc((void **)&ctx->small.out, sizeof(SmallOut), cudaHostAllocMapped);
        if (e != cudaSuccess) return fail(ctx, ADSB_E_NOMEM, "pinned result block", e);
        memset(ctx->small.out, 0, sizeof(SmallOut));
    }
    {
        size_t have = ctx->small.in_bytes;
        unsigned char *p = static_cast<unsigned char *>(ctx->small.in);
        if ((rc = pinned_grow(ctx, &p, &have, ((size_t)kSmallTiles * ctx->T + ctx->rh + ctx->hist + 64) * 8, "pinned input staging"))) return rc;
        ctx->small.in = p;
        ctx->small.in_bytes = have;
    }
    const bool want_mag = fmt == kFmtC64;
    if (want_mag && (rc = pinned_grow(ctx, &ctx->small.mag, &ctx->small.mag_items, (size_t)kSmallTiles * ctx->T + ctx->rh + ctx->hist + 64,
                                      "pinned |IQ|^2 staging")))
        return rc;
    const long long n_tiles = tiles_for(ctx, (long long)n_out, (long long)lead);
    if (!ctx->small.clean) {        // a throughput-path call came in between: one header reset puts the device back in step
        if ((rc = begin_call(ctx, n_tiles, (long long)n_out, s))) return rc;
        ctx->small.clean = true;
    }
    const double th0 = host_now();
    memcpy(ctx->small.in, in, n_in * isz);      // GNU Radio's buffers are pageable: stage through pinned memory
    const double th1 = host_now();
    const int prev_bit = (lead != 0) ? (0.0f >= threshold) : (ctx->prev_in0 >= threshold);
    const long long eob_in = ctx->prev_eob;
    // up to 128 KiB the kernel reads the pinned staging buffer in place (one PCIe round trip, no copy call); beyond that
    // the copy engine moves it at twice the bandwidth zero-copy reads reach and the kernel reads device memory
    const void *src = ctx->small.in;
    if (n_in * isz > (128u << 10)) {
        if ((rc = ensure_input(ctx, n_in * isz))) return rc;
        CK(cudaMemcpyAsync(ctx->d_in, ctx->small.in, n_in * isz, cudaMemcpyHostToDevice, s));
        src = ctx->d_in;
    }
    ScanParams P = make_params(ctx, src, n_in, lead, n_out, 0, threshold, base, prev_bit, do_demod);
    P.prefetch_tiles = 0;
    SmallParams Q;
    memset(&Q, 0, sizeof(Q));
    Q.out = ctx->small.out;
    Q.seq = ++ctx->small.seq;
    Q.eob_in = eob_in;
    Q.blank = 63 * ctx->sps;
    Q.n_tiles = (int)n_tiles;
    if (want_mag) {
        Q.mag_out = ctx->small.mag;
        Q.mag_first = (long long)lead;          // window index of in[0]
        Q.mag_n = (long long)n_in;
    }
    CK(launch_small(ctx, P, Q, fmt, s));
    const double th2 = host_now();
    ctx->small.calls++;
    if ((rc = wait_stamp(ctx, &ctx->small.out->seq, Q.seq, s))) {
        ctx->small.clean = false;
        return rc;
    }
    ctx->small.host_t[0] = th1 - th0;
    ctx->small.host_t[1] = th2 - th1;
    ctx->small.host_t[2] = host_now() - th2;
    ctx->staged_fmt = -1;
    ctx->small.last_was_small = true;
    ctx->staged_n = n_in;
    ctx->small_fmt = fmt;
    ctx->last.pending = true;
    ctx->last.n_out = (long long)n_out;
    ctx->last.lead = (long long)lead;
    ctx->last.base = base;
    ctx->last.eob_in = eob_in;
    ctx->last.left_is_start = 1;
    ctx->last.stream = s;
    Header h = ctx->small.out->hdr;
    if (h.flags & kFlagSmallOverflow) {
        // more candidates than the in-kernel resolve holds: the resolve kernels finish the call from the same records
        ctx->small.fallbacks++;
        ctx->small.clean = false;
        ctx->small.use_device_hits = true;
        if ((rc = zero_cand_totals(ctx, (long long)n_out, s))) return rc;
        if ((rc = launch_resolve(ctx, n_tiles, base, eob_in, s))) return rc;
        return adsb_burst_finish(ctx, nullptr, nullptr, 0, n_hits);
    }
    ctx->small.use_device_hits = false;
    ctx->last.pending = false;
    rc = status_of(ctx, h);
    if (rc) {
        ctx->prev_in0 = ((long long)n_out - 1 >= (long long)lead) ? h.x_last : 0.0f;
        return rc;
    }
    adsb_counters &c = ctx->counters;
    c.n_samples = (uint64_t)n_out;
    c.n_candidates = h.n_cand;
    c.n_hits = h.n_hits;
    c.n_demod = h.n_demod;
    c.n_crc_ok = h.n_crc;
    c.last_pulse = h.last_pulse;
    c.last_accept = h.last_accept;
    c.first_candidate = h.first_cand;
    c.carry_unstable = 0;
    c.n_fec_fixed = 0;
    c.any_rise = (h.flags & kFlagRise) ? 1u : 0u;
    c.any_fall = (h.flags & kFlagFall) ? 1u : 0u;
    ctx->prev_eob = next_eob(h, eob_in, (long long)n_out, ctx->sps);      // framer.py:121-123,165,177-179
    ctx->prev_in0 = ((long long)n_out - 1 >= (long long)lead) ? h.x_last : 0.0f;   // framer.py:87
    *n_hits = (size_t)h.n_hits;
    return ADSB_OK;
}

// where the hit records of the latest small_call are (pinned result block, or the device after a fallback)
int small_fetch_hits(adsb_ctx *ctx, Rec *dst, size_t n)
{
    if (!n) return ADSB_OK;
    if (!ctx->small.use_device_hits) {
        memcpy(dst, ctx->small.out->hits, n * sizeof(Rec));
        return ADSB_OK;
    }
    CK(cudaMemcpy(dst, ctx->d_hits, n * sizeof(Rec), cudaMemcpyDeviceToHost));
    return ADSB_OK;
}

}  // namespace

// ---------------------------------------------------------------------------------------------------
extern "C" {

const char *adsb_strerror(int code)
{
    switch (code) {
    case ADSB_OK: return "ok";
    case ADSB_E_INVAL: return "invalid argument";
    case ADSB_E_CUDA: return "CUDA runtime error";
    case ADSB_E_NOMEM: return "out of memory / capacity";
    case ADSB_E_CAPACITY: return "output capacity too small";
    case ADSB_E_NODEVICE: return "no usable sm_100 CUDA device";
    case ADSB_E_HALO: return "pulse run crosses the left halo of a shard";
    case ADSB_E_PEER: return "multi-GPU peer exchange failed";
    case ADSB_E_INTERNAL: return "internal error";
    default: return "unknown error";
    }
}

int adsb_abi_version(void) { return ADSB_B200_ABI_VERSION; }

int adsb_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

const char *adsb_last_error(const adsb_ctx *ctx) { return ctx ? ctx->err : "null context"; }

void adsb_destroy(adsb_ctx *ctx)
{
    if (!ctx) return;
    DeviceGuard guard(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    adsb_shard_detach(ctx);
    if (ctx->shard.own_base) cudaFree(ctx->shard.own_base);
    cudaFree(ctx->d_in);
    cudaFree(ctx->d_cand);
    cudaFree(ctx->d_ord);
    cudaFree(ctx->d_hits);
    cudaFree(ctx->d_tile_tab);
    cudaFree(ctx->d_tile_chunk_tot);
    cudaFree(ctx->d_cand_chunk_tot);
    cudaFree(ctx->d_hdr);
    cudaFree(ctx->d_crc);
    cudaFree(ctx->d_fec_t25);
    cudaFree(ctx->d_fec_keys);
    cudaFree(ctx->d_fec_fix);
    cudaFree(ctx->d_sob);
    cudaFree(ctx->d_bits);
    cudaFree(ctx->d_valid);
    cudaFree(ctx->d_mag);
    adsb_stream_close(ctx);
    if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
    for (int i = 0; i < 2; i++) {
        if (ctx->h_ring[i]) cudaFreeHost(ctx->h_ring[i]);
        if (ctx->ev_ring[i]) cudaEventDestroy(ctx->ev_ring[i]);
    }
    if (ctx->small.out) cudaFreeHost(ctx->small.out);
    if (ctx->small.in) cudaFreeHost(ctx->small.in);
    if (ctx->small.mag) cudaFreeHost(ctx->small.mag);
    if (ctx->small.sob) cudaFreeHost(ctx->small.sob);
    if (ctx->small.bits) cudaFreeHost(ctx->small.bits);
    if (ctx->small.valid) cudaFreeHost(ctx->small.valid);
    cudaFree(ctx->small.d_done);
    free(ctx->h_tmp);
    free(ctx->h_bits);
    for (int i = 0; i < kEvRing; i++)
        if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
    for (int i = 0; i < 3; i++)
        if (ctx->pev[i]) cudaEventDestroy(ctx->pev[i]);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    delete ctx;
}

int adsb_create(adsb_ctx **out, int device, int sps, size_t max_items)
{
    if (!out) return ADSB_E_INVAL;
    *out = nullptr;
    if (sps < 2 || sps > 32 || (sps & 1) || max_items == 0) return ADSB_E_INVAL;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return ADSB_E_NODEVICE;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess || prop.major < 10) return ADSB_E_NODEVICE;
    DeviceGuard guard(device);
    if (guard.status != cudaSuccess) return ADSB_E_NODEVICE;

    adsb_ctx *ctx = new (std::nothrow) adsb_ctx();
    if (!ctx) return ADSB_E_NOMEM;
    memset(ctx, 0, sizeof(*ctx));
    ctx->device = device;
    ctx->sps = sps;
    ctx->hist = 8 * sps - 1;
    ctx->nt = kThreads;
    ctx->T = tile_items(sps, ctx->nt);
    ctx->rh = right_halo(sps, ctx->nt);
    ctx->max_items = max_items;
    ctx->prev_in0 = 0.0f;
    ctx->prev_eob = -1;
    ctx->staged_fmt = -1;
    ctx->num_sms = prop.multiProcessorCount;
    {
        const char *e = getenv("ADSB_PREFETCH_TILES");      // tuning knob; default: two tiles per SM ahead
        ctx->prefetch_tiles = e ? (unsigned)atoi(e) : (unsigned)(ctx->num_sms * 2);
    }
    ctx->err[0] = 0;
    ctx->small.clean = false;
    {
        const char *e = getenv("ADSB_NO_LATENCY_PATH");
        ctx->no_small = (e && e[0] == '1') ? 1 : 0;
        e = getenv("ADSB_EARLY_TILE");
        ctx->early_tile = (e && e[0] == '0') ? 0 : 1;
    }

    auto die = [&](int code, const char *what, cudaError_t e) {
        char msg[256];
        snprintf(msg, sizeof(msg), "%s: %s", what, cudaGetErrorString(e));
        adsb_destroy(ctx);
        fprintf(stderr, "adsb_create: %s\n", msg);
        return code;
    };
    cudaError_t e;
    if ((e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess) return die(ADSB_E_CUDA, "stream", e);
    if ((e = cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking)) != cudaSuccess) return die(ADSB_E_CUDA, "stream", e);
    for (int i = 0; i < kEvRing; i++)
        if ((e = cudaEventCreateWithFlags(&ctx->ev[i], cudaEventDisableTiming)) != cudaSuccess) return die(ADSB_E_CUDA, "event", e);
    for (int i = 0; i < 3; i++)
        if ((e = cudaEventCreate(&ctx->pev[i])) != cudaSuccess) return die(ADSB_E_CUDA, "event", e);

    // candidates: at most one per 2 samples in theory, ~1 per 600 on busy air; size for 1 per 16.
    size_t cap = max_items / 16 + 4096;
    {
        const size_t blocks = (max_items / (size_t)ctx->T + 4) * (size_t)kTileSlots;   // fixed per-tile blocks
        if (cap < 2 * blocks) cap = 2 * blocks;
    }
    if (cap > 0x7fffff00u) cap = 0x7fffff00u;
    ctx->cand_cap = (unsigned)cap;
    ctx->hit_cap = (unsigned)cap;
    ctx->tile_cap = max_items / ctx->T + 4;
    ctx->tile_chunk_n = ctx->tile_cap / kTilesPerChunk + 4;
    ctx->cand_chunk_n = cap / kCandPerChunk + 4;
    ctx->tag_cap = 4096;
#define ALLOC(ptr, bytes)                                                                       \
    if ((e = cudaMalloc((void **)&(ptr), (bytes))) != cudaSuccess) return die(ADSB_E_NOMEM, #ptr, e)
    ALLOC(ctx->d_cand, cap * sizeof(Rec));
    ALLOC(ctx->d_ord, cap * sizeof(Rec));
    ALLOC(ctx->d_hits, cap * sizeof(Rec));
    ALLOC(ctx->d_tile_tab, ctx->tile_cap * sizeof(uint2));
    ALLOC(ctx->d_tile_chunk_tot, ctx->tile_chunk_n * sizeof(unsigned));
    ALLOC(ctx->d_cand_chunk_tot, ctx->cand_chunk_n * sizeof(unsigned));
    ALLOC(ctx->d_hdr, sizeof(Header));
    ALLOC(ctx->d_crc, kMsgBits * sizeof(unsigned));
    ALLOC(ctx->d_fec_t25, kMsgBits * sizeof(unsigned));
    ALLOC(ctx->d_fec_keys, kFecPatterns * sizeof(unsigned));
    ALLOC(ctx->d_fec_fix, kFecPatterns * sizeof(unsigned short));
    ALLOC(ctx->d_sob, ctx->tag_cap * sizeof(long long));
    ALLOC(ctx->d_bits, ctx->tag_cap * kMsgBits);
    ALLOC(ctx->d_valid, ctx->tag_cap);
#undef ALLOC
    if ((e = cudaMallocHost((void **)&ctx->h_stage, sizeof(HostStage))) != cudaSuccess) return die(ADSB_E_NOMEM, "pinned stage", e);
    // the header's sticky carry_unstable counter survives init_header_kernel: start it from zero
    if ((e = cudaMemset(ctx->d_hdr, 0, sizeof(Header))) != cudaSuccess) return die(ADSB_E_CUDA, "memset(header)", e);
    unsigned tab[kMsgBits];
    crc_table(tab);
    if ((e = cudaMemcpy(ctx->d_crc, tab, sizeof(tab), cudaMemcpyHostToDevice)) != cudaSuccess) return die(ADSB_E_CUDA, "crc table", e);
    {
        unsigned t25[kMsgBits], keys[kFecPatterns];
        unsigned short fix[kFecPatterns];
        fec_tables(t25, keys, fix);
        if ((e = cudaMemcpy(ctx->d_fec_t25, t25, sizeof(t25), cudaMemcpyHostToDevice)) != cudaSuccess ||
            (e = cudaMemcpy(ctx->d_fec_keys, keys, sizeof(keys), cudaMemcpyHostToDevice)) != cudaSuccess ||
            (e = cudaMemcpy(ctx->d_fec_fix, fix, sizeof(fix), cudaMemcpyHostToDevice)) != cudaSuccess)
            return die(ADSB_E_CUDA, "fec tables", e);
    }
    *out = ctx;
    return ADSB_OK;
}

int adsb_reset_state(adsb_ctx *ctx)
{
    if (!ctx) return ADSB_E_INVAL;
    ctx->prev_in0 = 0.0f;
    ctx->prev_eob = -1;
    return ADSB_OK;
}

int adsb_get_state(const adsb_ctx *ctx, float *prev_in0, int64_t *prev_eob_idx)
{
    if (!ctx) return ADSB_E_INVAL;
    if (prev_in0) *prev_in0 = ctx->prev_in0;
    if (prev_eob_idx) *prev_eob_idx = ctx->prev_eob;
    return ADSB_OK;
}

int adsb_set_state(adsb_ctx *ctx, float prev_in0, int64_t prev_eob_idx)
{
    if (!ctx) return ADSB_E_INVAL;
    ctx->prev_in0 = prev_in0;
    ctx->prev_eob = prev_eob_idx;
    return ADSB_OK;
}

int adsb_set_error_correction(adsb_ctx *ctx, int mode)
{
    if (!ctx || (mode != ADSB_FEC_NONE && mode != ADSB_FEC_CONSERVATIVE)) return ADSB_E_INVAL;
    ctx->fec_mode = mode;
    return ADSB_OK;
}

int adsb_get_counters(const adsb_ctx *ctx, adsb_counters *out)
{
    if (!ctx || !out) return ADSB_E_INVAL;
    *out = ctx->counters;
    return ADSB_OK;
}

int adsb_latency_stats(const adsb_ctx *ctx, uint64_t *calls, uint64_t *fallbacks)
{
    if (!ctx) return ADSB_E_INVAL;
    if (calls) *calls = ctx->small.calls;
    if (fallbacks) *fallbacks = ctx->small.fallbacks;
    return ADSB_OK;
}

/* A/B builds (-DADSB_SMALL_TIMING=1): device timestamps of the latest latency-path call + host timestamps around it */
int adsb_latency_trace(const adsb_ctx *ctx, uint64_t *dev8, double *host3)
{
    if (!ctx || !ctx->small.out) return ADSB_E_INVAL;
    if (dev8) memcpy(dev8, ctx->small.out->t, sizeof(ctx->small.out->t));
    if (host3) memcpy(host3, ctx->small.host_t, sizeof(ctx->small.host_t));
    return ADSB_OK;
}

int adsb_set_latency_path(adsb_ctx *ctx, int on)
{
    if (!ctx) return ADSB_E_INVAL;
    ctx->no_small = on ? 0 : 1;
    return ADSB_OK;
}

int adsb_profile_enable(adsb_ctx *ctx, int on)
{
    if (!ctx) return ADSB_E_INVAL;
    ctx->profile = on ? 1 : 0;
    return ADSB_OK;
}

int adsb_profile_last(adsb_ctx *ctx, float *scan_ms, float *resolve_ms)
{
    if (!ctx || !ctx->profile) return ADSB_E_INVAL;
    CK(cudaEventSynchronize(ctx->pev[2]));
    float a = 0.f, b = 0.f;
    CK(cudaEventElapsedTime(&a, ctx->pev[0], ctx->pev[1]));
    CK(cudaEventElapsedTime(&b, ctx->pev[1], ctx->pev[2]));
    if (scan_ms) *scan_ms = a;
    if (resolve_ms) *resolve_ms = b;
    return ADSB_OK;
}

int adsb_burst_work_dev(adsb_ctx *ctx, const void *d_in, int is_c64, size_t n_in, size_t lead, size_t n_out,
                        size_t halo_left, float threshold, int64_t base, int64_t eob_in, void *stream,
                        const adsb_hit **d_hits)
{
    ADSB_RANGE("adsb_burst_work_dev");
    if (!ctx || !d_in || is_c64 < 0 || is_c64 >= kNumFmt) return ADSB_E_INVAL;
    int rc = check_geometry(ctx, n_in, lead, n_out, halo_left, true);
    if (rc) return rc;
    ON_DEVICE(ctx);
    cudaStream_t s = stream ? (cudaStream_t)stream : ctx->stream;
    const int prev_bit = (lead != 0) ? (0.0f >= threshold) : (ctx->prev_in0 >= threshold);
    const long long n_tiles = tiles_for(ctx, (long long)n_out, (long long)lead);
    if ((size_t)n_tiles > ctx->tile_cap) return fail(ctx, ADSB_E_INVAL, "too many tiles for this context");
    if ((rc = begin_call(ctx, n_tiles, (long long)n_out, s))) return rc;
    ScanParams P = make_params(ctx, d_in, n_in, lead, n_out, halo_left, threshold, base, prev_bit, 1);
    if (ctx->profile) CK(cudaEventRecord(ctx->pev[0], s));
    CK(launch_scan(ctx, P, is_c64, 0u, (unsigned)n_tiles, s));
    if (ctx->profile) CK(cudaEventRecord(ctx->pev[1], s));
    if ((rc = launch_resolve(ctx, n_tiles, base, eob_in, s))) return rc;
    if (ctx->profile) CK(cudaEventRecord(ctx->pev[2], s));
    ctx->last.pending = true;
    ctx->last.n_out = (long long)n_out;
    ctx->last.lead = (long long)lead;
    ctx->last.base = base;
    ctx->last.eob_in = eob_in;
    ctx->last.left_is_start = halo_left == 0;
    ctx->last.stream = s;
    if (d_hits) *d_hits = reinterpret_cast<const adsb_hit *>(ctx->d_hits);
    return ADSB_OK;
}

int adsb_burst_reresolve_dev(adsb_ctx *ctx, int64_t eob_in, void *stream)
{
    if (!ctx || !ctx->last.pending) return ADSB_E_INVAL;
    ON_DEVICE(ctx);
    cudaStream_t s = stream ? (cudaStream_t)stream : ctx->last.stream;
    const long long n_tiles = tiles_for(ctx, ctx->last.n_out, ctx->last.lead);
    ctx->rerun_count++;
    int rc = zero_cand_totals(ctx, ctx->last.n_out, s);
    if (rc) return rc;
    rc = launch_resolve(ctx, n_tiles, ctx->last.base, eob_in, s, true);
    if (rc) return rc;
    ctx->last.eob_in = eob_in;
    ctx->last.stream = s;
    return ADSB_OK;
}

int adsb_burst_carry_round_dev(adsb_ctx *ctx, const int64_t *d_gathered, int world, int rank, int final_check,
                               void *stream)
{
    if (!ctx || !ctx->last.pending || !d_gathered || world < 1 || rank < 0 || rank >= world) return ADSB_E_INVAL;
    ON_DEVICE(ctx);
    cudaStream_t s = stream ? (cudaStream_t)stream : ctx->last.stream;
    ADSB_LAUNCH(carry_plan_kernel, 1, 1, 0, s, ctx->d_hdr, reinterpret_cast<const long long *>(d_gathered), world, rank,
                                      63LL * ctx->sps, final_check);
    CK(cudaGetLastError());
    if (final_check) return ADSB_OK;
    const long long n_tiles = tiles_for(ctx, ctx->last.n_out, ctx->last.lead);
    int rc = zero_cand_totals(ctx, ctx->last.n_out, s);
    if (rc) return rc;
    return launch_resolve(ctx, n_tiles, ctx->last.base, 0, s, true, true);   // no-ops unless hdr->carry_need
}

int adsb_burst_export_carry_dev(adsb_ctx *ctx, int64_t used_global, int64_t *d_out3, void *stream)
{
    if (!ctx || !ctx->last.pending || !d_out3) return ADSB_E_INVAL;
    ON_DEVICE(ctx);
    cudaStream_t s = stream ? (cudaStream_t)stream : ctx->last.stream;
    ADSB_LAUNCH(export_carry_kernel, 1, 1, 0, s, ctx->d_hdr, ctx->last.base, used_global, used_global == ADSB_CARRY_USED_ON_DEVICE,
                                        reinterpret_cast<long long *>(d_out3));
    CK(cudaGetLastError());
    return ADSB_OK;
}

// ---- multi-GPU peer mode -----------------------------------------------------------------------------------
int adsb_shard_alloc(adsb_ctx *ctx, size_t sample_bytes, void **d_samples, void *ipc_handle)
{
    if (!ctx || !d_samples || sample_bytes == 0) return ADSB_E_INVAL;
    ON_DEVICE(ctx);
    if (ctx->shard.world) return fail(ctx, ADSB_E_INVAL, "adsb_shard_alloc: detach first");
    if (ctx->shard.own_base) {
        cudaFree(ctx->shard.own_base);
        ctx->shard.own_base = nullptr;
    }
    const size_t bytes = ADSB_SHARD_CTRL_BYTES + ((sample_bytes + 255) & ~(size_t)255);
    void *p = nullptr;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) return fail(ctx, ADSB_E_NOMEM, "cudaMalloc(shard)", e);
    CK(cudaMemset(p, 0, ADSB_SHARD_CTRL_BYTES));
    ctx->shard.own_base = p;
    ctx->shard.own_bytes = bytes;
    *d_samples = static_cast<char *>(p) + ADSB_SHARD_CTRL_BYTES;
    if (ipc_handle) {
#if defined(ADSB_HOST_EMU)
        memset(ipc_handle, 0, ADSB_IPC_HANDLE_BYTES);
#else
        static_assert(sizeof(cudaIpcMemHandle_t) == ADSB_IPC_HANDLE_BYTES, "IPC handle size");
        cudaIpcMemHandle_t h;
        e = cudaIpcGetMemHandle(&h, p);
        if (e != cudaSuccess) return fail(ctx, ADSB_E_PEER, "cudaIpcGetMemHandle", e);
        memcpy(ipc_handle, &h, sizeof(h));
#endif
    }
    return ADSB_OK;
}

int adsb_shard_write(adsb_ctx *ctx, size_t byte_offset, const void *src, size_t bytes, void *stream)
{
    if (!ctx || !src) return ADSB_E_INVAL;
    if (!ctx->shard.own_base || byte_offset + bytes > ctx->shard.own_bytes - ADSB_SHARD_CTRL_BYTES)
        return fail(ctx, ADSB_E_INVAL, "adsb_shard_write: outside the allocation of adsb_shard_alloc");
    ON_DEVICE(ctx);
    cudaStream_t s = stream ? (cudaStream_t)stream : ctx->stream;
    CK(cudaMemcpyAsync(static_cast<char *>(ctx->shard.own_base) + ADSB_SHARD_CTRL_BYTES + byte_offset, src, bytes,
                       cudaMemcpyDefault, s));
    if (!stream) CK(cudaStreamSynchronize(s));
    return ADSB_OK;
}

int adsb_shard_detach(adsb_ctx *ctx)
{
    if (!ctx) return ADSB_E_INVAL;
    DeviceGuard guard(ctx->device);
#if !defined(ADSB_HOST_EMU)
    if (ctx->shard.exec) cudaGraphExecDestroy(ctx->shard.exec);
    if (ctx->shard.graph) cudaGraphDestroy(ctx->shard.graph);
    ctx->shard.exec = nullptr;
    ctx->shard.graph = nullptr;
    for (int r = 0; r < ctx->shard.world; r++)
        if (ctx->shard.ipc_opened[r]) cudaIpcCloseMemHandle(ctx->shard.peer_base[r]);
#endif
    for (int r = 0; r < kMaxWorld; r++) {
        ctx->shard.peer_base[r] = nullptr;
        ctx->shard.ipc_opened[r] = false;
    }
    ctx->shard.world = 0;
    return ADSB_OK;
}

int adsb_shard_attach_ptrs(adsb_ctx *ctx, int rank, int world, void *const *peer_bases)
{
    if (!ctx || !peer_bases || world < 1 || world > kMaxWorld || rank < 0 || rank >= world) return ADSB_E_INVAL;
    adsb_shard_detach(ctx);
    for (int r = 0; r < world; r++) {
        if (!peer_bases[r]) return fail(ctx, ADSB_E_INVAL, "adsb_shard_attach_ptrs: null base pointer");
        ctx->shard.peer_base[r] = peer_bases[r];
    }
    ctx->shard.rank = rank;
    ctx->shard.world = world;
    return ADSB_OK;
}

int adsb_shard_attach(adsb_ctx *ctx, int rank, int world, const void *ipc_handles)
{
    if (!ctx || !ipc_handles || world < 1 || world > kMaxWorld || rank < 0 || rank >= world) return ADSB_E_INVAL;
    if (!ctx->shard.own_base) return fail(ctx, ADSB_E_INVAL, "adsb_shard_attach: call adsb_shard_alloc first");
#if defined(ADSB_HOST_EMU)
    return fail(ctx, ADSB_E_PEER, "no CUDA IPC in the host emulation");
#else
    ON_DEVICE(ctx);
    adsb_shard_detach(ctx);
    for (int r = 0; r < world; r++) {
        if (r == rank) {
            ctx->shard.peer_base[r] = ctx->shard.own_base;
            continue;
        }
        cudaIpcMemHandle_t h;
        memcpy(&h, static_cast<const char *>(ipc_handles) + (size_t)r * ADSB_IPC_HANDLE_BYTES, sizeof(h));
        void *p = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            ctx->shard.world = r;          // close what was opened so far
            adsb_shard_detach(ctx);
            return fail(ctx, ADSB_E_PEER, "cudaIpcOpenMemHandle (peer access between the GPUs needed)", e);
        }
        ctx->shard.peer_base[r] = p;
        ctx->shard.ipc_opened[r] = true;
    }
    ctx->shard.rank = rank;
    ctx->shard.world = world;
    return ADSB_OK;
#endif
}

namespace {

// one step of shard `rank`, issued on stream s (phases: ADSB_SHARD_PH_*)
int shard_issue(adsb_ctx *ctx, int fmt, size_t items, float threshold, int phases, cudaStream_t s)
{
    const int rank = ctx->shard.rank, world = ctx->shard.world;
    const bool first = rank == 0, lastr = rank == world - 1;
    auto ctrl_of = [&](int r) { return reinterpret_cast<ShardCtrl *>(ctx->shard.peer_base[r]); };
    auto samples_of = [&](int r) { return static_cast<char *>(ctx->shard.peer_base[r]) + ADSB_SHARD_CTRL_BYTES; };
    ShardCtrl *own = ctrl_of(rank);
    ShardCtrl *left = first ? nullptr : ctrl_of(rank - 1), *right = lastr ? nullptr : ctrl_of(rank + 1);
    // window of this shard: [left neighbour's shard | own samples | right neighbour's shard]; rank 0 has the
    // 8*sps-1 zeros of GNU Radio's history instead of a left neighbour (whole-buffer semantics)
    const long long L = (long long)items, h1 = ctx->hist;
    const long long lead = first ? h1 : L;
    const long long own_lo = first ? 0 : lead;
    const long long own_hi = lead + L - (lastr ? h1 : 0);      // falls owned: [own_lo, own_hi)
    const long long base = first ? 0 : (long long)rank * L + h1 - lead;
    if (own_hi <= own_lo) return fail(ctx, ADSB_E_INVAL, "shard too small");
    const long long n_tiles = tiles_for(ctx, own_hi, lead);
    if ((size_t)n_tiles > ctx->tile_cap) return fail(ctx, ADSB_E_INVAL, "too many tiles for this context");
    int rc;
    const bool ph_sig = phases & ADSB_SHARD_PH_SIGNAL, ph_scan = phases & ADSB_SHARD_PH_SCAN;
    const bool ph_exp = phases & ADSB_SHARD_PH_EXPORT, ph_plan = phases & ADSB_SHARD_PH_PLAN, ph_check = phases & ADSB_SHARD_PH_CHECK;
    if (ph_sig && !ph_scan && world > 1) {
        ADSB_LAUNCH(shard_sync_kernel, 1, 1, 0, s, own, left, right, ctx->d_hdr, 1, 0);
        CK(cudaGetLastError());
    }
    if (ph_scan) {
        if ((rc = begin_call(ctx, n_tiles, own_hi - lead + 1, s))) return rc;
        if (world > 1 && ph_sig) {     // the wait for the neighbours' signals happens inside the scan kernel's edge tiles
            ADSB_LAUNCH(shard_sync_kernel, 1, 1, 0, s, own, left, right, ctx->d_hdr, 1, 0);
            CK(cudaGetLastError());
        }
        ScanParams P = make_params(ctx, samples_of(rank), (size_t)L, (size_t)lead, (size_t)own_hi, (size_t)own_lo, threshold,
                                   base, /*prev_bit*/ first ? (0.0f >= threshold) : 0, 1);
        if (!first) {
            P.in_left = samples_of(rank - 1);
            P.left_off = 0;                              // lead == L: window index i < lead is the neighbour's item i
        }
        if (!lastr) {
            P.in_right = samples_of(rank + 1);
            P.n_right = L;
        }
        P.win_hi = P.n_loc + P.n_right;
        if (world > 1) P.peer_ready = own->ready;     // ready[0], ready[1], step are consecutive (ShardCtrl)
        if (ctx->profile) CK(cudaEventRecord(ctx->pev[0], s));
        CK(launch_scan(ctx, P, fmt, 0u, (unsigned)n_tiles, s));
        if (ctx->profile) CK(cudaEventRecord(ctx->pev[1], s));
        if ((rc = launch_resolve(ctx, n_tiles, base, -1, s))) return rc;
        if (ctx->profile) CK(cudaEventRecord(ctx->pev[2], s));
        ctx->last.pending = true;
        ctx->last.n_out = own_hi;
        ctx->last.lead = lead;
        ctx->last.base = base;
        ctx->last.eob_in = -1;
        ctx->last.left_is_start = 0;       // sharded mode never advances the carried framer state
        ctx->last.stream = s;
    }
    if (world > 1) {
        PeerCtrls pc;
        memset(&pc, 0, sizeof(pc));
        for (int r = 0; r < world; r++) pc.p[r] = ctrl_of(r);
        const long long blank = 63LL * ctx->sps;
        if ((ph_exp || ph_plan || ph_check) && !ctx->last.pending)
            return fail(ctx, ADSB_E_INVAL, "adsb_shard_step: carry phase without a scan phase");
        // round 0: (look at the outcome of the previous exchange,) export the latest resolve's summary to every rank /
        // wait for everyone's and plan
        if (ph_exp || ph_plan)
            ADSB_LAUNCH(shard_carry_kernel, 1, kMaxWorld, 0, s, ctx->d_hdr, own, pc, world, rank, base, blank, 0, ph_exp ? 1 : 0,
                        ph_plan ? 1 : 0, ph_exp ? 1 : 0);
        if (ph_plan) {
            if ((rc = zero_cand_totals(ctx, ctx->last.n_out - ctx->last.lead + 1, s))) return rc;
            if ((rc = launch_resolve(ctx, n_tiles, base, 0, s, true, true, /*reset_done=*/true))) return rc;   // no-ops unless carry_need
        }
        // round 1: the same exchange after the fix-ups, only to check that nothing is left to fix; the check itself waits
        // for the next exchange (or an explicit CHECK phase) so that it costs the step no second rendezvous
        if (ph_plan || ph_check)
            ADSB_LAUNCH(shard_carry_kernel, 1, kMaxWorld, 0, s, ctx->d_hdr, own, pc, world, rank, base, blank, 1, ph_plan ? 1 : 0,
                        ph_check ? 1 : 0, 0);
        CK(cudaGetLastError());
    }
    return ADSB_OK;
}

}  // namespace

int adsb_shard_step(adsb_ctx *ctx, int fmt, size_t shard_items, float threshold, int phases, int use_graph, void *stream)
{
    ADSB_RANGE("adsb_shard_step");
    if (!ctx || fmt < 0 || fmt >= kNumFmt || shard_items == 0 || !(phases & (ADSB_SHARD_PH_ALL | ADSB_SHARD_PH_CHECK))) return ADSB_E_INVAL;
    if (!ctx->shard.world) return fail(ctx, ADSB_E_INVAL, "adsb_shard_step: attach first");
    if (shard_items > ctx->max_items) return fail(ctx, ADSB_E_INVAL, "shard_items exceeds max_items given to adsb_create");
    if (shard_items < (size_t)(120 * ctx->sps + 256)) return fail(ctx, ADSB_E_INVAL, "shard shorter than a burst plus the noise window");
    ON_DEVICE(ctx);
    cudaStream_t s = stream ? (cudaStream_t)stream : ctx->stream;
#if !defined(ADSB_HOST_EMU)
    if (use_graph && !ctx->profile && ctx->scan_configured[fmt]) {
        auto &k = ctx->shard.key;
        const bool same = ctx->shard.exec && k.fmt == fmt && k.phases == phases && k.items == shard_items &&
                          k.thr == threshold && k.stream == s;
        if (!same) {
            if (ctx->shard.exec) cudaGraphExecDestroy(ctx->shard.exec);
            if (ctx->shard.graph) cudaGraphDestroy(ctx->shard.graph);
            ctx->shard.exec = nullptr;
            ctx->shard.graph = nullptr;
            CK(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
            const int rc = shard_issue(ctx, fmt, shard_items, threshold, phases, s);
            cudaGraph_t g = nullptr;
            const cudaError_t e = cudaStreamEndCapture(s, &g);
            if (rc) {
                if (g) cudaGraphDestroy(g);
                return rc;
            }
            if (e != cudaSuccess) return fail(ctx, ADSB_E_CUDA, "cudaStreamEndCapture", e);
            ctx->shard.graph = g;
            CK(cudaGraphInstantiate(&ctx->shard.exec, g, 0));
            k.fmt = fmt;
            k.phases = phases;
            k.items = shard_items;
            k.thr = threshold;
            k.stream = s;
        }
        CK(cudaGraphLaunch(ctx->shard.exec, s));
        ctx->shard.steps_graph++;
        return ADSB_OK;
    }
#else
    (void)use_graph;
#endif
    ctx->shard.steps_plain++;
    return shard_issue(ctx, fmt, shard_items, threshold, phases, s);
}


// ---- streaming session -------------------------------------------------------------------------------------
namespace {

// pageable -> pinned staging copy: one core moves ~10 GB/s, PCIe takes 50; large pushes are split over a few threads
void big_memcpy(void *dst, const void *src, size_t bytes)
{
#if !defined(ADSB_HOST_EMU)
    constexpr size_t kPart = 4u << 20;
    unsigned hw = std::thread::hardware_concurrency();
    size_t parts = bytes / kPart;
    if (parts > 16) parts = 16;
    if (hw && parts > hw) parts = hw;
    if (parts >= 2) {
        std::thread th[16];
        const size_t per = (bytes / parts + 63) & ~(size_t)63;
        size_t off = 0;
        unsigned n = 0;
        for (; n < parts && off < bytes; n++) {
            const size_t len = (off + per < bytes && n + 1 < parts) ? per : bytes - off;
            th[n] = std::thread([=]() { memcpy(static_cast<char *>(dst) + off, static_cast<const char *>(src) + off, len); });
            off += len;
        }
        for (unsigned i = 0; i < n; i++) th[i].join();
        return;
    }
#endif
    memcpy(dst, src, bytes);
}

// scan + resolve of chunk k of the stream (k < submitted): its window is [chunk k-1 | chunk k | chunk k+1], read in place
int stream_process(adsb_ctx *ctx, long long k, bool last, unsigned long long s_tot)
{
    auto &S = ctx->st;
    const long long C = (long long)S.C, h1 = ctx->hist;
    const bool first = k == 0;
    const long long n_own = (long long)S.d_items[k % 3];
    const long long lead = first ? h1 : C;
    // falls owned: stream index [k*C, min((k+1)*C, s_tot - h1)) -- the last 8*sps-1 items of the stream are look-ahead
    // only, exactly as in one work() call over the whole stream (framer.py:72-77)
    long long hi_stream = (k + 1) * C;
    if (last || (long long)s_tot - h1 < hi_stream) hi_stream = (long long)s_tot - h1;
    const long long own_lo = first ? 0 : lead;
    const long long own_hi = lead + (hi_stream - k * C);
    const long long base = first ? 0 : k * C + h1 - lead;
    cudaStream_t s = ctx->stream;
    S.processed = k + 1;
    if (own_hi <= own_lo) {          // nothing owned (a tail shorter than the look-ahead)
        ctx->last.pending = false;
        return ADSB_OK;
    }
    const long long n_tiles = tiles_for(ctx, own_hi, lead);
    if ((size_t)n_tiles > ctx->tile_cap) return fail(ctx, ADSB_E_INVAL, "chunk_items exceeds max_items given to adsb_create");
    int rc;
    CK(cudaStreamWaitEvent(s, S.ev_h2d[k % 3], 0));
    if (!last) CK(cudaStreamWaitEvent(s, S.ev_h2d[(k + 1) % 3], 0));
    if ((rc = begin_call(ctx, n_tiles, own_hi - lead + 1, s))) return rc;
    ScanParams P = make_params(ctx, S.d_buf[k % 3], (size_t)n_own, (size_t)lead, (size_t)own_hi, (size_t)own_lo, S.thr, base,
                               first ? (0.0f >= S.thr) : 0, 1);
    if (!first) {
        P.in_left = S.d_buf[(k - 1) % 3];
        P.left_off = 0;
    }
    if (!last) {
        P.in_right = S.d_buf[(k + 1) % 3];
        P.n_right = (long long)S.d_items[(k + 1) % 3];
    }
    P.win_hi = P.n_loc + P.n_right;
    CK(launch_scan(ctx, P, S.fmt, 0u, (unsigned)n_tiles, s));
    const long long eob_in = (S.e_global == kCarryNone) ? -1 - base : S.e_global - base;     // window coordinates of this chunk
    if ((rc = launch_resolve(ctx, n_tiles, base, eob_in, s))) return rc;
    CK(cudaEventRecord(S.ev_scan[k % 3], s));
    ctx->last.pending = true;
    ctx->last.n_out = own_hi;
    ctx->last.lead = lead;
    ctx->last.base = base;
    ctx->last.eob_in = eob_in;
    ctx->last.left_is_start = 0;       // a streaming session never touches the framer state of the block entry points
    ctx->last.stream = s;
    return ADSB_OK;
}

// fetch the hits of the chunk processed last into the session's result list
int stream_collect(adsb_ctx *ctx)
{
    auto &S = ctx->st;
    if (S.collected >= S.processed) return ADSB_OK;
    S.collected = S.processed;
    if (!ctx->last.pending) return ADSB_OK;
    adsb_counters c;
    size_t n = 0;
    int rc = adsb_burst_finish(ctx, &c, nullptr, 0, &n);
    if (rc) return rc;
    if (S.out_n + n > S.out_cap) {
        size_t want = S.out_cap ? S.out_cap * 2 : 4096;
        while (want < S.out_n + n) want *= 2;
        Rec *p = (Rec *)realloc(S.out, want * sizeof(Rec));
        if (!p) return fail(ctx, ADSB_E_NOMEM, "realloc(stream hits)");
        S.out = p;
        S.out_cap = want;
    }
    if (n) {
        CK(cudaMemcpy(S.out + S.out_n, ctx->d_hits, n * sizeof(Rec), cudaMemcpyDeviceToHost));
        S.out_n += n;
    }
    S.n_cand += c.n_candidates;
    S.n_crc += c.n_crc_ok + c.n_fec_fixed;
    S.n_demod += c.n_demod;
    if (c.last_accept >= 0) S.e_global = ctx->last.base + c.last_accept + 63LL * ctx->sps;
    ctx->last.pending = false;
    return ADSB_OK;
}

// hand the staging buffer being filled (fill items) to the copy engine as the next chunk; scan the chunk before it
int stream_submit(adsb_ctx *ctx, bool final_chunk)
{
    auto &S = ctx->st;
    const long long j = S.submitted;
    int rc;
    // chunk j goes where chunk j-3 was: chunk j-2, whose scan reads that buffer as its left neighbour, must be done
    if (j >= 3) CK(cudaStreamWaitEvent(ctx->copy_stream, S.ev_scan[(j - 2) % 3], 0));
    CK(cudaMemcpyAsync(S.d_buf[j % 3], S.h_pin[S.cur], S.fill * S.isz, cudaMemcpyHostToDevice, ctx->copy_stream));
    CK(cudaEventRecord(S.ev_copy[S.cur], ctx->copy_stream));
    CK(cudaEventRecord(S.ev_h2d[j % 3], ctx->copy_stream));
    S.d_items[j % 3] = S.fill;
    S.submitted = j + 1;
    S.cur ^= 1;
    S.fill = 0;
    if (j >= 1) {
        if ((rc = stream_collect(ctx))) return rc;            // chunk j-2 (its header must be read before the next reset)
        if ((rc = stream_process(ctx, j - 1, false, final_chunk ? S.total_items : ~0ULL >> 1))) return rc;
    }
    // the other staging buffer is filled next: its previous copy must have left it
    CK(cudaEventSynchronize(S.ev_copy[S.cur]));
    return ADSB_OK;
}

}  // namespace

int adsb_stream_close(adsb_ctx *ctx)
{
    if (!ctx) return ADSB_E_INVAL;
    auto &S = ctx->st;
    if (!S.open && !S.out && !S.d_buf[0]) return ADSB_OK;
    DeviceGuard guard(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    if (ctx->copy_stream) cudaStreamSynchronize(ctx->copy_stream);
    for (int i = 0; i < 3; i++) {
        cudaFree(S.d_buf[i]);
        if (S.ev_h2d[i]) cudaEventDestroy(S.ev_h2d[i]);
        if (S.ev_scan[i]) cudaEventDestroy(S.ev_scan[i]);
    }
    for (int i = 0; i < 2; i++) {
        if (S.h_pin[i]) cudaFreeHost(S.h_pin[i]);
        if (S.ev_copy[i]) cudaEventDestroy(S.ev_copy[i]);
    }
    free(S.out);
    memset(&S, 0, sizeof(S));
    ctx->last.pending = false;
    return ADSB_OK;
}

int adsb_stream_open(adsb_ctx *ctx, int fmt, float threshold, size_t chunk_items)
{
    if (!ctx || fmt < 0 || fmt >= kNumFmt) return ADSB_E_INVAL;
    if (chunk_items > ctx->max_items) return fail(ctx, ADSB_E_INVAL, "chunk_items exceeds max_items given to adsb_create");
    if (chunk_items < (size_t)(120 * ctx->sps + 256)) return fail(ctx, ADSB_E_INVAL, "chunk shorter than a burst plus the noise window");
    ON_DEVICE(ctx);
    adsb_stream_close(ctx);
    auto &S = ctx->st;
    S.fmt = fmt;
    S.thr = threshold;
    S.C = chunk_items;
    S.isz = (fmt == kFmtC64) ? 8 : (fmt == kFmtCU8) ? 2 : 4;
    S.e_global = kCarryNone;
    cudaError_t e = cudaSuccess;
    for (int i = 0; i < 3 && e == cudaSuccess; i++) {
        e = cudaMalloc(&S.d_buf[i], S.C * S.isz + 256);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&S.ev_h2d[i], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&S.ev_scan[i], cudaEventDisableTiming);
    }
    for (int i = 0; i < 2 && e == cudaSuccess; i++) {
        e = cudaHostAlloc(&S.h_pin[i], S.C * S.isz, cudaHostAllocMapped);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&S.ev_copy[i], cudaEventDisableTiming);
    }
    if (e != cudaSuccess) {
        adsb_stream_close(ctx);
        return fail(ctx, ADSB_E_NOMEM, "stream buffers", e);
    }
    S.open = true;
    return ADSB_OK;
}

int adsb_stream_acquire(adsb_ctx *ctx, void **host_buf, size_t *cap_items)
{
    if (!ctx || !host_buf || !cap_items) return ADSB_E_INVAL;
    auto &S = ctx->st;
    if (!S.open || S.flushed) return fail(ctx, ADSB_E_INVAL, "no open stream");
    *host_buf = static_cast<char *>(S.h_pin[S.cur]) + S.fill * S.isz;
    *cap_items = S.C - S.fill;
    return ADSB_OK;
}

int adsb_stream_commit(adsb_ctx *ctx, size_t n_items)
{
    ADSB_RANGE("adsb_stream_commit");
    if (!ctx) return ADSB_E_INVAL;
    auto &S = ctx->st;
    if (!S.open || S.flushed) return fail(ctx, ADSB_E_INVAL, "no open stream");
    if (n_items > S.C - S.fill) return fail(ctx, ADSB_E_INVAL, "adsb_stream_commit: more than adsb_stream_acquire offered");
    ON_DEVICE(ctx);
    S.fill += n_items;
    S.total_items += n_items;
    if (S.fill == S.C) return stream_submit(ctx, false);
    return ADSB_OK;
}

int adsb_stream_push(adsb_ctx *ctx, const void *samples, size_t n_items)
{
    if (!ctx || (!samples && n_items)) return ADSB_E_INVAL;
    auto &S = ctx->st;
    if (!S.open || S.flushed) return fail(ctx, ADSB_E_INVAL, "no open stream");
    const char *src = static_cast<const char *>(samples);
    while (n_items) {
        void *dst;
        size_t cap;
        int rc = adsb_stream_acquire(ctx, &dst, &cap);
        if (rc) return rc;
        const size_t n = n_items < cap ? n_items : cap;
        big_memcpy(dst, src, n * S.isz);
        if ((rc = adsb_stream_commit(ctx, n))) return rc;
        src += n * S.isz;
        n_items -= n;
    }
    return ADSB_OK;
}

int adsb_stream_flush(adsb_ctx *ctx)
{
    ADSB_RANGE("adsb_stream_flush");
    if (!ctx) return ADSB_E_INVAL;
    auto &S = ctx->st;
    if (!S.open) return fail(ctx, ADSB_E_INVAL, "no open stream");
    if (S.flushed) return ADSB_OK;
    ON_DEVICE(ctx);
    int rc;
    if (S.fill > 0) {
        if ((rc = stream_submit(ctx, true))) return rc;       // also scans the chunk before it, with the true stream end
    }
    S.flushed = true;
    if (S.submitted == 0) return ADSB_OK;
    if ((rc = stream_collect(ctx))) return rc;
    if (S.processed < S.submitted) {
        if ((rc = stream_process(ctx, S.submitted - 1, true, S.total_items))) return rc;
        if ((rc = stream_collect(ctx))) return rc;
    }
    return ADSB_OK;
}

int adsb_stream_pop(adsb_ctx *ctx, adsb_hit *hits, size_t cap, size_t *n_hits)
{
    if (!ctx || !n_hits || (!hits && cap)) return ADSB_E_INVAL;
    auto &S = ctx->st;
    if (!S.open) return fail(ctx, ADSB_E_INVAL, "no open stream");
    size_t n = S.out_n - S.out_read;
    if (n > cap) n = cap;
    if (n) memcpy(hits, S.out + S.out_read, n * sizeof(Rec));
    S.out_read += n;
    *n_hits = n;
    return ADSB_OK;
}

int adsb_stream_stats(const adsb_ctx *ctx, uint64_t *items, uint64_t *chunks_done, uint64_t *hits, uint64_t *crc_ok)
{
    if (!ctx) return ADSB_E_INVAL;
    const auto &S = ctx->st;
    if (items) *items = S.total_items;
    if (chunks_done) *chunks_done = (uint64_t)S.collected;
    if (hits) *hits = S.out_n;
    if (crc_ok) *crc_ok = S.n_crc;
    return ADSB_OK;
}

int adsb_burst_finish(adsb_ctx *ctx, adsb_counters *counters, adsb_hit *hits, size_t cap, size_t *n_hits)
{
    ADSB_RANGE("adsb_burst_finish");
    if (!ctx || !ctx->last.pending) return ADSB_E_INVAL;
    ON_DEVICE(ctx);
    cudaStream_t s = ctx->last.stream;
    CK(cudaMemcpyAsync(&ctx->h_stage->hdr, ctx->d_hdr, sizeof(Header), cudaMemcpyDeviceToHost, s));
    if (hits && cap)
        CK(cudaMemcpyAsync(ctx->h_stage->first, ctx->d_hits, sizeof(Rec) * kFirstHits, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    const Header &h = ctx->h_stage->hdr;
    int rc = status_of(ctx, h);
    if (h.carry_unstable)       // per-call counter: hand it out once, then start the next call from zero
        cudaMemsetAsync(reinterpret_cast<char *>(ctx->d_hdr) + offsetof(Header, carry_unstable), 0, sizeof(unsigned), s);
    if (rc) {
        // the call's tags are lost, but the stream position is not: prev_in0 must become x[n_out - 1] (framer.py:87)
        if (ctx->last.left_is_start && !(h.flags & kFlagPeerTimeout))
            ctx->prev_in0 = (ctx->last.n_out - 1 >= ctx->last.lead) ? h.x_last : 0.0f;
        ctx->last.pending = false;
        return rc;
    }
    adsb_counters &c = ctx->counters;
    c.n_samples = (uint64_t)ctx->last.n_out;
    c.n_candidates = h.n_cand;
    c.n_hits = h.n_hits;
    c.n_demod = h.n_demod;
    c.n_crc_ok = h.n_crc;
    c.last_pulse = h.last_pulse;
    c.last_accept = h.last_accept;
    c.first_candidate = h.first_cand;
    c.carry_unstable = h.carry_unstable;
    c.n_fec_fixed = h.n_fec;
    c.any_rise = (h.flags & kFlagRise) ? 1u : 0u;
    c.any_fall = (h.flags & kFlagFall) ? 1u : 0u;
    if (counters) *counters = c;
    if (n_hits) *n_hits = (size_t)h.n_hits;
    if (ctx->last.left_is_start) {
        // advance the carried state (framer.py:87,121-123,165,177-179)
        ctx->prev_eob = next_eob(h, ctx->last.eob_in, ctx->last.n_out, ctx->sps);
        ctx->prev_in0 = (ctx->last.n_out - 1 >= ctx->last.lead) ? h.x_last : 0.0f;
    }
    if (hits) {
        if (h.n_hits > cap) return fail(ctx, ADSB_E_CAPACITY, "hit capacity too small");
        const size_t n = (size_t)h.n_hits;
        const size_t nf = n < (size_t)kFirstHits ? n : (size_t)kFirstHits;
        memcpy(hits, ctx->h_stage->first, nf * sizeof(Rec));
        if (n > nf)
            CK(cudaMemcpy(hits + nf, ctx->d_hits + nf, (n - nf) * sizeof(Rec), cudaMemcpyDeviceToHost));
    }
    return ADSB_OK;
}

int adsb_burst_work(adsb_ctx *ctx, const void *in, int is_c64, size_t n_in, size_t lead, size_t n_out,
                    float threshold, int64_t base, adsb_hit *hits, size_t cap, size_t *n_hits)
{
    ADSB_RANGE("adsb_burst_work");
    if (!ctx || !in || is_c64 < 0 || is_c64 >= kNumFmt) return ADSB_E_INVAL;
    int rc = check_geometry(ctx, n_in, lead, n_out, 0);
    if (rc) return rc;
    ON_DEVICE(ctx);
    if (small_eligible(ctx, is_c64, n_in, lead, n_out)) {
        // GNU-Radio-sized call: one kernel launch, no copy calls (adsb_small.cuh)
        size_t n = 0;
        if ((rc = small_call(ctx, in, is_c64, n_in, lead, n_out, threshold, base, 1, &n))) return rc;
        if (n_hits) *n_hits = n;
        if (hits) {
            if (n > cap) return fail(ctx, ADSB_E_CAPACITY, "hit capacity too small");
            if ((rc = small_fetch_hits(ctx, reinterpret_cast<Rec *>(hits), n))) return rc;
        }
        return ADSB_OK;
    }
    const size_t isz = (is_c64 == kFmtC64) ? 8 : (is_c64 == kFmtCU8) ? 2 : 4;
    ctx->staged_fmt = -1;
    if ((rc = ensure_input(ctx, n_in * isz))) return rc;
    const long long n_tiles = tiles_for(ctx, (long long)n_out, (long long)lead);
    if ((size_t)n_tiles > ctx->tile_cap) return fail(ctx, ADSB_E_INVAL, "too many tiles for this context");
    const int prev_bit = (lead != 0) ? (0.0f >= threshold) : (ctx->prev_in0 >= threshold);
    const long long eob_in = ctx->prev_eob;
    cudaStream_t s = ctx->stream;
    if ((rc = begin_call(ctx, n_tiles, (long long)n_out, s))) return rc;
    ScanParams P = make_params(ctx, ctx->d_in, n_in, lead, n_out, 0, threshold, base, prev_bit, 1);

    // H2D in chunks on the copy stream; tiles are launched as soon as their samples (incl. the right
    // halo) are resident, so the scan overlaps the transfer of the following chunk.
    const char *src = static_cast<const char *>(in);
    char *dst = static_cast<char *>(ctx->d_in);
    size_t copied = 0;
    long long tiles_done = 0;
    int ring = 0;
    // A pageable source (np.fromfile, GNU Radio buffers ...) would make every cudaMemcpyAsync a synchronous, staged copy:
    // stage it ourselves through two pinned chunk buffers, a few host threads copying chunk i+1 while chunk i crosses PCIe.
    bool pageable = true;
#if !defined(ADSB_HOST_EMU)
    {
        cudaPointerAttributes attr;
        if (cudaPointerGetAttributes(&attr, in) == cudaSuccess) pageable = attr.type == cudaMemoryTypeUnregistered;
        else cudaGetLastError();
    }
#endif
    if (pageable) {
        for (int i = 0; i < 2; i++) {
            if (!ctx->h_ring[i]) {
                cudaError_t e = cudaHostAlloc(&ctx->h_ring[i], kPipeChunk * 8, cudaHostAllocMapped);
                if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_ring[i], cudaEventDisableTiming);
                if (e != cudaSuccess) return fail(ctx, ADSB_E_NOMEM, "pinned staging ring", e);
            }
        }
    }
    int hb = 0;
    while (copied < n_in) {
        const size_t n = (n_in - copied < kPipeChunk) ? (n_in - copied) : kPipeChunk;
        if (pageable) {
            CK(cudaEventSynchronize(ctx->ev_ring[hb]));          // the copy that last used this staging buffer has left it
            big_memcpy(ctx->h_ring[hb], src + copied * isz, n * isz);
            CK(cudaMemcpyAsync(dst + copied * isz, ctx->h_ring[hb], n * isz, cudaMemcpyHostToDevice, ctx->copy_stream));
            CK(cudaEventRecord(ctx->ev_ring[hb], ctx->copy_stream));
            hb ^= 1;
        } else {
            CK(cudaMemcpyAsync(dst + copied * isz, src + copied * isz, n * isz, cudaMemcpyHostToDevice, ctx->copy_stream));
        }
        copied += n;
        long long ready = (copied == n_in) ? n_tiles : ((long long)copied - ctx->rh) / ctx->T;
        if (ready > n_tiles) ready = n_tiles;
        if (ready > tiles_done || copied == n_in) {
            CK(cudaEventRecord(ctx->ev[ring], ctx->copy_stream));
            CK(cudaStreamWaitEvent(s, ctx->ev[ring], 0));
            ring = (ring + 1) % kEvRing;
            if (ready > tiles_done) CK(launch_scan(ctx, P, is_c64, (unsigned)tiles_done, (unsigned)(ready - tiles_done), s));
            tiles_done = ready;
        }
    }
    if ((rc = launch_resolve(ctx, n_tiles, base, eob_in, s))) return rc;
    ctx->staged_fmt = is_c64;
    ctx->staged_n = n_in;
    ctx->last.pending = true;
    ctx->last.n_out = (long long)n_out;
    ctx->last.lead = (long long)lead;
    ctx->last.base = base;
    ctx->last.eob_in = eob_in;
    ctx->last.left_is_start = 1;
    ctx->last.stream = s;
    return adsb_burst_finish(ctx, nullptr, hits, cap, n_hits);
}

int adsb_burst_mag_out(adsb_ctx *ctx, float *out, size_t n_out)
{
    if (!ctx || !out) return ADSB_E_INVAL;
    if (ctx->small.last_was_small && n_out && n_out <= ctx->staged_n) {
        // latency path: the kernel left |IQ|^2 of the call's samples in pinned memory (float32 input: the samples themselves)
        const float *src = ctx->small_fmt == kFmtC64 ? ctx->small.mag : static_cast<const float *>(ctx->small.in);
        memcpy(out, src + (ctx->staged_n - n_out), n_out * sizeof(float));
        return ADSB_OK;
    }
    if (ctx->staged_fmt < 0 || n_out == 0 || n_out > ctx->staged_n)
        return fail(ctx, ADSB_E_INVAL, "adsb_burst_mag_out: needs a preceding adsb_burst_work with n_in >= n_out");
    ON_DEVICE(ctx);
    if (ctx->d_mag_cap < n_out) {
        cudaFree(ctx->d_mag);
        ctx->d_mag = nullptr;
        ctx->d_mag_cap = 0;
        const size_t want = ctx->max_items > n_out ? ctx->max_items : n_out;
        cudaError_t e = cudaMalloc((void **)&ctx->d_mag, want * sizeof(float));
        if (e != cudaSuccess) return fail(ctx, ADSB_E_NOMEM, "cudaMalloc(|IQ|^2 output)", e);
        ctx->d_mag_cap = want;
    }
    cudaStream_t s = ctx->stream;
    const long long first = (long long)(ctx->staged_n - n_out);
    size_t want_blocks = (n_out + kThreads * 4 - 1) / (kThreads * 4);
    const size_t max_blocks = (size_t)ctx->num_sms * 8;
    const unsigned blocks = (unsigned)(want_blocks < max_blocks ? want_blocks : max_blocks);
    switch (ctx->staged_fmt) {
    case kFmtC64: ADSB_LAUNCH(mag_out_kernel<kFmtC64>, blocks, kThreads, 0, s, ctx->d_in, first, (long long)n_out, ctx->d_mag); break;
    case kFmtSC16: ADSB_LAUNCH(mag_out_kernel<kFmtSC16>, blocks, kThreads, 0, s, ctx->d_in, first, (long long)n_out, ctx->d_mag); break;
    case kFmtCU8: ADSB_LAUNCH(mag_out_kernel<kFmtCU8>, blocks, kThreads, 0, s, ctx->d_in, first, (long long)n_out, ctx->d_mag); break;
    default: ADSB_LAUNCH(mag_out_kernel<kFmtF32>, blocks, kThreads, 0, s, ctx->d_in, first, (long long)n_out, ctx->d_mag); break;
    }
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out, ctx->d_mag, n_out * sizeof(float), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return ADSB_OK;
}

int adsb_framer_work(adsb_ctx *ctx, const float *in0, size_t n_in, size_t n_out, float threshold,
                     int64_t *hit_idx, float *hit_snr, size_t cap, size_t *n_hits)
{
    ADSB_RANGE("adsb_framer_work");
    if (!ctx || !in0 || !n_hits) return ADSB_E_INVAL;
    *n_hits = 0;
    int rc = check_geometry(ctx, n_in, 0, n_out, 0);
    if (rc) return rc;
    ON_DEVICE(ctx);
    if (small_eligible(ctx, kFmtF32, n_in, 0, n_out)) {
        size_t n = 0;
        if ((rc = small_call(ctx, in0, kFmtF32, n_in, 0, n_out, threshold, 0, 0, &n))) return rc;
        *n_hits = n;
        if (n > cap || (n && (!hit_idx || !hit_snr))) return fail(ctx, ADSB_E_CAPACITY, "hit capacity too small");
        if (ctx->h_tmp_cap < n) {
            free(ctx->h_tmp);
            ctx->h_tmp = (Rec *)malloc((n + 64) * sizeof(Rec));
            ctx->h_tmp_cap = ctx->h_tmp ? n + 64 : 0;
            if (!ctx->h_tmp) return fail(ctx, ADSB_E_NOMEM, "malloc(hit scratch)");
        }
        if ((rc = small_fetch_hits(ctx, ctx->h_tmp, n))) return rc;
        for (size_t i = 0; i < n; i++) {
            hit_idx[i] = ctx->h_tmp[i].idx;
            hit_snr[i] = ctx->h_tmp[i].snr;
        }
        return ADSB_OK;
    }
    if ((rc = ensure_input(ctx, n_in * 4))) return rc;
    ctx->staged_fmt = -1;
    cudaStream_t s = ctx->stream;
    const long long n_tiles = tiles_for(ctx, (long long)n_out, 0);
    const int prev_bit = ctx->prev_in0 >= threshold;
    const long long eob_in = ctx->prev_eob;
    CK(cudaMemcpyAsync(ctx->d_in, in0, n_in * 4, cudaMemcpyHostToDevice, s));
    if ((rc = begin_call(ctx, n_tiles, (long long)n_out, s))) return rc;
    ScanParams P = make_params(ctx, ctx->d_in, n_in, 0, n_out, 0, threshold, 0, prev_bit, 0);
    CK(launch_scan(ctx, P, kFmtF32, 0u, (unsigned)n_tiles, s));
    if ((rc = launch_resolve(ctx, n_tiles, 0, eob_in, s))) return rc;
    ctx->last.pending = true;
    ctx->last.n_out = (long long)n_out;
    ctx->last.lead = 0;
    ctx->last.base = 0;
    ctx->last.eob_in = eob_in;
    ctx->last.left_is_start = 1;
    ctx->last.stream = s;
    // header + hits in one synchronisation (the first kFirstHits ride along with the header)
    const size_t tcap = n_out / (size_t)(63 * ctx->sps) + 2;      // blanking bounds the hits of one call
    if (ctx->h_tmp_cap < tcap) {
        free(ctx->h_tmp);
        ctx->h_tmp = (Rec *)malloc(tcap * sizeof(Rec));
        ctx->h_tmp_cap = ctx->h_tmp ? tcap : 0;
        if (!ctx->h_tmp) return fail(ctx, ADSB_E_NOMEM, "malloc(hit scratch)");
    }
    Rec *tmp = ctx->h_tmp;
    size_t n = 0;
    rc = adsb_burst_finish(ctx, nullptr, reinterpret_cast<adsb_hit *>(tmp), tcap, &n);
    if (rc == ADSB_OK) {
        *n_hits = n;
        if (n > cap || (n && (!hit_idx || !hit_snr))) {
            rc = fail(ctx, ADSB_E_CAPACITY, "hit capacity too small");
        } else {
            for (size_t i = 0; i < n; i++) {
                hit_idx[i] = tmp[i].idx;
                hit_snr[i] = tmp[i].snr;
            }
        }
    }
    return rc;
}

int adsb_demod_work(adsb_ctx *ctx, const float *in0, size_t n, const int64_t *sob_idx, size_t n_tags,
                    uint8_t *bits, uint8_t *valid)
{
    ADSB_RANGE("adsb_demod_work");
    if (!ctx) return ADSB_E_INVAL;
    if (n_tags == 0) return ADSB_OK;             // the common work() call: nothing tagged, nothing to do
    if (!in0 || !sob_idx || !bits || !valid) return ADSB_E_INVAL;
    if (n > ctx->max_items + (size_t)ctx->hist) return fail(ctx, ADSB_E_INVAL, "n exceeds max_items");
    ON_DEVICE(ctx);
    int rc;
    {
        // latency path: only the span the tags touch travels, through pinned memory the kernel reads and writes in place
        long long lo = (long long)n, hi = 0;
        for (size_t t = 0; t < n_tags; t++) {
            const long long a = sob_idx[t], b = sob_idx[t] + 112LL * ctx->sps;
            if (a < lo) lo = a;
            if (b > hi) hi = b;
        }
        if (lo < 0) lo = 0;
        if (hi > (long long)n) hi = (long long)n;
        const size_t span = hi > lo ? (size_t)(hi - lo) : 0;
        const size_t ring = ((size_t)kSmallTiles * ctx->T + ctx->rh + ctx->hist + 64) * 2;     // floats in the staging buffer
        if (!ctx->no_small && n_tags <= 1024 && span <= ring) {
            cudaStream_t s = ctx->stream;
            unsigned char *p = static_cast<unsigned char *>(ctx->small.in);
            size_t have = ctx->small.in_bytes;
            if ((rc = pinned_grow(ctx, &p, &have, ring * 4, "pinned input staging"))) return rc;
            ctx->small.in = p;
            ctx->small.in_bytes = have;
            if (!ctx->small.out) {
                cudaError_t e = cudaHostAlloc((void **)&ctx->small.out, sizeof(SmallOut), cudaHostAllocMapped);
                if (e != cudaSuccess) return fail(ctx, ADSB_E_NOMEM, "pinned result block", e);
                memset(ctx->small.out, 0, sizeof(SmallOut));
            }
            if (ctx->small.tag_cap < n_tags) {
                size_t c1 = ctx->small.tag_cap, c2 = ctx->small.tag_cap, c3 = ctx->small.tag_cap;
                const size_t want = n_tags < 64 ? 64 : n_tags;
                if ((rc = pinned_grow(ctx, &ctx->small.sob, &c1, want, "pinned tag staging"))) return rc;
                c2 *= kMsgBits;
                if ((rc = pinned_grow(ctx, &ctx->small.bits, &c2, want * kMsgBits, "pinned bit staging"))) return rc;
                if ((rc = pinned_grow(ctx, &ctx->small.valid, &c3, want, "pinned tag staging"))) return rc;
                ctx->small.tag_cap = want;
            }
            if (!ctx->small.d_done) {
                CK(cudaMalloc((void **)&ctx->small.d_done, sizeof(unsigned)));
                CK(cudaMemset(ctx->small.d_done, 0, sizeof(unsigned)));
            }
            if (span) memcpy(ctx->small.in, in0 + lo, span * sizeof(float));
            memcpy(ctx->small.sob, sob_idx, n_tags * sizeof(long long));
            ctx->small.last_was_small = false;
            ctx->staged_fmt = -1;
            unsigned blocks = (unsigned)((n_tags + kWarps - 1) / kWarps);
            if (blocks > 64u) blocks = 64u;
            const unsigned long long seq = ++ctx->small.seq;
            ADSB_LAUNCH(small_demod_kernel, blocks, kThreads, 0, s, static_cast<const float *>(ctx->small.in), lo, (long long)n,
                        ctx->small.sob, (int)n_tags, ctx->sps, ctx->small.bits, ctx->small.valid, &ctx->small.out->seq, seq,
                        ctx->small.d_done);
            CK(cudaGetLastError());
            if ((rc = wait_stamp(ctx, &ctx->small.out->seq, seq, s))) return rc;
            for (size_t t = 0; t < n_tags; t++) {
                valid[t] = ctx->small.valid[t];
                if (valid[t]) memcpy(bits + t * kMsgBits, ctx->small.bits + t * kMsgBits, kMsgBits);   // straddlers' bits stay untouched
            }
            return ADSB_OK;
        }
    }
    if ((rc = ensure_input(ctx, n * 4))) return rc;
    ctx->staged_fmt = -1;
    ctx->small.last_was_small = false;
    if (n_tags > ctx->tag_cap) {
        cudaFree(ctx->d_sob);
        cudaFree(ctx->d_bits);
        cudaFree(ctx->d_valid);
        ctx->d_sob = nullptr;
        ctx->d_bits = ctx->d_valid = nullptr;
        ctx->tag_cap = 0;
        size_t want = n_tags * 2;
        CK(cudaMalloc((void **)&ctx->d_sob, want * sizeof(long long)));
        CK(cudaMalloc((void **)&ctx->d_bits, want * kMsgBits));
        CK(cudaMalloc((void **)&ctx->d_valid, want));
        ctx->tag_cap = want;
    }
    cudaStream_t s = ctx->stream;
    // only the span the tags can touch needs to travel
    long long lo = (long long)n, hi = 0;
    for (size_t t = 0; t < n_tags; t++) {
        long long a = sob_idx[t], b = sob_idx[t] + 112LL * ctx->sps;
        if (a < lo) lo = a;
        if (b > hi) hi = b;
    }
    if (lo < 0) lo = 0;
    if (hi > (long long)n) hi = (long long)n;
    if (hi > lo)
        CK(cudaMemcpyAsync((float *)ctx->d_in + lo, in0 + lo, (size_t)(hi - lo) * 4, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(ctx->d_sob, sob_idx, n_tags * sizeof(long long), cudaMemcpyHostToDevice, s));
    unsigned blocks = (unsigned)((n_tags + kWarps - 1) / kWarps);
    if (blocks > (unsigned)ctx->num_sms * 8u) blocks = (unsigned)ctx->num_sms * 8u;
    ADSB_LAUNCH(demod_kernel, blocks, kThreads, 0, s, (const float *)ctx->d_in, (long long)n, ctx->d_sob, (int)n_tags, ctx->sps,
                                             ctx->d_bits, ctx->d_valid);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(valid, ctx->d_valid, n_tags, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    // bits of straddlers stay untouched in the caller's buffer
    bool all = true;
    for (size_t t = 0; t < n_tags; t++) all = all && valid[t];
    if (all) {
        CK(cudaMemcpy(bits, ctx->d_bits, n_tags * kMsgBits, cudaMemcpyDeviceToHost));
    } else {
        if (ctx->h_bits_cap < n_tags * kMsgBits) {
            free(ctx->h_bits);
            ctx->h_bits = (unsigned char *)malloc(n_tags * kMsgBits);
            ctx->h_bits_cap = ctx->h_bits ? n_tags * kMsgBits : 0;
            if (!ctx->h_bits) return fail(ctx, ADSB_E_NOMEM, "malloc(bit scratch)");
        }
        CK(cudaMemcpy(ctx->h_bits, ctx->d_bits, n_tags * kMsgBits, cudaMemcpyDeviceToHost));
        for (size_t t = 0; t < n_tags; t++)
            if (valid[t]) memcpy(bits + t * kMsgBits, ctx->h_bits + t * kMsgBits, kMsgBits);
    }
    return ADSB_OK;
}

}  // extern "C"
