// The latency path: ONE kernel launch per work() call of GNU-Radio size (a few tiles), no copy calls.
//
// gr-adsb's blocks are called by the GNU Radio scheduler with a few thousand items at a time (framer.py:72-77:
// len(output_items[0]) is typically 8192).  At that size the throughput pipeline (H2D copy, header init, two memsets,
// scan, gather, chain, compact, D2H copy, stream sync) is nine driver operations for a few microseconds of work.  Here the
// samples are read straight from a pinned, device-mapped staging buffer over PCIe, every tile is scanned and sliced by
// the same scan_tile / slice_tile code as the throughput kernel, and the CTA that finishes last resolves the call on the
// spot -- ordered candidate list, 63-symbol blanking (framer.py:121-123,165), hit records and counters -- writing the results
// into pinned host memory and a completion stamp the host spins on.  The device-side header is left clean for the next call.
#pragma once
#include "adsb_resolve.cuh"
#include "adsb_scan.cuh"

namespace adsb {

// L2-coherent 16-byte load (records other CTAs of the same launch wrote)
__device__ __forceinline__ uint4 ld_cg_u4(const uint4 *p)
{
#if defined(ADSB_HOST_EMU)
    return *p;
#else
    return __ldcg(p);
#endif
}

constexpr int kSmallTiles = 64;          // tiles per call on this path (320 Ki samples at 2 / 4 Msps)
constexpr int kSmallCand = 2048;         // candidates resolved in shared memory; more -> the host runs the resolve kernels
constexpr int kSmallHits = 2624;         // >= (kSmallTiles * 5120) / (63 * 2) + 2 hit records in the pinned result block (more -> fallback)
constexpr unsigned kFlagSmallOverflow = 64u;

#ifndef ADSB_SMALL_TIMING
#define ADSB_SMALL_TIMING 0              // A/B builds: phase timestamps (%globaltimer) of the last CTA in SmallOut::t
#endif

struct SmallOut {                        // pinned, device-mapped host memory
    Header hdr;                          // the call's summary, as adsb_burst_finish reads it
    unsigned long long seq;              // completion stamp (written last)
    unsigned long long pad;
    unsigned long long t[8];             // ADSB_SMALL_TIMING: kernel entry, tile staged, detected, sliced, ticket, resolved
    Rec hits[kSmallHits];
};
#if ADSB_SMALL_TIMING
#define ADSB_TSTAMP(k) do { if (threadIdx.x == 0) tstamp[k] = now_ns(); } while (0)
#else
#define ADSB_TSTAMP(k) do { } while (0)
#endif

struct SmallParams {
    SmallOut *out;
    unsigned long long seq;
    long long eob_in;                    // incoming prev_eob_idx, window coordinates
    int blank;                           // 63 * sps
    int n_tiles;
    float *mag_out;                      // |IQ|^2 of window items [mag_first, mag_first + mag_n) (pinned host), or null
    long long mag_first, mag_n;
};

// (one CTA per SM is all this kernel ever gets to run: no register cap, so that an edge tile's loads all go out at once)
template <int SPS, int FMT>
__global__ void __launch_bounds__(kThreads, 1) small_work_kernel(const ScanParams P, const SmallParams Q)
{
#if defined(ADSB_HOST_EMU)
    unsigned char *smem_raw = cuda_emu::dynamic_smem();
#else
    extern __shared__ __align__(16) unsigned char smem_raw[];
#endif
    using G = Geo<SPS, kThreads>;
    const TileSmem sm = carve(smem_raw, G::T, SPS ? G::RH : P.rh, G::CW);
    const long long tile = (long long)P.tile_first + blockIdx.x;
    const int tid = threadIdx.x;
    const CtaBar gb;
#if ADSB_SMALL_TIMING
    unsigned long long tstamp[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#endif
    ADSB_TSTAMP(0);
    scan_tile<SPS, FMT, kThreads, CtaBar, 30>(P, sm, tile, gb);   // 30 rows in flight: a whole tile (22-29 rows) in one batch
    gb.sync();
    ADSB_TSTAMP(2);
    slice_tile<SPS, FMT, kThreads>(P, sm, tile, tid, kWarps, gb);
    ADSB_TSTAMP(3);
    if (Q.mag_out) {
        // the stream blocks.complex_to_mag_squared fed downstream (adsb_rx.py:180): this tile's share, from shared memory
        const long long w_first = tile * G::T + P.lead;          // window index of shared-memory slot kLeftHalo
        // the last tile also hands out its right halo: the 8*sps-1 look-ahead items no tile owns
        const int lim = (tile == P.n_tiles_total - 1) ? G::T + (SPS ? G::RH : P.rh) : G::T;
        for (int i = tid; i < lim; i += kThreads) {
            const long long o = w_first + i - Q.mag_first;
            if (o >= 0 && o < Q.mag_n) Q.mag_out[o] = sm.mag[kLeftHalo + i];
        }
    }

    // ---- the CTA that finishes last resolves the call ----
    __shared__ int s_last;
    __threadfence();
    gb.sync();
    if (tid == 0) s_last = (atomicAdd(&P.hdr->tiles_done, 1u) == (unsigned)(Q.n_tiles - 1));
    gb.sync();
    if (!s_last) return;
    __threadfence();
    ADSB_TSTAMP(4);

    // shared memory is free now: candidate centres (relative to base), their record slots, accepted flags
    int *s_idx = reinterpret_cast<int *>(smem_raw);
    unsigned *s_src = reinterpret_cast<unsigned *>(smem_raw) + kSmallCand;
    unsigned char *s_acc = smem_raw + 8 * kSmallCand;
    __shared__ int scan_scratch[16];
    __shared__ unsigned s_cnt[4];       // hits, demod, crc
    const volatile uint2 *tt = reinterpret_cast<const volatile uint2 *>(P.tile_tab);
    unsigned my_off = 0, my_cnt = 0;
    if (tid < Q.n_tiles) {
        my_off = tt[tid].x;
        my_cnt = tt[tid].y;
    }
    int M;
    const int excl = block_excl_scan((int)my_cnt, scan_scratch, M);
    Header *h = P.hdr;
    SmallOut *out = Q.out;
    if (tid < 3) s_cnt[tid] = 0;
    bool overflow = M > kSmallCand;
    if (!overflow) {
        // ascending order: tiles in order, a tile's records in order (slice_tile)
        for (unsigned k = 0; k < my_cnt; k++) {
            const unsigned src = my_off + k;
            if (src < P.cand_cap) {
                s_idx[excl + k] = (int)(*reinterpret_cast<const volatile long long *>(&P.cand[src].idx) - P.base);
                s_src[excl + k] = src;
            } else {
                s_idx[excl + k] = 0x7fffffff;
                s_src[excl + k] = 0xffffffffu;
            }
        }
        __syncthreads();
        if (tid == 0) {
            // framer.py:121-123,165: a candidate is taken iff it lies beyond the blanking interval of the one before
            long long lim = Q.eob_in;
            for (int i = 0; i < M; i++) {
                const long long c = s_idx[i];
                const bool acc = c > lim && s_src[i] != 0xffffffffu;
                s_acc[i] = acc ? 1 : 0;
                if (acc) lim = c + Q.blank;
            }
        }
        __syncthreads();
        // accepted candidates -> hit records in the pinned result block, in order
        int done = 0;
        long long last_acc = -1;
        for (int i0 = 0; i0 < M; i0 += kThreads) {
            const int i = i0 + tid;
            const int a = (i < M) ? s_acc[i] : 0;
            int total;
            const int pos = done + block_excl_scan(a, scan_scratch, total);
            if (a) {
                if (pos < kSmallHits) {
                    Rec r;
                    const uint4 *src = reinterpret_cast<const uint4 *>(P.cand + s_src[i]);
                    reinterpret_cast<uint4 *>(&r)[0] = ld_cg_u4(src);          // written by other CTAs of this launch
                    reinterpret_cast<uint4 *>(&r)[1] = ld_cg_u4(src + 1);
                    r.snr = snr_db_from_ratio(r.snr);           // K1 left x[p]/median here
                    if (r.flags & kHitDemodValid) atomicAdd(&s_cnt[1], 1u);
                    if (r.flags & kHitCrcOk) atomicAdd(&s_cnt[2], 1u);
                    uint4 *dst = reinterpret_cast<uint4 *>(out->hits + pos);
                    dst[0] = reinterpret_cast<const uint4 *>(&r)[0];
                    dst[1] = reinterpret_cast<const uint4 *>(&r)[1];
                } else {
                    overflow = true;
                }
                last_acc = s_idx[i];
            }
            done += total;
        }
        // the largest accepted centre: candidates are ascending, so it is the last accepted one
        __shared__ long long s_last_acc[kWarps];
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            const long long o = __shfl_xor_sync(kFull, last_acc, d);
            last_acc = o > last_acc ? o : last_acc;
        }
        if ((tid & 31) == 0) s_last_acc[tid >> 5] = last_acc;
        overflow = __syncthreads_or(overflow) != 0;
        if (tid == 0) {
            long long la = -1;
#pragma unroll
            for (int q = 0; q < kWarps; q++) la = s_last_acc[q] > la ? s_last_acc[q] : la;
            out->hdr.n_hits = (unsigned long long)done;
            out->hdr.n_demod = s_cnt[1];
            out->hdr.n_crc = s_cnt[2];
            out->hdr.last_accept = la;
            out->hdr.first_cand = M > 0 ? (long long)s_idx[0] : -1;
        }
    } else {
        __syncthreads();
    }
    if (tid == 0) {
        const volatile Header *vh = h;
        out->hdr.n_cand = (unsigned long long)M;
        out->hdr.last_pulse = vh->last_pulse;
        out->hdr.flags = vh->flags | (overflow ? kFlagSmallOverflow : 0u);
        out->hdr.x_last = vh->x_last;
        out->hdr.n_fec = 0;
        out->hdr.carry_unstable = 0;
        if (overflow) {
            // too many candidates / hits for this path: leave the counter the resolve kernels expect and let the host run them
            h->n_cand = (unsigned long long)M;
            h->tiles_done = 0;
        } else {
            // leave the device header as init_header_kernel would: the next small call starts without one
            h->n_cand = 0;
            h->n_hits = 0;
            h->n_demod = 0;
            h->n_crc = 0;
            h->last_pulse = -1;
            h->last_accept = -1;
            h->flags = 0;
            h->n_fec = 0;
            h->tiles_done = 0;
            h->x_last = 0.0f;
            h->first_cand = -1;
            P.tile_chunk_tot[0] = 0;        // the tiles of a small call all fall into the first gather chunk
            h->carry_ein = kCarryNone;
            h->carry_used = kCarryNone;
            h->carry_need = 0;
        }
#if ADSB_SMALL_TIMING
        tstamp[5] = now_ns();
        for (int k = 0; k < 8; k++) out->t[k] = tstamp[k];
#endif
        fence_system();
        *reinterpret_cast<volatile unsigned long long *>(&out->seq) = Q.seq;
    }
}

// stand-alone demod for a handful of tags (demod.py:75-95), samples and results in pinned host memory, one launch
__global__ void __launch_bounds__(kThreads) small_demod_kernel(const float *__restrict__ in0, long long lo, long long n,
                                                                const long long *__restrict__ sob, int n_tags, int sps,
                                                                unsigned char *bits, unsigned char *valid,
                                                                unsigned long long *seq_out, unsigned long long seq,
                                                                unsigned *done_ctr)
{
    const int lane = threadIdx.x & 31;
    const int h = sps >> 1;
    for (int t = blockIdx.x * kWarps + (threadIdx.x >> 5); t < n_tags; t += gridDim.x * kWarps) {
        const long long s = sob[t];
        // eob_idx = sob_idx + 111*sps + sps/2 must be < len(in0) (demod.py:76,80,82); in0[i] is staged at in0[i - lo]
        const bool ok = (s >= 0) && (s + 111LL * sps + h < n);
        if (lane == 0) valid[t] = ok ? 1 : 0;
        if (!ok) continue;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int b = j * 32 + lane;
            if (b < kMsgBits)
                bits[(long long)t * kMsgBits + b] = in0[s - lo + (long long)b * sps] > in0[s - lo + (long long)b * sps + h];
        }
    }
    __shared__ int s_lastb;
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) s_lastb = (atomicAdd(done_ctr, 1u) == gridDim.x - 1);
    __syncthreads();
    if (s_lastb && threadIdx.x == 0) {
        *done_ctr = 0;
        fence_system();
        *reinterpret_cast<volatile unsigned long long *>(seq_out) = seq;
    }
}

}  // namespace adsb
