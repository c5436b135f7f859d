// K2: resolve the sparse candidate list produced by K1.
//   gather_kernel   tile table (offset,count) -> candidates in ascending sample order
//   chain_kernel    63-symbol blanking (framer.py:121-123,165): greedy chain, parallel over the
//                   independent segments that gaps > 63 symbols cut the candidate list into
//   compact_kernel  accepted candidates -> hit list (ordered), counters
//   demod_kernel    stand-alone PPM slice for the demod block (demod.py:82-95)
#pragma once
#include "adsb_device.cuh"

namespace adsb {

// ---- decoupled look-back over chunk aggregates (one warp) -------------------------------------------
// status word: high 32 bits = (epoch << 2) | flag, low 32 bits = value; flag 1 = aggregate, 2 = inclusive.
__device__ __forceinline__ unsigned long long pack_status(unsigned epoch, unsigned flag, unsigned v)
{
    return ((unsigned long long)((epoch << 2) | flag) << 32) | v;
}

__device__ __forceinline__ unsigned warp_sum_u32(unsigned v)
{
    return __reduce_add_sync(kFull, v);
}

// called by all 32 lanes of one warp; returns the exclusive prefix of `agg` over chunks < idx
__device__ unsigned lookback_exclusive(volatile unsigned long long *st, unsigned idx, unsigned agg,
                                       unsigned epoch, unsigned *err_flags)
{
    const int lane = threadIdx.x & 31;
    if (idx == 0) {
        if (lane == 0) st[0] = pack_status(epoch, 2u, agg);
        return 0u;
    }
    if (lane == 0) st[idx] = pack_status(epoch, 1u, agg);
    unsigned excl = 0;
    long long base = (long long)idx - 1;
    unsigned spins = 0;
    while (true) {
        const long long t = base - lane;
        unsigned long long s = 0;
        bool ready = true, inc = true;
        if (t >= 0) {
            s = st[t];
            const unsigned hi = (unsigned)(s >> 32);
            ready = ((hi >> 2) == epoch) && ((hi & 3u) != 0u);
            inc = ready && ((hi & 3u) == 2u);
        }
        const unsigned notready = __ballot_sync(kFull, !ready);
        const unsigned usable = notready ? ((1u << (__ffs(notready) - 1)) - 1u) : kFull;
        const unsigned incmask = __ballot_sync(kFull, inc) & usable;
        if (incmask) {
            const int first = __ffs(incmask) - 1;
            excl += warp_sum_u32((lane <= first && t >= 0) ? (unsigned)s : 0u);
            break;
        }
        if (usable == kFull) {
            excl += warp_sum_u32((unsigned)s);
            base -= 32;
            continue;
        }
        if (++spins > (1u << 22)) {
            if (lane == 0) atomicOr(err_flags, kFlagSpin);
            break;
        }
        __nanosleep(40);
    }
    if (lane == 0) st[idx] = pack_status(epoch, 2u, excl + agg);
    return excl;
}

struct ResolveParams {
    const uint2 *tile_tab;
    long long n_tiles;
    const Rec *cand;            // K1 order (tile blocks at atomic offsets)
    Rec *ord;                   // ascending order
    Rec *hits;
    unsigned hit_cap;
    unsigned cand_cap;
    volatile unsigned long long *status_g;
    volatile unsigned long long *status_c;
    Header *hdr;
    unsigned epoch;
    long long base;             // records carry base + window index
    long long eob_in;           // incoming prev_eob_idx, window coordinates (framer.py:57,121)
    int blank;                  // 63 * sps (framer.py:165)
    int dev_carry;              // 1: conditional re-run -- take eob from hdr->carry_ein, do nothing unless hdr->carry_need
};

// ---- gather: order the tile blocks -------------------------------------------------------------------
constexpr int kTilesPerChunk = kThreads;      // one tile per thread: enough blocks to fill the GPU

__global__ void __launch_bounds__(kThreads) gather_kernel(const ResolveParams P)
{
    __shared__ int scratch[16];
    __shared__ unsigned s_chunk, s_excl;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const unsigned n_chunks = (unsigned)((P.n_tiles + kTilesPerChunk - 1) / kTilesPerChunk);
    while (true) {
        if (tid == 0) s_chunk = atomicAdd(&P.hdr->chunk_ticket_g, 1u);
        __syncthreads();
        const unsigned chunk = s_chunk;
        if (chunk >= n_chunks) break;
        const long long t0 = (long long)chunk * kTilesPerChunk + tid;
        const uint2 tt = (t0 < P.n_tiles) ? P.tile_tab[t0] : make_uint2(0u, 0u);
        int total;
        const int off = block_excl_scan((int)tt.y, scratch, total);
        if (wid == 0) {
            unsigned e = lookback_exclusive(P.status_g, chunk, (unsigned)total, P.epoch, &P.hdr->flags);
            if (lane == 0) s_excl = e;
        }
        __syncthreads();
        // the last chunk knows the total: K1 only used hdr->n_cand as its overflow allocator
        if (chunk == n_chunks - 1 && tid == 0) P.hdr->n_cand = (unsigned long long)s_excl + (unsigned long long)total;
        const unsigned dst = s_excl + (unsigned)off;
        // warp-cooperative copy: the 32 tiles of a warp one after the other, 16 bytes per lane
        const unsigned any = __ballot_sync(kFull, tt.y != 0);
        for (unsigned m = any; m; m &= m - 1) {
            const int t = __ffs(m) - 1;
            const unsigned src_t = __shfl_sync(kFull, tt.x, t);
            const unsigned cnt_t = __shfl_sync(kFull, tt.y, t);
            const unsigned dst_t = __shfl_sync(kFull, dst, t);
            for (unsigned i = lane; i < 2 * cnt_t; i += 32) {
                const unsigned r = i >> 1;
                if (src_t + r < P.cand_cap && dst_t + r < P.cand_cap)
                    reinterpret_cast<uint4 *>(P.ord + dst_t + r)[i & 1] = reinterpret_cast<const uint4 *>(P.cand + src_t + r)[i & 1];
            }
        }
        __syncthreads();
    }
}

// ---- chain: blanking -------------------------------------------------------------------------------
// A candidate is eligible when its centre lies beyond the incoming prev_eob_idx (framer.py:121).  An
// eligible candidate whose predecessor is ineligible or more than 63 symbols behind starts a segment
// and is always accepted; its thread then walks the segment greedily.  Accepted = flags bit 31.
constexpr unsigned kAccepted = 0x80000000u;

__global__ void __launch_bounds__(kThreads) chain_kernel(const ResolveParams P)
{
    if (P.dev_carry && !P.hdr->carry_need) return;
    const unsigned long long m64 = P.hdr->n_cand;
    const long long M = (long long)(m64 < P.cand_cap ? m64 : P.cand_cap);
    const long long e_in = P.dev_carry ? P.hdr->carry_ein : P.base + P.eob_in;
    const long long D = P.blank;
    // every record's accepted bit is assigned exactly once per launch (by the record's own thread when
    // it is ineligible, otherwise by the head of its segment), so the kernel can be re-run with
    // another eob_in (multi-GPU carry fix-up)
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < M;
         i += (long long)gridDim.x * blockDim.x) {
        const long long c = P.ord[i].idx;
        if (i == 0) P.hdr->first_cand = c - P.base;
        if (c <= e_in) {                            // framer.py:121: still blanked by the incoming state
            P.ord[i].flags &= ~kAccepted;
            continue;
        }
        bool head = (i == 0);
        if (!head) {
            const long long cp = P.ord[i - 1].idx;
            head = (cp <= e_in) || (c - cp > D);
        }
        if (!head) continue;
        long long last = c, prev = c;
        P.ord[i].flags |= kAccepted;
        for (long long j = i + 1; j < M; j++) {
            const long long cj = P.ord[j].idx;
            if (cj - prev > D) break;               // next segment (its own head handles it)
            prev = cj;
            if (cj > last + D) {                    // framer.py:121 with prev_eob_idx = last + 63*sps
                P.ord[j].flags |= kAccepted;
                last = cj;
            } else {
                P.ord[j].flags &= ~kAccepted;
            }
        }
    }
}

// reset what compact_kernel accumulates, before a re-run of chain + compact
__global__ void reset_resolve_kernel(Header *h, int dev_carry)
{
    if (dev_carry && !h->carry_need) return;
    h->n_hits = 0;
    h->n_demod = 0;
    h->n_crc = 0;
    h->last_accept = -1;
    h->chunk_ticket_c = 0;
}

// ---- compact: accepted candidates -> ordered hit list ------------------------------------------------
constexpr int kCandPerThread = 4;
constexpr int kCandPerChunk = kThreads * kCandPerThread;   // big chunks keep the look-back chain short

__global__ void __launch_bounds__(kThreads) compact_kernel(const ResolveParams P)
{
    __shared__ int scratch[16];
    __shared__ unsigned s_chunk, s_excl;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (P.dev_carry && !P.hdr->carry_need) return;
    const unsigned long long m64 = P.hdr->n_cand;
    const long long M = (long long)(m64 < P.cand_cap ? m64 : P.cand_cap);
    const unsigned n_chunks = (unsigned)((M + kCandPerChunk - 1) / kCandPerChunk);
    while (true) {
        if (tid == 0) s_chunk = atomicAdd(&P.hdr->chunk_ticket_c, 1u);
        __syncthreads();
        const unsigned chunk = s_chunk;
        if (chunk >= n_chunks) break;
        const long long i0 = (long long)chunk * kCandPerChunk + (long long)tid * kCandPerThread;
        Rec r[kCandPerThread];
        int acc = 0;
        unsigned n_demod = 0, n_crc = 0;
        long long mx = -1;
#pragma unroll
        for (int q = 0; q < kCandPerThread; q++) {
            r[q].flags = 0;
            r[q].idx = -1;
            if (i0 + q < M) {
                const uint4 *a = reinterpret_cast<const uint4 *>(P.ord + i0 + q);
                reinterpret_cast<uint4 *>(&r[q])[0] = a[0];
                reinterpret_cast<uint4 *>(&r[q])[1] = a[1];
            }
            if (r[q].flags & kAccepted) {
                acc++;
                n_demod += (r[q].flags & kHitDemodValid) ? 1u : 0u;
                n_crc += (r[q].flags & kHitCrcOk) ? 1u : 0u;
                mx = r[q].idx;                              // ascending: the last accepted one is the largest
            }
        }
        int total;
        const int off = block_excl_scan(acc, scratch, total);
        if (wid == 0) {
            unsigned e = lookback_exclusive(P.status_c, chunk, (unsigned)total, P.epoch, &P.hdr->flags);
            if (lane == 0) s_excl = e;
        }
        // block counters
        n_demod = __reduce_add_sync(kFull, n_demod);
        n_crc = __reduce_add_sync(kFull, n_crc);
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            const long long o = __shfl_xor_sync(kFull, mx, d);
            mx = o > mx ? o : mx;
        }
        if (lane == 0) {
            if (n_demod) atomicAdd(&P.hdr->n_demod, (unsigned long long)n_demod);
            if (n_crc) atomicAdd(&P.hdr->n_crc, (unsigned long long)n_crc);
            if (mx >= 0) atomicMax(&P.hdr->last_accept, mx - P.base);
        }
        __syncthreads();
        unsigned dst = s_excl + (unsigned)off;
#pragma unroll
        for (int q = 0; q < kCandPerThread; q++) {
            if (r[q].flags & kAccepted) {
                if (dst < P.hit_cap) {
                    r[q].flags &= ~kAccepted;
                    r[q].snr = snr_db_from_ratio(r[q].snr);       // K1 left x[p]/median here
                    uint4 *b = reinterpret_cast<uint4 *>(P.hits + dst);
                    b[0] = reinterpret_cast<const uint4 *>(&r[q])[0];
                    b[1] = reinterpret_cast<const uint4 *>(&r[q])[1];
                } else {
                    atomicOr(&P.hdr->flags, kFlagHitOverflow);
                }
                dst++;
            }
        }
        if (tid == 0 && total) atomicAdd(&P.hdr->n_hits, (unsigned long long)total);
        __syncthreads();
    }
}

// ---- multi-GPU blanking carry, resolved without a host round trip -----------------------------------------
// gathered[r] = (first candidate, last hit, incoming state assumed) of shard r, global window coordinates,
// kCarryNone where absent.  Same plan as gr_adsb_b200/sharding.py:carry_plan, evaluated by every rank.
__global__ void carry_plan_kernel(Header *h, const long long *gathered, int world, int rank, long long blank,
                                  int final_check)
{
    long long e = kCarryNone, my_e = kCarryNone;
    int my_need = 0, any = 0;
    for (int r = 0; r < world; r++) {
        const long long first = gathered[3 * r], last = gathered[3 * r + 1], used = gathered[3 * r + 2];
        const bool changed = e != used;
        const long long mx = e > used ? e : used;
        const bool need = changed && first != kCarryNone && first <= mx;
        if (r == rank) {
            my_e = e;
            my_need = need;
        }
        any |= need;
        if (last != kCarryNone) e = last + blank;
    }
    if (final_check) {
        if (any) h->carry_unstable += 1;
        h->carry_need = 0;
    } else {
        h->carry_need = my_need;
        h->carry_ein = my_e;
        h->carry_used = my_e;       // where no re-run is needed the result is the same under either value
    }
}

// ---- stand-alone demod (demod.py:75-95) ----------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) demod_kernel(const float *__restrict__ in0, long long n,
                                                          const long long *__restrict__ sob, int n_tags, int sps,
                                                          unsigned char *__restrict__ bits,
                                                          unsigned char *__restrict__ valid)
{
    const int lane = threadIdx.x & 31;
    const int h = sps >> 1;
    for (int t = blockIdx.x * kWarps + (threadIdx.x >> 5); t < n_tags; t += gridDim.x * kWarps) {
        const long long s = sob[t];
        // eob_idx = sob_idx + 111*sps + sps/2 must be < len(in0) (demod.py:76,80,82)
        const bool ok = (s >= 0) && (s + 111LL * sps + h < n);
        if (lane == 0) valid[t] = ok ? 1 : 0;
        if (!ok) continue;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int b = j * 32 + lane;
            if (b < kMsgBits)
                bits[(long long)t * kMsgBits + b] = in0[s + (long long)b * sps] > in0[s + (long long)b * sps + h];
        }
    }
}

}  // namespace adsb
