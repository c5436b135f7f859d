// K2: resolve the sparse candidate list produced by K1.
//   gather_kernel   tile table (offset,count) -> candidates in ascending sample order
//   chain_kernel    63-symbol blanking (framer.py:121-123,165): greedy chain, parallel over the
//                   independent segments that gaps > 63 symbols cut the candidate list into
//   compact_kernel  accepted candidates -> hit list (ordered), counters
//   demod_kernel    stand-alone PPM slice for the demod block (demod.py:82-95)
#pragma once
#include "adsb_device.cuh"

namespace adsb {

// Ordered offsets without inter-block waiting: per-chunk totals are accumulated with fire-and-forget
// atomics by the producer (K1 for tiles, chain_kernel for accepted candidates); a consumer block sums the
// totals of the chunks before its own.
__device__ __forceinline__ unsigned block_sum_before(const unsigned *tot, unsigned n, unsigned *scratch)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    unsigned v = 0;
    for (unsigned i = threadIdx.x; i < n; i += kThreads) v += tot[i];
    v = __reduce_add_sync(kFull, v);
    if (lane == 0) scratch[wid] = v;
    __syncthreads();
    unsigned t = 0;
#pragma unroll
    for (int q = 0; q < kWarps; q++) t += scratch[q];
    __syncthreads();
    return t;
}

struct ResolveParams {
    const uint2 *tile_tab;
    long long n_tiles;
    const Rec *cand;            // K1 order (tile blocks at atomic offsets)
    Rec *ord;                   // ascending order
    Rec *hits;
    unsigned hit_cap;
    unsigned cand_cap;
    unsigned *tile_chunk_tot;   // candidates per gather chunk (kTilesPerChunk tiles), filled by K1
    unsigned *cand_chunk_tot;   // accepted candidates per compact chunk (kCandPerChunk), filled by chain_kernel
    Header *hdr;
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
    __shared__ unsigned uscratch[kWarps];
    const int tid = threadIdx.x, lane = tid & 31;
    const unsigned n_chunks = (unsigned)((P.n_tiles + kTilesPerChunk - 1) / kTilesPerChunk);
    for (unsigned chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x) {
        const long long t0 = (long long)chunk * kTilesPerChunk + tid;
        const uint2 tt = (t0 < P.n_tiles) ? P.tile_tab[t0] : make_uint2(0u, 0u);
        int total;
        const int off = block_excl_scan((int)tt.y, scratch, total);
        const unsigned excl = block_sum_before(P.tile_chunk_tot, chunk, uscratch);
        // the last chunk knows the total: K1 only used hdr->n_cand as its overflow allocator
        if (chunk == n_chunks - 1 && tid == 0) P.hdr->n_cand = (unsigned long long)excl + (unsigned long long)total;
        const unsigned dst = excl + (unsigned)off;
        // warp-cooperative copy, flattened over the warp's 32 tiles so that every lane has an independent
        // 16-byte load in flight each iteration (a per-tile loop serialises on the load latency)
        __shared__ unsigned s_pre[kWarps][33], s_src[kWarps][32], s_dst[kWarps][32];
        const int wid = tid >> 5;
        unsigned inc = tt.y;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned t = __shfl_up_sync(kFull, inc, d);
            if (lane >= d) inc += t;
        }
        s_pre[wid][lane + 1] = inc;                  // inclusive prefix of the warp's tile counts
        if (lane == 0) s_pre[wid][0] = 0;
        s_src[wid][lane] = tt.x;
        s_dst[wid][lane] = dst;
        __syncwarp();
        const unsigned wtot = s_pre[wid][32];
#pragma unroll 4
        for (unsigned hidx = lane; hidx < 2 * wtot; hidx += 32) {
            const unsigned r = hidx >> 1;
            int t = 0;                               // tile with s_pre[t] <= r < s_pre[t+1]
#pragma unroll
            for (int step = 16; step > 0; step >>= 1)
                if (s_pre[wid][t + step] <= r) t += step;
            const unsigned k = r - s_pre[wid][t];
            const unsigned src = s_src[wid][t] + k, dd = s_dst[wid][t] + k;
            if (src < P.cand_cap && dd < P.cand_cap)
                reinterpret_cast<uint4 *>(P.ord + dd)[hidx & 1] = reinterpret_cast<const uint4 *>(P.cand + src)[hidx & 1];
        }
        __syncwarp();
    }
}

// ---- chain: blanking -------------------------------------------------------------------------------
// A candidate is eligible when its centre lies beyond the incoming prev_eob_idx (framer.py:121).  An
// eligible candidate whose predecessor is ineligible or more than 63 symbols behind starts a segment
// and is always accepted; its thread then walks the segment greedily.  Accepted = flags bit 31.
constexpr unsigned kAccepted = 0x80000000u;
constexpr int kCandPerThread = 4;
constexpr int kCandPerChunk = kThreads * kCandPerThread;   // candidates per compact_kernel block iteration

__global__ void __launch_bounds__(kThreads) chain_kernel(const ResolveParams P)
{
    if (P.dev_carry && !P.hdr->carry_need) return;
    const unsigned long long m64 = P.hdr->n_cand;
    const long long M = (long long)(m64 < P.cand_cap ? m64 : P.cand_cap);
    const long long e_in = P.dev_carry ? P.hdr->carry_ein : P.base + P.eob_in;
    const long long D = P.blank;
    const int lane = threadIdx.x & 31;
    // every record's accepted bit is assigned exactly once per launch (by the record's own thread when
    // it is ineligible, otherwise by the head of its segment), so the kernel can be re-run with
    // another eob_in (multi-GPU carry fix-up).  Accepted candidates are counted per compact chunk: a warp's
    // 32 consecutive candidates share one chunk, so the count goes out as one atomic per warp.
    for (long long base = (long long)blockIdx.x * blockDim.x; base < M; base += (long long)gridDim.x * blockDim.x) {
        const long long i = base + threadIdx.x;
        const long long my_chunk = i / kCandPerChunk;
        unsigned local = 0;
        if (i < M) {
            const long long c = P.ord[i].idx;
            if (i == 0) P.hdr->first_cand = c - P.base;
            if (c <= e_in) {                            // framer.py:121: still blanked by the incoming state
                P.ord[i].flags &= ~kAccepted;
            } else {
                bool head = (i == 0);
                if (!head) {
                    const long long cp = P.ord[i - 1].idx;
                    head = (cp <= e_in) || (c - cp > D);
                }
                if (head) {
                    long long last = c, prev = c;
                    P.ord[i].flags |= kAccepted;
                    local++;
                    for (long long j = i + 1; j < M; j++) {
                        const long long cj = P.ord[j].idx;
                        if (cj - prev > D) break;               // next segment (its own head handles it)
                        prev = cj;
                        if (cj > last + D) {                    // framer.py:121 with prev_eob_idx = last + 63*sps
                            P.ord[j].flags |= kAccepted;
                            last = cj;
                            if (j / kCandPerChunk == my_chunk) local++;
                            else atomicAdd(&P.cand_chunk_tot[j / kCandPerChunk], 1u);
                        } else {
                            P.ord[j].flags &= ~kAccepted;
                        }
                    }
                }
            }
        }
        const unsigned wsum = __reduce_add_sync(kFull, local);
        if (lane == 0 && wsum) atomicAdd(&P.cand_chunk_tot[(base + (threadIdx.x & ~31)) / kCandPerChunk], wsum);
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
    h->n_fec = 0;
}

// ---- compact: accepted candidates -> ordered hit list ------------------------------------------------

__global__ void __launch_bounds__(kThreads) compact_kernel(const ResolveParams P)
{
    __shared__ int scratch[16];
    __shared__ unsigned uscratch[kWarps];
    const int tid = threadIdx.x, lane = tid & 31;
    if (P.dev_carry && !P.hdr->carry_need) return;
    const unsigned long long m64 = P.hdr->n_cand;
    const long long M = (long long)(m64 < P.cand_cap ? m64 : P.cand_cap);
    const unsigned n_chunks = (unsigned)((M + kCandPerChunk - 1) / kCandPerChunk);
    for (unsigned chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x) {
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
        const unsigned excl = block_sum_before(P.cand_chunk_tot, chunk, uscratch);
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
        unsigned dst = excl + (unsigned)off;
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

// ---- multi-GPU peer mode: neighbour hand-shake and blanking carry through peer-mapped control blocks ----------
// Every rank owns one ShardCtrl in its symmetric allocation (adsb_shard_alloc); the other ranks write into it over
// NVLink.  All waits are bounded (kPeerTimeoutNs) and report through Header.flags instead of hanging the GPU.
constexpr int kMaxWorld = 32;

struct ShardCtrl {
    unsigned long long ready[2];      // step counter of the left [0] / right [1] neighbour, written by that neighbour
    unsigned long long step;          // this rank's step counter
    unsigned long long xchg;          // this rank's carry-exchange counter (one per round-0 export; stamps the slots)
    unsigned long long checked;       // the latest exchange whose outcome (round 1) this rank has evaluated
    unsigned long long pad[3];
    long long slot[2][kMaxWorld][4];  // [round][rank] = first candidate, last hit, incoming state assumed, step stamp
};
static_assert(sizeof(ShardCtrl) <= 8192, "control block must fit ADSB_SHARD_CTRL_BYTES");

struct PeerCtrls {
    ShardCtrl *p[kMaxWorld];
};

// signal: bump the own step counter and tell both neighbours that this rank's samples of the step are resident.
// wait: hold the stream until both neighbours have said the same (the scan kernel's edge tiles do this wait
// themselves -- wait_neighbours() in adsb_scan.cuh -- so the production step only launches the signal).  One thread.
__global__ void shard_sync_kernel(ShardCtrl *own, ShardCtrl *left, ShardCtrl *right, Header *hdr, int do_signal, int do_wait)
{
    if (do_signal) {
        const unsigned long long k = own->step + 1;
        own->step = k;
        fence_system();                 // the samples written earlier on this stream are visible before the flag is
        if (left) *reinterpret_cast<volatile unsigned long long *>(&left->ready[1]) = k;
        if (right) *reinterpret_cast<volatile unsigned long long *>(&right->ready[0]) = k;
    }
    if (do_wait) {
        const unsigned long long k = own->step;
        bool ok = true;
        if (left) ok = wait_ge(&own->ready[0], k) && ok;
        if (right) ok = wait_ge(&own->ready[1], k) && ok;
        fence_system();
        if (!ok) atomicOr(&hdr->flags, kFlagPeerTimeout);
    }
}

// Blanking carry without NCCL and without the host.  export: this rank's (first candidate, last hit, incoming state its
// latest resolve assumed) goes into slot[round][rank] of EVERY rank's control block, stamped with the step counter.
// plan: wait until all `world` slots of the own block carry this step's stamp, then evaluate the same plan as
// gr_adsb_b200/sharding.py:carry_plan.  round 0 arms the conditional re-resolve (chain + compact) that follows on the
// stream; round 1 only checks that nothing is left to fix (Header.carry_unstable otherwise).
__device__ __forceinline__ bool carry_plan_any(const ShardCtrl *own, int round, int world, int rank, long long blank,
                                               long long &my_e, int &my_need)
{
    long long e = kCarryNone;
    int any = 0;
    my_e = kCarryNone;
    my_need = 0;
    for (int r = 0; r < world; r++) {
        const volatile long long *sl = own->slot[round][r];
        const long long first = sl[0], last = sl[1], used = sl[2];
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
    return any != 0;
}

// round 0: do_export publishes this rank's summary, do_plan waits for everybody's and arms the conditional re-resolve.
// round 1: do_export re-publishes after the fix-ups; the CHECK that nothing is left to fix (Header.carry_unstable
// otherwise) does not sit on the step's critical path: it is made at the start of the NEXT exchange (do_prev_check, before
// this rank publishes again -- nobody can overwrite a round-1 slot before every rank has published the next round 0), or
// by an explicit round-1 do_plan call (the end of a run, tests).
// wait for every rank's round-1 slot of exchange x and count the exchange as unstable if some rank still needs a fix-up
__device__ __forceinline__ void carry_check(Header *h, ShardCtrl *own, int world, int rank, long long blank,
                                            unsigned long long x, int *s_bad)
{
    const int t = threadIdx.x;
    if (t == 0) *s_bad = 0;
    __syncthreads();
    if (t < world && !wait_ge(reinterpret_cast<const unsigned long long *>(&own->slot[1][t][3]), x)) *s_bad = 1;
    fence_system();
    __syncthreads();
    if (t == 0) {
        if (*s_bad) {
            atomicOr(&h->flags, kFlagPeerTimeout);
        } else {
            long long e_;
            int n_;
            if (carry_plan_any(own, 1, world, rank, blank, e_, n_)) h->carry_unstable += 1;
        }
        own->checked = x;
    }
    __syncthreads();
}

__global__ void __launch_bounds__(kMaxWorld) shard_carry_kernel(Header *h, ShardCtrl *own, const PeerCtrls peers, int world,
                                                                 int rank, long long base, long long blank, int round,
                                                                 int do_export, int do_plan, int do_prev_check)
{
    const int t = threadIdx.x;
    __shared__ unsigned long long s_k, s_chk;
    __shared__ int s_bad;
    if (t == 0) {
        s_k = own->xchg;
        s_chk = own->checked;
    }
    __syncthreads();
    // (A) the outcome of the previous exchange, looked at before this rank publishes again
    const unsigned long long x0 = s_k, c0 = s_chk;
    if (do_prev_check && round == 0 && x0 > c0) carry_check(h, own, world, rank, blank, x0, &s_bad);
    __syncthreads();            // every thread has read the cells thread 0 rewrites next
    // (B) publish; every rank runs the same sequence of exchanges, so the counters agree without being communicated
    if (t == 0) {
        unsigned long long x = own->xchg;
        if (do_export && round == 0) own->xchg = ++x;
        s_k = x;
        s_chk = own->checked;
        s_bad = 0;
    }
    __syncthreads();
    const unsigned long long k = s_k, chk = s_chk;
    if (do_export && t < world) {
        long long *dst = peers.p[t]->slot[round][rank];
        volatile long long *vd = dst;
        vd[0] = h->first_cand >= 0 ? base + h->first_cand : kCarryNone;
        vd[1] = h->last_accept >= 0 ? base + h->last_accept : kCarryNone;
        vd[2] = h->carry_used;
        fence_system();
        vd[3] = (long long)k;
    }
    if (!do_plan) return;
    if (round == 1) {
        // (C1) explicit check of this exchange (end of a run, tests)
        if (k > chk) carry_check(h, own, world, rank, blank, k, &s_bad);
        if (t == 0) h->carry_need = 0;
        return;
    }
    // (C0) wait for everybody's round-0 summary and plan this rank's fix-up
    if (t < world) {
        if (!wait_ge(reinterpret_cast<const unsigned long long *>(&own->slot[0][t][3]), k)) s_bad = 1;
        fence_system();
    }
    __syncthreads();
    if (t != 0) return;
    if (s_bad) {
        atomicOr(&h->flags, kFlagPeerTimeout);
        h->carry_need = 0;
        return;
    }
    long long my_e;
    int my_need;
    carry_plan_any(own, 0, world, rank, blank, my_e, my_need);
    h->carry_need = my_need;
    h->carry_ein = my_e;
    h->carry_used = my_e;       // where no re-run is needed the result is the same under either value
    if (my_need) {              // what reset_resolve_kernel does, for the conditional chain + compact that follow
        h->n_hits = 0;
        h->n_demod = 0;
        h->n_crc = 0;
        h->last_accept = -1;
        h->n_fec = 0;
    }
}

// ---- "Conservative" error correction (decoder.py:304-323, 738-763): one thread per hit -----------------------
// A hit that was demodulated, fails the parity check and carries DF 17..19 gets its 25-bit residual
// (decoder.compute_crc_2: the division stops one step early) looked up among the residuals of all 1-bit and
// 2-contiguous-bit error patterns; a match flips those bits and the message counts as passed.
constexpr int kFecPatterns = 112 + 111;

struct FecParams {
    Rec *hits;
    Header *hdr;
    unsigned hit_cap;
    const unsigned *t25;            // [112] residual25 of the unit vector at bit i
    const unsigned *keys;           // [kFecPatterns] residuals of the error patterns, ascending
    const unsigned short *fix;      // [kFecPatterns] first bit | (length << 8), same order
};

__global__ void __launch_bounds__(kThreads) fec_kernel(const FecParams P)
{
    const unsigned long long n64 = P.hdr->n_hits;
    const unsigned n = (unsigned)(n64 < P.hit_cap ? n64 : P.hit_cap);
    unsigned fixed = 0;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const unsigned flags = P.hits[i].flags;
        if (!(flags & kHitDemodValid) || (flags & kHitCrcOk)) continue;
        unsigned w[4];
#pragma unroll
        for (int j = 0; j < 4; j++) w[j] = P.hits[i].bits[j];
        // message bit k sits in byte k/8 (little-endian words), bit 7 - k%8
        const unsigned df = (w[0] & 0xffu) >> 3;
        if (df < 17 || df > 19) continue;                       // payload_length never set (decoder.py:667-670)
        unsigned r = 0;
        for (int k = 0; k < kMsgBits; k++) {
            const unsigned bit = (w[k >> 5] >> (8 * ((k >> 3) & 3) + 7 - (k & 7))) & 1u;
            r ^= bit ? P.t25[k] : 0u;
        }
        int lo = 0, hi = kFecPatterns - 1, found = -1;
        while (lo <= hi) {
            const int mid = (lo + hi) >> 1;
            const unsigned key = P.keys[mid];
            if (key == r) { found = mid; break; }
            if (key < r) lo = mid + 1; else hi = mid - 1;
        }
        if (found < 0) continue;
        const unsigned f = P.fix[found];
        for (unsigned q = 0; q < (f >> 8); q++) {
            const unsigned k = (f & 0xffu) + q;
            w[k >> 5] ^= 1u << (8 * ((k >> 3) & 3) + 7 - (k & 7));
        }
#pragma unroll
        for (int j = 0; j < 4; j++) P.hits[i].bits[j] = w[j];
        P.hits[i].flags = flags | kHitCrcOk | kHitFecFixed;
        fixed++;
    }
    fixed = __reduce_add_sync(kFull, fixed);
    if ((threadIdx.x & 31) == 0 && fixed) atomicAdd(&P.hdr->n_fec, fixed);
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
