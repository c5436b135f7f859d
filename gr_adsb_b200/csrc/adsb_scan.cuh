// K1: fused |IQ|^2 + threshold/run/centre + preamble match + (per match) SNR, PPM slice, CRC-24.
// One CTA per tile of T samples; the tile plus a 128-sample left halo and a 120-symbol right halo
// is staged once in shared memory as float32 |IQ|^2, so every sample is fetched from HBM once.
#pragma once
#include "adsb_device.cuh"

namespace adsb {

// framer.py:137-147 -- taps x[p + k*h], k = 0..15, against half the centre amplitude; template
// [1,0,1,0,0,0,0,1,0,1,0,0,0,0,0,0] (framer.py:50).  Tests are ordered so that pulses inside a
// burst's data block are rejected after ~2 loads (two consecutive half symbols of one PPM symbol
// cannot both be low).
template <int HS>
__device__ __forceinline__ bool preamble_match(const float *m, int h_rt)
{
    const int h = HS ? HS : h_rt;
    const float xp = m[0];
    const float half = __fmul_rn(xp, 0.5f);      // == xp/2 exactly (framer.py:141)
    if (m[4 * h] > half) return false;
    if (m[5 * h] > half) return false;
    if (m[3 * h] > half) return false;
    if (!(m[2 * h] > half)) return false;
    if (!(m[7 * h] > half)) return false;
    if (!(m[9 * h] > half)) return false;
    if (m[1 * h] > half) return false;
    if (m[6 * h] > half) return false;
    if (m[8 * h] > half) return false;
#pragma unroll
    for (int k = 10; k < 16; k++)
        if (m[k * h] > half) return false;
    return xp > half;
}

// ---- slow path: a pulse whose run is too long for the shared-memory bitmap (>= 32 samples) ------
template <bool C64>
__device__ bool slow_centre(const ScanParams &P, long long i_f, long long &p)
{
    long long i = i_f - 1;
    while (i >= 0 && window_x<C64>(P, i) >= P.thr) --i;
    long long r;
    if (i < 0) {
        if (!P.left_is_start) {
            atomicOr(&P.hdr->flags, kFlagHalo);
            return false;
        }
        if (P.prev_bit) return false;           // run began before this window: no rise here (framer.py:98-100)
        r = 0;
    } else {
        r = i + 1;
    }
    p = (r + i_f) >> 1;                          // framer.py:113
    return true;
}

template <bool C64>
__device__ bool slow_match(const ScanParams &P, long long p)
{
    const int h = P.sps >> 1;
    const float xp = window_x<C64>(P, p);
    const float half = __fmul_rn(xp, 0.5f);
    const unsigned tmpl = 0x0285u;               // bits 0,2,7,9
    for (int k = 0; k < 16; k++) {
        bool hi = window_x<C64>(P, p + (long long)k * h) > half;
        if (hi != (((tmpl >> k) & 1u) != 0)) return false;
    }
    return true;
}

__device__ __forceinline__ bool demod_in_range(const ScanParams &P, long long p)
{
    // demod.py:67,82 in window coordinates: tag not late, last sample p + 119*sps + sps/2 inside in0
    const long long last = p + 119LL * P.sps + (P.sps >> 1);
    return (!P.left_is_start || p >= P.hist) && last <= P.n_loc - 1;
}

template <bool C64>
__device__ void slow_slice(const ScanParams &P, long long p, Rec &rec)
{
    const int sps = P.sps, h = sps >> 1;
    const float xp = window_x<C64>(P, p);
    long long lo = p - kNoise;
    if (lo < 0) lo = 0;
    const int n = (int)(p - lo);
    float med = median_serial([&](int i) { return window_x<C64>(P, lo + i); }, n);
    rec.idx = P.base + p;
    rec.snr = snr_db(xp, med);
    unsigned flags = (P.left_is_start && p < P.hist) ? kHitLate : 0u;
    unsigned w[4] = {0, 0, 0, 0};
    if (P.do_demod && demod_in_range(P, p)) {
        flags |= kHitDemodValid;
        unsigned syn = 0;
        for (int b = 0; b < kMsgBits; b++) {
            long long s = p + 8LL * sps + (long long)b * sps;
            if (window_x<C64>(P, s) > window_x<C64>(P, s + h)) {
                w[b >> 5] |= 1u << (b & 31);
                syn ^= P.crc_tab[b];
            }
        }
        const unsigned df = __brev(w[0]) >> 27;
        if (syn == 0 && df >= 17 && df <= 19) flags |= kHitCrcOk;
    }
    rec.flags = flags;
#pragma unroll
    for (int j = 0; j < 4; j++) rec.bits[j] = pack_msb_first(w[j]);
}

// ---- shared-memory layout ------------------------------------------------------------------------
struct TileSmem {
    float *mag;             // S
    unsigned *bits;         // S/32
    unsigned short *list;   // T/2 + 2
    Rec *stage;             // kStage
    int *scratch;           // 16 ints
};

__host__ __device__ inline size_t tile_smem_bytes(int T, int rh)
{
    size_t S = (size_t)kLeftHalo + T + rh;
    size_t b = S * 4;                    // mag
    b += ((S / 32) * 4 + 15) & ~(size_t)15;  // bits
    b += (((size_t)T / 2 + 2) * 2 + 15) & ~(size_t)15;   // list
    b += (size_t)kStage * sizeof(Rec);
    b += 16 * 4;
    return b;
}

template <int T>
__device__ __forceinline__ TileSmem carve(unsigned char *raw, int rh)
{
    TileSmem s;
    const size_t S = (size_t)kLeftHalo + T + rh;
    s.mag = reinterpret_cast<float *>(raw);
    raw += S * 4;
    s.bits = reinterpret_cast<unsigned *>(raw);
    raw += ((S / 32) * 4 + 15) & ~(size_t)15;
    s.list = reinterpret_cast<unsigned short *>(raw);
    raw += (((size_t)T / 2 + 2) * 2 + 15) & ~(size_t)15;
    s.stage = reinterpret_cast<Rec *>(raw);
    raw += (size_t)kStage * sizeof(Rec);
    s.scratch = reinterpret_cast<int *>(raw);
    return s;
}

// run length of ones ending just below bit `pos` (32 <= pos <= 63) of V
__device__ __forceinline__ int ones_below(unsigned long long V, int pos)
{
    return __clzll((long long)~(V << (64 - pos)));
}

// ---- the kernel ------------------------------------------------------------------------------------
template <int T, int HS, bool C64>
__global__ void __launch_bounds__(kThreads, (T <= 4096) ? 4 : 2) scan_slice_kernel(const ScanParams P)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const TileSmem sm = carve<T>(smem_raw, P.rh);
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int S = kLeftHalo + T + P.rh;
    const int h = HS ? HS : (P.sps >> 1);

    // ---- tile id (ticket: tiles start in index order whatever the CTA dispatch order) ----
    if (tid == 0) sm.scratch[12] = (int)atomicAdd(&P.hdr->tile_ticket, 1u);
    __syncthreads();
    const long long tile = sm.scratch[12];
    const long long a0 = tile * T;                 // array index of the tile's first owned sample
    const long long w0 = a0 + P.lead - kLeftHalo;  // window index of shared-memory slot 0
    const bool interior = (a0 - kLeftHalo >= 0) && (a0 + T + P.rh <= P.n_arr) &&
                          (w0 + kLeftHalo >= P.own_lo) && (w0 + kLeftHalo + T <= P.own_hi) && (w0 > 0);

    // ---- phase 1: HBM -> |IQ|^2 tile + threshold bitmap ----
    {
        constexpr int U = 6;
        const float thr = P.thr;
        for (int kb = wid * 32; kb < S; kb += kThreads * U) {
            float x[U];
            if (interior) {
                if (C64) {
                    const float2 *src = reinterpret_cast<const float2 *>(P.in) + (a0 - kLeftHalo) + lane;
                    float2 v[U];
#pragma unroll
                    for (int u = 0; u < U; u++)
                        if (kb + u * kThreads < S) v[u] = ldg_stream_f2(src + kb + u * kThreads);
#pragma unroll
                    for (int u = 0; u < U; u++)
                        if (kb + u * kThreads < S) x[u] = mag2(v[u]);
                } else {
                    const float *src = reinterpret_cast<const float *>(P.in) + (a0 - kLeftHalo) + lane;
#pragma unroll
                    for (int u = 0; u < U; u++)
                        if (kb + u * kThreads < S) x[u] = ldg_stream_f1(src + kb + u * kThreads);
                }
#pragma unroll
                for (int u = 0; u < U; u++) {
                    const int k = kb + u * kThreads;
                    if (k < S) {
                        sm.mag[k + lane] = x[u];
                        unsigned w = __ballot_sync(kFull, x[u] >= thr);
                        if (lane == 0) sm.bits[k >> 5] = w;
                    }
                }
            } else {
#pragma unroll
                for (int u = 0; u < U; u++) {
                    const int k = kb + u * kThreads;
                    if (k < S) {
                        const long long i = w0 + k + lane;
                        const float v = (i >= 0 && i < P.n_loc) ? window_x<C64>(P, i) : 0.0f;
                        bool bit = v >= thr;
                        if (i < 0) bit = P.left_is_start ? (i == -1 && P.prev_bit) : false;
                        sm.mag[k + lane] = v;
                        unsigned w = __ballot_sync(kFull, bit);
                        if (lane == 0) sm.bits[k >> 5] = w;
                    }
                }
            }
        }
    }
    __syncthreads();

    // ---- phase 2: falls -> pulse centres -> preamble match ----
    constexpr int NCH = (T / 16 + kThreads - 1) / kThreads;
    unsigned cmask[NCH], smask[NCH];
    int tmax = -1;              // largest centre (shared-memory slot) of a complete pulse
    long long tmax_slow = -1;   // same for slow-path pulses (window index)
    unsigned tflags = 0;
#pragma unroll
    for (int ci = 0; ci < NCH; ci++) {
        const int c = tid + ci * kThreads;
        cmask[ci] = 0;
        smask[ci] = 0;
        if (c >= T / 16) continue;
        const int w = kLeftHalo / 32 + (c >> 1), sh = (c & 1) * 16;
        const unsigned long long V = ((unsigned long long)sm.bits[w] << 32) | sm.bits[w - 1];
        const unsigned long long F = ~V & (V << 1);
        const unsigned long long R = V & ~(V << 1);
        unsigned fm = (unsigned)(F >> (32 + sh)) & 0xffffu;
        unsigned rm = (unsigned)(R >> (32 + sh)) & 0xffffu;
        const int k0 = kLeftHalo + c * 16;       // slot of chunk bit 0
        if (!interior) {
            const long long i0 = w0 + k0;
            long long lo = P.own_lo - i0, hi = P.own_hi - i0;
            lo = lo < 0 ? 0 : (lo > 16 ? 16 : lo);
            hi = hi < 0 ? 0 : (hi > 16 ? 16 : hi);
            const unsigned own = (hi > lo) ? (((1u << hi) - 1u) & ~((1u << lo) - 1u)) : 0u;
            fm &= own;
            rm &= own;
        }
        if (rm) tflags |= kFlagRise;
        if (fm) tflags |= kFlagFall;
        unsigned cm = 0, sl = 0;
        while (fm) {
            const int b = __ffs(fm) - 1;
            fm &= fm - 1;
            const int pos = 32 + sh + b;
            const int run = ones_below(V, pos);
            const int kf = k0 + b;
            if (run >= pos) {
                // run longer than the bitmap window: resolve from global memory
                long long p;
                if (slow_centre<C64>(P, w0 + kf, p)) {
                    tmax_slow = p > tmax_slow ? p : tmax_slow;
                    if (slow_match<C64>(P, p)) { cm |= 1u << b; sl |= 1u << b; }
                }
                continue;
            }
            if (!interior && (w0 + kf - run) < 0) {
                // run reaches window index -1: it began before this window (or before the halo)
                if (!P.left_is_start) atomicOr(&P.hdr->flags, kFlagHalo);
                continue;
            }
            const int kp = kf - ((run + 1) >> 1);    // framer.py:113: (rise + fall) // 2
            tmax = kp > tmax ? kp : tmax;
            if (preamble_match<HS>(sm.mag + kp, h)) cm |= 1u << b;
        }
        cmask[ci] = cm;
        smask[ci] = sl;
    }

    // ---- ordered compaction of this tile's candidates into sm.list ----
    int n_cand = 0;
#pragma unroll
    for (int ci = 0; ci < NCH; ci++) {
        int total;
        int off = n_cand + block_excl_scan(__popc(cmask[ci]), sm.scratch, total);
        unsigned cm = cmask[ci];
        const int k0 = kLeftHalo + (tid + ci * kThreads) * 16;
        while (cm) {
            const int b = __ffs(cm) - 1;
            cm &= cm - 1;
            sm.list[off++] = (unsigned short)((k0 + b) | (((smask[ci] >> b) & 1u) << 15));
        }
        n_cand += total;
    }

    // ---- per-tile reductions: edge flags, last pulse; candidate slot allocation ----
    {
        __shared__ long long s_max[kWarps];
        __shared__ unsigned s_flag[kWarps];
        const unsigned f = __reduce_or_sync(kFull, tflags);
        const int m = __reduce_max_sync(kFull, tmax);
        long long ms = tmax_slow;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            const long long o = __shfl_xor_sync(kFull, ms, d);
            ms = o > ms ? o : ms;
        }
        long long mw = (m >= 0) ? (w0 + m) : -1;
        mw = ms > mw ? ms : mw;
        if (lane == 0) {
            s_flag[wid] = f;
            s_max[wid] = mw;
        }
        __syncthreads();
        if (tid == 0) {
            unsigned ff = 0;
            long long mm = -1;
#pragma unroll
            for (int q = 0; q < kWarps; q++) {
                ff |= s_flag[q];
                mm = s_max[q] > mm ? s_max[q] : mm;
            }
            if (ff) atomicOr(&P.hdr->flags, ff);
            if (mm >= 0) atomicMax(&P.hdr->last_pulse, mm);
            unsigned off = 0;
            if (n_cand > 0) off = (unsigned)atomicAdd(&P.hdr->n_cand, (unsigned long long)n_cand);
            P.tile_tab[tile] = make_uint2(off, (unsigned)n_cand);
            sm.scratch[13] = (int)off;
            // x[n_out - 1] for the carried prev_in0 (framer.py:87)
            const long long kl = (P.own_hi - 1) - w0;
            if (kl >= kLeftHalo && kl < kLeftHalo + T) P.hdr->x_last = sm.mag[kl];
        }
        __syncthreads();
    }
    if (n_cand == 0) return;
    const unsigned g_off = (unsigned)sm.scratch[13];

    // ---- phase 3: per-candidate SNR + PPM slice + CRC, one warp per candidate ----
    unsigned crc_t[4];
#pragma unroll
    for (int j = 0; j < 4; j++) crc_t[j] = (j * 32 + lane < kMsgBits) ? P.crc_tab[j * 32 + lane] : 0u;

    for (int b0 = 0; b0 < n_cand; b0 += kStage) {
        const int nb = min(kStage, n_cand - b0);
        for (int c = wid; c < nb; c += kWarps) {
            const unsigned e = sm.list[b0 + c];
            const int kf = e & 0x7fff;
            Rec &rec = sm.stage[c];
            if (e >> 15) {
                if (lane == 0) {
                    long long p;
                    slow_centre<C64>(P, w0 + kf, p);
                    slow_slice<C64>(P, p, rec);
                }
                continue;
            }
            const int wq = kf >> 5, bq = kf & 31;
            const unsigned long long V = ((unsigned long long)sm.bits[wq] << 32) | sm.bits[wq - 1];
            const int run = ones_below(V, 32 + bq);
            const int kp = kf - ((run + 1) >> 1);
            const long long p = w0 + kp;
            // noise window x[max(0, p-100) : p] (framer.py:156-159)
            int n = kNoise;
            if (p < kNoise) n = (int)p;
            const float med = warp_median(sm.mag + kp - n, n, lane);
            const float xp = sm.mag[kp];
            unsigned flags = (P.left_is_start && p < P.hist) ? kHitLate : 0u;
            unsigned w[4] = {0u, 0u, 0u, 0u};
            if (P.do_demod && demod_in_range(P, p)) {
                flags |= kHitDemodValid;
                const float *sob = sm.mag + kp + 8 * P.sps;       // demod.py:75
                unsigned syn = 0;
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const int bi = j * 32 + lane;
                    bool bit = false;
                    if (bi < kMsgBits) bit = sob[bi * P.sps] > sob[bi * P.sps + h];   // demod.py:87-95
                    w[j] = __ballot_sync(kFull, bit);
                    syn ^= bit ? crc_t[j] : 0u;
                }
                syn = __reduce_xor_sync(kFull, syn);
                const unsigned df = __brev(w[0]) >> 27;
                if (syn == 0 && df >= 17 && df <= 19) flags |= kHitCrcOk;    // decoder.py:667-685
            }
            if (lane == 0) {
                rec.idx = P.base + p;
                rec.snr = snr_db(xp, med);
                rec.flags = flags;
            }
            if (lane < 4) rec.bits[lane] = pack_msb_first(lane == 0 ? w[0] : lane == 1 ? w[1] : lane == 2 ? w[2] : w[3]);
        }
        __syncthreads();
        // coalesced copy-out: 2 x 16 bytes per record
        if (tid < nb * 2) {
            const unsigned slot = g_off + b0 + (tid >> 1);
            if (slot < P.cand_cap)
                reinterpret_cast<uint4 *>(P.cand + slot)[tid & 1] = reinterpret_cast<const uint4 *>(sm.stage + (tid >> 1))[tid & 1];
            else if ((tid & 1) == 0)
                atomicOr(&P.hdr->flags, kFlagCandOverflow);
        }
        __syncthreads();
    }
}

}  // namespace adsb
