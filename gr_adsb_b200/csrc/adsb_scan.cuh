// K1: fused |IQ|^2 + threshold/run/centre + preamble match + (per match) SNR, PPM slice, CRC-24.
// One CTA per tile of T samples; the tile plus a 128-sample left halo and a >= 120-symbol right halo
// is staged once in shared memory as float32 |IQ|^2, so every sample is fetched from HBM once.
//
//   phase 1  HBM -> registers -> |IQ|^2 -> shared memory, threshold bitmap by warp ballot
//   phase 2  bitmap -> falling edges -> pulse centres (bit-parallel for runs <= LMAX samples),
//            dense 3-tap pre-filter, full 16-tap preamble test on the survivors
//   phase 3  two kinds of task spread over the warps: exact medians of the 100-sample noise windows of four candidates
//            (8 lanes and 16 keys per lane each: register sorting network + selection of the two middle ranks), and
//            2 x 112 PPM decisions of a pair by ballot with CRC-24 by XOR-reduction; edge tiles: one warp per candidate
// Interior tiles of the specialised sample rates (2, 4, 8 samples/symbol) take the fast paths; tiles
// touching the window edges, and other rates, take a generic per-edge path with identical results.
#pragma once
#include "adsb_device.cuh"

namespace adsb {

#ifndef ADSB_NCH2
#define ADSB_NCH2 1           // detect chunks per thread at 2 / 4 samples per symbol (tile = ADSB_NCH2 * ADSB_CW * 256 samples)
#endif
#ifndef ADSB_CW8
#define ADSB_CW8 12           // samples per detect chunk of the other kernels (8 Msps, unspecialised rates): two chunks per thread
#endif
#ifndef ADSB_MIN_CTAS_BIG
#define ADSB_MIN_CTAS_BIG (ADSB_CW8 == 16 ? 4 : ADSB_CW8 == 12 ? 5 : 3)  // resident CTAs per SM of those kernels
#endif
#ifndef ADSB_CW
#define ADSB_CW 20            // samples per detect chunk at 2 / 4 samples per symbol: with 20 the float4 loads of the dense pre-filter
#endif                        // (thread stride 80 B) fall into 8 different 16-byte bank groups per quarter-warp; with 16 (stride 64 B)
                              // they are 4-way bank conflicted -- 40 % of the kernel's shared-memory wavefronts (ncu, round 2)
#ifndef ADSB_LMAX2
#define ADSB_LMAX2 8          // longest run resolved bit-parallel at 2 samples/symbol (8 otherwise)
#endif

// ---- compile-time geometry per samples-per-symbol and CTA size -------------------------------------------
// NT = threads per CTA (256; 128- and 512-thread CTAs measured slower in round 2, profiles/r02/variants_call1.jsonl).  A thread
// owns one 20-sample chunk of the tile (two 12-sample chunks at 8 Msps and unspecialised rates), so the tile grows with the CTA.
template <int SPS, int NT = kThreads>
struct Geo {
    static constexpr int CW = (SPS == 2 || SPS == 4) ? ADSB_CW : ADSB_CW8;  // samples per detect chunk (one 32-bit window)
    static constexpr int LB = 32 - CW;                                // look-back bits of the window below the chunk
    static constexpr int T = ((SPS == 2 || SPS == 4) ? ADSB_NCH2 * CW : 2 * CW) * NT;
    static constexpr int H = SPS / 2;                                  // 0: run-time
    // right halo: >= 120*SPS and such that the staged span is a multiple of NT (384 / 640 / 1152 for NT = 256)
    static constexpr int RH = SPS == 0 ? 0 : ((120 * SPS + kLeftHalo + NT - 1) / NT) * NT - kLeftHalo;
    static constexpr int S = kLeftHalo + T + RH;
    static constexpr int NPT = S / NT;                                 // slots per thread in phase 1
    static constexpr int LMAX = ADSB_LMAX2 > 0 && SPS == 2 ? ADSB_LMAX2 : 8;   // run lengths resolved bit-parallel
    static constexpr int DB = LMAX / 2;                                // how far a centre lies before its chunk
    static constexpr int NV = (CW + 3 + 5 * H + 3) / 4;                // float4 loads of the dense pre-filter
    static_assert(SPS == 0 || (S % NT == 0 && RH >= 120 * SPS && RH % 32 == 0), "tile geometry");
    static_assert(CW % 4 == 0 && CW >= 12 && LB >= LMAX + 1 && T % 32 == 0, "chunk geometry");
    static_assert(DB <= 4, "the dense pre-filter's register window starts 4 samples before the chunk (LMAX <= 9)");
};

__host__ __device__ inline int tile_items(int sps, int nt = kThreads) { return ((sps == 2 || sps == 4) ? ADSB_NCH2 * ADSB_CW : 2 * ADSB_CW8) * nt; }
__host__ __device__ inline int right_halo(int sps, int nt = kThreads)
{
    if (sps == 2 || sps == 4 || sps == 8) return ((120 * sps + kLeftHalo + nt - 1) / nt) * nt - kLeftHalo;
    return ((120 * sps + 31) / 32) * 32;
}

#ifndef ADSB_MIN_CTAS
#define ADSB_MIN_CTAS (ADSB_CW == 16 ? 8 : 7)   // resident CTAs per SM of the 2 / 4 samples-per-symbol kernels: 8 x 24.3 KB
                                                // (4096-sample tiles) or 7 x 31 KB (5120-sample tiles) of shared memory
#endif
constexpr int kTileSlots = 32;    // candidate records reserved per tile (typical busy tile: 5-10)
constexpr int kSlowCap = 256;     // long-run falls deferred per tile before falling back to serial handling

// Phase 3 takes the noise medians four candidates per warp where tiles usually hold several candidates, two per warp
// otherwise.  Measured on the B200 (profiles/r02/variants_quad_median.jsonl, variants_median_policy.jsonl, bench lines):
// 2 Msps, 10k bursts/s (8.7 candidates per 5120-sample tile): 0.460 ms against 0.483 ms; 8 Msps, 10k bursts/s (2.6 per
// 6144-sample tile: a quad task mostly sorts pads): 0.474 ms against 0.444 ms.  One policy per kernel on purpose: a
// build with both median routines in the same kernel, chosen per task, ran 3-8 % slower than either.
#ifndef ADSB_QUAD_MEDIANS_SPS
#define ADSB_QUAD_MEDIANS_SPS 2
#endif
template <int SPS>
constexpr bool kQuadMedians = SPS == ADSB_QUAD_MEDIANS_SPS;

// ---- shared-memory layout ------------------------------------------------------------------------
struct TileSmem {
    float *mag;             // S
    unsigned *bits;         // S/32
    unsigned short *list;   // T/2 + 2
    unsigned *xmask;        // T/CW   extra candidate bits (by fall position) from the long-run pass
    unsigned *xlong;        // T/CW   which of them are long-run entries
    unsigned short *slow;   // kSlowCap fall slots
    int *scratch;           // scratch_ints(NW): slots below + per-warp reduction cells
};

__host__ __device__ inline size_t align16(size_t b) { return (b + 15) & ~(size_t)15; }

__host__ __device__ constexpr int scratch_ints(int nw) { return 32 + 4 * nw; }      // 64 for the 8-warp CTA

__host__ __device__ inline size_t tile_smem_bytes(int T, int rh, int cw, int nw = kWarps)
{
    size_t S = (size_t)kLeftHalo + T + rh;
    size_t b = S * 4;
    b += align16((S / 32 + 1) * 4);
    b += align16(((size_t)T / 2 + 2) * 2);
    b += align16((size_t)(T / cw) * 4) * 2;      // one word per detect chunk
    b += align16((size_t)kSlowCap * 2);
    b += (size_t)scratch_ints(nw) * 4;
    return b;
}

__device__ __forceinline__ TileSmem carve(unsigned char *raw, int T, int rh, int cw)
{
    TileSmem s;
    const size_t S = (size_t)kLeftHalo + T + rh;
    s.mag = reinterpret_cast<float *>(raw);
    raw += S * 4;
    s.bits = reinterpret_cast<unsigned *>(raw);
    raw += align16((S / 32 + 1) * 4);
    s.list = reinterpret_cast<unsigned short *>(raw);
    raw += align16(((size_t)T / 2 + 2) * 2);
    s.xmask = reinterpret_cast<unsigned *>(raw);
    raw += align16((size_t)(T / cw) * 4);
    s.xlong = reinterpret_cast<unsigned *>(raw);
    raw += align16((size_t)(T / cw) * 4);
    s.slow = reinterpret_cast<unsigned short *>(raw);
    raw += align16((size_t)kSlowCap * 2);
    s.scratch = reinterpret_cast<int *>(raw);
    return s;
}

// scratch slots
enum { kScrFast = 11, kScrTile = 12, kScrOff = 13, kScrSlowN = 14, kScrNCand = 15, kScrCnt = 16 /* [2][NW] */ };
// per-warp cells behind the counts, NW = warps per CTA: [NW] last pulse, [NW] edge flags, one long long
__host__ __device__ constexpr int scr_max(int nw) { return kScrCnt + 2 * nw; }
__host__ __device__ constexpr int scr_flag(int nw) { return kScrCnt + 3 * nw; }
__host__ __device__ constexpr int scr_max_long(int nw) { return kScrCnt + 4 * nw; }


// ---- thread-group barrier: scan_tile / slice_tile are written against this interface so that the phases can
// be moved to warp groups with named barriers; a persistent 8+4-warp producer/consumer variant was measured in
// round 1 and lost to the plain one-CTA-per-tile kernel (0.93 ms vs 0.61 ms: half the scanning warps per SM).
struct CtaBar {
    __device__ __forceinline__ void sync() const { __syncthreads(); }
    __device__ __forceinline__ bool sync_or(bool p) const { return __syncthreads_or(p) != 0; }
};
// ---- preamble test -----------------------------------------------------------------------------------
// framer.py:137-147 -- taps x[p + k*h], k = 0..15, against half the centre amplitude; template
// [1,0,1,0,0,0,0,1,0,1,0,0,0,0,0,0] (framer.py:50).
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

// same test for a centre that already passed the dense pre-filter (taps 3, 4, 5 known low)
template <int HS>
__device__ __forceinline__ bool preamble_match_rest(const float *m)
{
    const float xp = m[0];
    const float half = __fmul_rn(xp, 0.5f);
    if (!(m[2 * HS] > half)) return false;
    if (!(m[7 * HS] > half)) return false;
    if (!(m[9 * HS] > half)) return false;
    if (m[1 * HS] > half) return false;
    if (m[6 * HS] > half) return false;
    if (m[8 * HS] > half) return false;
#pragma unroll
    for (int k = 10; k < 16; k++)
        if (m[k * HS] > half) return false;
    return xp > half;
}

// a > b as a word whose SIGN BIT is the answer, for a, b that are +0, positive, +inf or the canonical NaN (see the
// dense pre-filter).  The host emulation compares directly: x86 gives inf - inf and propagated NaNs a sign bit.
__device__ __forceinline__ unsigned gt_sign(float a, float b)
{
#if defined(ADSB_HOST_EMU)
    return (a > b) ? 0x80000000u : 0u;
#else
    return __float_as_uint(__fadd_rn(b, -a));
#endif
}

__device__ __forceinline__ bool demod_in_range(const ScanParams &P, long long p)
{
    // demod.py:67,82 in window coordinates: tag not late, last sample p + 119*sps + sps/2 inside in0
    const long long last = p + 119LL * P.sps + (P.sps >> 1);
    return (!P.left_is_start || p >= P.hist) && last <= P.win_hi - 1;
}

// window sample i through shared memory when staged, global memory otherwise
template <int FMT>
struct Access {
    const ScanParams &P;
    const float *mag;
    long long w0;
    int S;
    __device__ __forceinline__ float operator()(long long i) const
    {
        const long long s = i - w0;
        if (s >= 0 && s < S) return mag[s];
        return window_x<FMT>(P, i);
    }
};

// Slot of the rising edge of the run that ends just below slot kf, walking the staged bitmap word by
// word; -1 when the run is still high at slot 0 (its start lies outside the staged span).
__device__ __forceinline__ int walk_back_smem(const unsigned *bits, int kf)
{
    const int top = kf - 1;
    int w = top >> 5;
    unsigned z = ~bits[w];
    if ((top & 31) != 31) z &= (2u << (top & 31)) - 1u;
    while (z == 0) {
        if (--w < 0) return -1;
        z = ~bits[w];
    }
    return w * 32 + (31 - __clz(z)) + 1;
}

// ---- long runs: warp-cooperative walk back to the rising edge ------------------------------------------
// Fall at shared-memory slot kf (window index w0 + kf).  Returns false when the pulse has no rising edge
// inside this window (framer.py:98-100).  All 32 lanes participate and get the same answer.
template <int FMT>
__device__ bool resolve_long(const ScanParams &P, const TileSmem &sm, long long w0, int kf, int lane, long long &p)
{
    // 1) bitmap in shared memory, 32 words per step
    int top = kf - 1;                       // highest slot that is part of the run
    long long r = -(1LL << 62);
    bool found = false;
    for (int wbase = top >> 5; wbase >= 0 && !found; wbase -= 32) {
        const int wi = wbase - lane;
        unsigned z = 0;
        if (wi >= 0) {
            z = ~sm.bits[wi];
            if (wi == (top >> 5) && (top & 31) != 31) z &= (2u << (top & 31)) - 1u;
        }
        const unsigned has = __ballot_sync(kFull, z != 0);
        if (has) {
            const int src = __ffs(has) - 1;
            const unsigned zz = __shfl_sync(kFull, z, src);
            const int wz = wbase - src;
            r = w0 + (long long)wz * 32 + (31 - __clz(zz)) + 1;
            found = true;
        }
    }
    // 2) beyond the staged span: global memory, 32 samples per step
    if (!found) {
        long long i0 = w0 - 1;              // next window index to examine
        while (true) {
            const long long i = i0 - lane;
            bool below;
            if (i >= 0) below = !(window_x<FMT>(P, i) >= P.thr);
            else if (i == -1) below = !(P.left_is_start && P.prev_bit);
            else below = true;
            const unsigned b = __ballot_sync(kFull, below);
            if (b) {
                r = i0 - (__ffs(b) - 1) + 1;
                break;
            }
            i0 -= 32;
        }
    }
    if (r < 0) {
        // the run reaches window index -1 with the bitmap still high there
        if (!P.left_is_start && lane == 0) atomicOr(&P.hdr->flags, kFlagHalo);
        return false;
    }
    if (r == 0 && !P.left_is_start) {
        // a halo'd shard cannot tell whether the run continues to the left of its halo
        if (lane == 0) atomicOr(&P.hdr->flags, kFlagHalo);
        return false;
    }
    p = (r + (w0 + kf)) >> 1;                // framer.py:113
    return true;
}

// 16-tap test by one warp through the generic accessor
template <int FMT>
__device__ __forceinline__ bool warp_match(const Access<FMT> &X, long long p, int h, int lane)
{
    const float xp = X(p);
    const float half = __fmul_rn(xp, 0.5f);
    bool hi = false;
    if (lane < 16) hi = X(p + (long long)lane * h) > half;
    return (__ballot_sync(kFull, hi) & 0xffffu) == 0x0285u;
}

// ---- per-candidate slice ------------------------------------------------------------------------------
// SNR ratio x[p] / median(noise window), 112 PPM decisions, CRC.  One warp; get(d) returns window sample p + d.
template <class Get>
__device__ __forceinline__ void slice_candidate(const ScanParams &P, Get get, long long p, int sps, int h, int lane,
                                                const unsigned (&crc_t)[4], bool always_valid, Rec &rec)
{
    int n = kNoise;                                   // framer.py:156-159
    if (!always_valid && p < kNoise) {
        n = (int)p;                                   // window start: the noise window is cut short
        if (!P.left_is_start && lane == 0) atomicOr(&P.hdr->flags, kFlagHalo);   // a shard cannot know what lies before its view
    }
    float v[4];
#pragma unroll
    for (int r = 0; r < 4; r++) {
        const int e = r * 32 + lane;
        v[r] = (e < n) ? get(e - n) : 0.0f;
    }
    const float med = warp_median_regs(v, n, lane);
    const float xp = get(0);
    unsigned flags = (!always_valid && P.left_is_start && p < P.hist) ? kHitLate : 0u;
    unsigned w[4] = {0u, 0u, 0u, 0u};
    if (P.do_demod && (always_valid || demod_in_range(P, p))) {
        flags |= kHitDemodValid;
        const int sob = 8 * sps;                          // demod.py:75
        unsigned syn = 0;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int bi = j * 32 + lane;
            bool bit = false;
            if (bi < kMsgBits) bit = get(sob + bi * sps) > get(sob + bi * sps + h);   // demod.py:87-95
            w[j] = __ballot_sync(kFull, bit);
            syn ^= bit ? crc_t[j] : 0u;
        }
        syn = __reduce_xor_sync(kFull, syn);
        const unsigned df = __brev(w[0]) >> 27;
        if (syn == 0 && df >= 17 && df <= 19) flags |= kHitCrcOk;    // decoder.py:667-685
    }
    if (lane == 0) {
        rec.idx = P.base + p;
        rec.snr = __fdiv_rn(xp, med);      // ratio; the compact kernel applies 10*log10(.)+1.6
        rec.flags = flags;
    }
    if (lane < 4) rec.bits[lane] = pack_msb_first(lane == 0 ? w[0] : lane == 1 ? w[1] : lane == 2 ? w[2] : w[3]);
}

// Four candidates of an interior tile per warp (8 lanes each): the noise medians / SNR ratios.  Full 100-sample noise
// window, burst fully staged, so no edge cases; mp = shared-memory address of the candidate's centre sample.
template <int FMT>
__device__ __forceinline__ void slice_quad_snr(const ScanParams &P, const float *mp, long long p, int lane,
                                               bool write, Rec &rec)
{
    const int l8 = lane & 7;
    unsigned k[16];
    bool nan = false;
    // 100 samples + 14 pads below everything + 14 pads above everything: the median's two order statistics
    // (ranks 49, 50 of 100) are ranks 63, 64 of the 128, the two middle ones
    constexpr int kLowPads = (128 - kNoise) / 2;
#pragma unroll
    for (int r = 0; r < 16; r++) {
        const int e = r * 8 + l8;
        const float v = (e < kNoise) ? mp[e - kNoise] : 0.0f;
        if (FMT != kFmtF32) {
            k[r] = (e < kNoise) ? __float_as_uint(v) : (e < kNoise + kLowPads ? 0u : 0x7f800000u);
        } else {
            nan |= (e < kNoise) && (v != v);
            k[r] = (e < kNoise) ? f2key(v) : (e < kNoise + kLowPads ? 0u : 0xffffffffu);
        }
    }
    unsigned lane_top;
    const unsigned m = quarter_mid128(k, lane, lane_top);        // lanes 0-3 of the group: rank 63, lanes 4-7: rank 64
    // |IQ|^2 computed in phase 1 is never negative, so NaN bits sort above +inf (the high pads): a NaN among the
    // samples shows as some lane's largest key; a vote costs no shuffle
    if (FMT != kFmtF32) nan = lane_top > 0x7f800000u;
    const unsigned anyn = __ballot_sync(kFull, nan);
    const unsigned mo = __shfl_xor_sync(kFull, m, 4);
    float med;
    if (FMT != kFmtF32)
        med = __fmul_rn(__fadd_rn(__uint_as_float(m), __uint_as_float(mo)), 0.5f);   // fl32(fl32(a+b)/2)
    else
        med = __fmul_rn(__fadd_rn(key2f(m), key2f(mo)), 0.5f);
    if ((anyn >> (lane & 24)) & 0xffu) med = __int_as_float(0x7fc00000);
    if (write && l8 == 0) {
        rec.idx = P.base + p;
        rec.snr = __fdiv_rn(mp[0], med);       // ratio; the compact kernel applies 10*log10(.)+1.6
    }
}

// The same for two candidates per warp (lanes 0-15 / 16-31, 8 keys per lane): 330 instead of 550 instructions per task,
// the cheaper routine where a tile rarely holds more than two or three candidates (kernels of the other sample rates).
template <int SPS, int FMT>
__device__ __forceinline__ void slice_pair_snr(const ScanParams &P, const float *mp, long long p, int lane,
                                               bool write, Rec &rec)
{
    const int hl = lane & 15;
    unsigned k[8];
    bool nan = false;
    // 100 samples + 14 pads below everything + 14 pads above everything: the median's two order statistics
    // (ranks 49, 50 of 100) are ranks 63, 64 of the 128, the two middle ones
    constexpr int kLowPads = (128 - kNoise) / 2;
#pragma unroll
    for (int r = 0; r < 8; r++) {
        const int e = r * 16 + hl;
        const float v = (e < kNoise) ? mp[e - kNoise] : 0.0f;
        if (FMT != kFmtF32) {
            // |IQ|^2 computed here is never negative: the raw bits order like the values; NaN bits sort above +inf
            k[r] = (e < kNoise) ? __float_as_uint(v) : (e < kNoise + kLowPads ? 0u : 0x7f800000u);
        } else {
            nan |= (e < kNoise) && (v != v);
            k[r] = (e < kNoise) ? f2key(v) : (e < kNoise + kLowPads ? 0u : 0xffffffffu);
        }
    }
    unsigned ka, kb, top;
    half_mid128(k, lane, ka, kb, top);
    float med;
    if (FMT != kFmtF32) {
        med = __fmul_rn(__fadd_rn(__uint_as_float(ka), __uint_as_float(kb)), 0.5f);   // fl32(fl32(a+b)/2)
        if (top > 0x7f800000u) med = __int_as_float(0x7fc00000);
    } else {
        const unsigned anyn = __ballot_sync(kFull, nan);
        med = __fmul_rn(__fadd_rn(key2f(ka), key2f(kb)), 0.5f);
        if ((anyn >> (lane & 16)) & 0xffffu) med = __int_as_float(0x7fc00000);
    }
    if (write && hl == 0) {
        rec.idx = P.base + p;
        rec.snr = __fdiv_rn(mp[0], med);       // ratio; the compact kernel applies 10*log10(.)+1.6
    }
}

template <int SPS>
__device__ __forceinline__ void slice_pair_bits(const ScanParams &P, const float *mp, int lane,
                                                bool write, Rec &rec)
{
    constexpr int H = SPS / 2;
    const int hl = lane & 15;
    unsigned flags = 0;
    unsigned w[4] = {0u, 0u, 0u, 0u};
    if (P.do_demod) {
        flags = kHitDemodValid;
        const float *sob = mp + 8 * SPS;                       // demod.py:75
        unsigned syn = 0;
        unsigned bal[8];
        bal[7] = 0u;
#pragma unroll
        for (int j = 0; j < 7; j++) {
            const int bi = j * 16 + hl;
            const bool bit = sob[bi * SPS] > sob[bi * SPS + H];   // demod.py:87-95
            bal[j] = __ballot_sync(kFull, bit);
            syn ^= bit ? P.crc_tab[bi] : 0u;
        }
        // this half-warp's 16 bits of two consecutive ballots -> one word, by one byte permute
        const unsigned sel = (lane & 16) ? 0x7632u : 0x5410u;
#pragma unroll
        for (int q = 0; q < 4; q++) w[q] = __byte_perm(bal[2 * q], bal[2 * q + 1], sel);
#pragma unroll
        for (int d = 8; d > 0; d >>= 1) syn ^= __shfl_xor_sync(kFull, syn, d);
        const unsigned df = __brev(w[0]) >> 27;
        if (syn == 0 && df >= 17 && df <= 19) flags |= kHitCrcOk;    // decoder.py:667-685
    }
    if (write) {
        if (hl == 0) rec.flags = flags;
        if (hl < 4) rec.bits[hl] = pack_msb_first(hl == 0 ? w[0] : hl == 1 ? w[1] : hl == 2 ? w[2] : w[3]);
    }
}

// The 32-bit window of detect chunk c: its CW samples in bits LB..31, the LB samples before it in bits 0..LB-1
// (LB >= longest bit-parallel run + 1).  k0 = slot of the chunk's first sample.
template <int CW>
__device__ __forceinline__ unsigned chunk_window(const unsigned *bits, int k0)
{
    constexpr int LB = 32 - CW;
    const int s0 = k0 - LB;                       // slot of window bit 0
    if (CW == 16) {                               // s0 is a multiple of 16: one or two aligned words
        const unsigned cur = bits[k0 >> 5], prev = bits[(k0 >> 5) - 1];
        return (k0 & 16) ? cur : __funnelshift_l(prev, cur, 16);
    }
    return __funnelshift_r(bits[s0 >> 5], bits[(s0 >> 5) + 1], s0 & 31);
}

// ---- phases 1 + 2 of one tile, by the 256 scan threads (threads 0..255 of the CTA) --------------------------
// On return (after the caller's next group barrier) sm.list, scratch[kScrNCand|kScrOff|kScrFast] are valid.
// UG: loads an edge tile keeps in flight per thread (the latency-path kernel, which runs alone on its SM and has the
// registers, takes most of a tile in one batch: one or two PCIe round trips instead of four)
template <int SPS, int FMT, int NT, class GB, int UG = 6>
__device__ __forceinline__ void scan_tile(const ScanParams &P, const TileSmem &sm, const long long tile, const GB gb)
{
    using G = Geo<SPS, NT>;
    constexpr int NW = NT / 32;                    // warps per CTA
    constexpr int T = G::T;
    const int rh = SPS ? G::RH : P.rh;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;      // scan threads are threads 0..255
    const int S = kLeftHalo + T + rh;
    const int h = SPS ? G::H : (P.sps >> 1);
    const float thr = P.thr;

    if (tid == 0) {
        sm.scratch[kScrSlowN] = 0;
    }
    const long long a0 = tile * T;                 // array index of the tile's first owned sample
    const long long w0 = a0 + P.lead - kLeftHalo;  // window index of shared-memory slot 0
    // interior tile (host-computed range): staged span inside the array, chunk fully owned, w0 > 0
    const bool fast = SPS != 0 && ((unsigned)tile - P.fast_lo) <= P.fast_span && P.fast_span != 0xffffffffu;

    // ---- phase 1: HBM -> |IQ|^2 tile + threshold bitmap ----
    if (fast && P.prefetch_tiles) {
        // pull the tile one wave of CTAs ahead into L2 (one 128-byte line per thread): by the time a CTA
        // gets to it, its loads are L2 hits instead of HBM round trips
        const long long ta = tile + P.prefetch_tiles;
        if (((unsigned)ta - P.fast_lo) <= P.fast_span) {
            // only that tile's own T samples: its halos are the neighbouring tiles' own ranges, prefetched by the
            // CTAs that run next to this one; T * itemsize / 128 lines = 320 for complex64 at T = 5120
            const char *base = reinterpret_cast<const char *>(P.in) + ta * T * item_bytes<FMT>();
            constexpr int kLines = T * item_bytes<FMT>() / 128;
#if !defined(ADSB_HOST_EMU)
#pragma unroll
            for (int l = tid; l < kLines; l += NT)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(base + (size_t)l * 128));
#else
            (void)base;
            (void)kLines;
#endif
        }
    }
    if (fast) {
        constexpr int U = 6;
        // slot u of warp wid covers the 32 samples from (wid * NPT + u) * 32: every load is one coalesced row and a
        // warp's bitmap words are consecutive, so lane 0 stores them in pairs (when NPT is even: 8-byte alignment)
        const int p1_base = wid * (G::NPT * 32) + lane;
        const int p1_word = wid * G::NPT;
        constexpr bool p1_pairs = (G::NPT % 2) == 0;
        unsigned w_even = 0;
        auto put = [&](int u, float x) {
            sm.mag[p1_base + u * 32] = x;
            const unsigned w = __ballot_sync(kFull, x >= thr);
            if (p1_pairs) {
                if (u & 1) {
                    if (lane == 0) *reinterpret_cast<uint2 *>(sm.bits + p1_word + (u - 1)) = make_uint2(w_even, w);
                } else {
                    w_even = w;
                }
            } else if (lane == 0) {
                sm.bits[p1_word + u] = w;
            }
        };
        if (FMT == kFmtC64) {
            const float2 *src = reinterpret_cast<const float2 *>(P.in) + (a0 - kLeftHalo) + p1_base;
#pragma unroll
            for (int u0 = 0; u0 < G::NPT; u0 += U) {
                unsigned long long v[U];
#pragma unroll
                for (int u = 0; u < U; u++)
                    if (u0 + u < G::NPT) v[u] = ldg_stream_b64(src + (u0 + u) * 32);
#pragma unroll
                for (int u = 0; u < U; u++)
                    if (u0 + u < G::NPT) put(u0 + u, mag2_packed(v[u]));
            }
        } else if (FMT == kFmtSC16) {
            const unsigned *src = reinterpret_cast<const unsigned *>(P.in) + (a0 - kLeftHalo) + p1_base;
#pragma unroll
            for (int u0 = 0; u0 < G::NPT; u0 += U) {
                unsigned v[U];
#pragma unroll
                for (int u = 0; u < U; u++)
                    if (u0 + u < G::NPT) v[u] = ldg_stream_b32(src + (u0 + u) * 32);
#pragma unroll
                for (int u = 0; u < U; u++)
                    if (u0 + u < G::NPT) put(u0 + u, mag2_sc16(v[u]));
            }
        } else if (FMT == kFmtCU8) {
            const unsigned short *src = reinterpret_cast<const unsigned short *>(P.in) + (a0 - kLeftHalo) + p1_base;
#pragma unroll
            for (int u0 = 0; u0 < G::NPT; u0 += U) {
                unsigned v[U];
#pragma unroll
                for (int u = 0; u < U; u++)
                    if (u0 + u < G::NPT) v[u] = ldg_stream_b16(src + (u0 + u) * 32);
#pragma unroll
                for (int u = 0; u < U; u++)
                    if (u0 + u < G::NPT) put(u0 + u, mag2_cu8(v[u]));
            }
        } else {
            const float *src = reinterpret_cast<const float *>(P.in) + (a0 - kLeftHalo) + p1_base;
#pragma unroll
            for (int u0 = 0; u0 < G::NPT; u0 += U) {
                float v[U];
#pragma unroll
                for (int u = 0; u < U; u++)
                    if (u0 + u < G::NPT) v[u] = ldg_stream_f1(src + (u0 + u) * 32);
#pragma unroll
                for (int u = 0; u < U; u++)
                    if (u0 + u < G::NPT) put(u0 + u, v[u]);
            }
        }
    } else {
        if (P.peer_ready) {
            // multi-GPU peer mode: this tile may read a neighbour's samples -- not before that rank has said they are
            // resident for this step (its shard_sync_kernel writes the step counter into our control block)
            if (tid == 0) {
                const unsigned long long k = P.peer_ready[2];
                bool ok = true;
                if (P.in_left) ok = wait_ge(P.peer_ready + 0, k) && ok;
                if (P.in_right) ok = wait_ge(P.peer_ready + 1, k) && ok;
                fence_system();
                if (!ok) atomicOr(&P.hdr->flags, kFlagPeerTimeout);
            }
            gb.sync();
        }
        // loads in batches, issued back to back before the first one is used: on the latency path these are reads of
        // pinned host memory over PCIe (microseconds each), and a call of GNU Radio size consists of edge tiles only
        for (int kb = wid * 32; kb < S; kb += NT * UG) {
            unsigned long long raw[UG];
            bool have[UG];
#pragma unroll
            for (int u = 0; u < UG; u++) {
                const long long i = w0 + kb + u * NT + lane;
                raw[u] = 0;
                have[u] = kb + u * NT < S && i >= 0 && i < P.win_hi && window_raw<FMT>(P, i, raw[u]);
            }
#pragma unroll
            for (int u = 0; u < UG; u++) {
                const int k = kb + u * NT;
                if (k < S) {                                     // same for the whole warp
                    const long long i = w0 + k + lane;
                    const float v = have[u] ? raw_to_x<FMT>(raw[u]) : 0.0f;
                    bool bit = v >= thr;
                    if (i < 0) bit = P.left_is_start ? (i == -1 && P.prev_bit) : false;
                    sm.mag[k + lane] = v;
                    const unsigned w = __ballot_sync(kFull, bit);
                    if (lane == 0) sm.bits[k >> 5] = w;
                }
            }
        }
    }
    gb.sync();

    // ---- phase 2: falls -> pulse centres -> preamble match ----
    // Each thread owns the falls of CW-sample chunks; W holds the chunk in bits LB..31 and the LB samples
    // before it in bits 0..LB-1.  cmask: candidates keyed by the bit of their FALL inside the chunk.
    constexpr int CW = G::CW, LB = G::LB;
    constexpr unsigned kChunkBits = 0xffffffffu << LB;
    constexpr int NCH = (T / CW + NT - 1) / NT;
    unsigned cmask[NCH], lmask[NCH];
    int cent[NCH];              // centre slot of the candidate found last in this thread's chunk (valid when it has exactly one)
    int tmax = -1;              // largest centre (shared-memory slot) of a complete short pulse
    unsigned tflags = 0;
    bool have_long = false;
#pragma unroll
    for (int ci = 0; ci < NCH; ci++) {
        const int c = tid + ci * NT;
        cmask[ci] = 0;
        lmask[ci] = 0;
        cent[ci] = 0;
        if (c >= T / CW) continue;
        const int k0 = kLeftHalo + c * CW;       // slot of chunk bit 0 (= W bit LB)
        const unsigned W = chunk_window<CW>(sm.bits, k0);
        unsigned F = ~W & (W << 1) & kChunkBits;
        unsigned R = W & ~(W << 1) & kChunkBits;
        unsigned cm = 0, longf = 0, lm = 0;
        if (fast) {
            if (R) tflags |= kFlagRise;
            if (F) tflags |= kFlagFall;
            // bit-parallel centres: run of length L ending at fall f has centre f - ceil(L/2) (framer.py:113)
            unsigned acc = F, C = 0;
#pragma unroll
            for (int L = 1; L <= G::LMAX; L++) {
                const unsigned nxt = W << (L + 1);
                C |= (acc & ~nxt) >> ((L + 1) >> 1);
                acc &= nxt;
            }
            // runs longer than LMAX: walk the staged bitmap in-thread; only runs that leave the staged
            // span are deferred to the warp-cooperative path
            unsigned lf = acc >> LB;
            while (lf) {
                const int b = __ffs(lf) - 1;
                lf &= lf - 1;
                const int kf = k0 + b;
                const int pos = LB + b;
                const int run = __clz(~(W << (32 - pos)));      // ones below the fall inside W
                int r = kf - run;
                if (run >= pos) {                               // run leaves W: walk the staged bitmap
                    r = walk_back_smem(sm.bits, kf);
                    if (r < 0) {
                        longf |= 1u << b;
                        continue;
                    }
                }
                const int kp = (r + kf) >> 1;                   // framer.py:113
                tmax = kp > tmax ? kp : tmax;
                if (preamble_match<G::H>(sm.mag + kp, h)) {
                    cm |= 1u << b;
                    if (kp < kNoise) lm |= 1u << b;             // noise window leaves the staged span
                    cent[ci] = kp;
                }
            }
            if (C) {
                tmax = max(tmax, k0 + (31 - __clz(C)) - LB);
                // dense pre-filter: taps 3,4,5 must not exceed half the centre sample (framer.py:141,50)
                float vals[G::NV * 4];
#pragma unroll
                for (int q = 0; q < G::NV; q++) {
                    const float4 t = *reinterpret_cast<const float4 *>(sm.mag + k0 - 4 + 4 * q);
                    vals[4 * q] = t.x; vals[4 * q + 1] = t.y; vals[4 * q + 2] = t.z; vals[4 * q + 3] = t.w;
                }
                unsigned Q = 0;
                if (FMT != kFmtF32) {
                    // |IQ|^2 computed in phase 1 is +0, positive, +inf or the canonical NaN 0x7fffffff, for which
                    // `m3 > half` is exactly the sign bit of fl32(half - m3): distinct finite values never subtract
                    // to zero (no flush-to-zero here), inf - inf and anything with a NaN give the sign-less canonical
                    // NaN.  One FADD (FMA pipe) and one funnel shift replace FSETP + SEL + ADD, all three on the ALU pipe
                    // this kernel loads most.  (float32 input may hold -0.0 or negative values: compare as before.)
#pragma unroll
                    for (int j = CW - 2 + G::DB; j >= 0; j--) {
                        const int i0 = 4 - G::DB + j;
                        const float half = __fmul_rn(vals[i0], 0.5f);
                        const float m3 = fmaxf(fmaxf(vals[i0 + 3 * G::H], vals[i0 + 4 * G::H]), vals[i0 + 5 * G::H]);
                        Q = __funnelshift_l(gt_sign(m3, half), Q, 1);        // Q = Q << 1 | (m3 > half)
                    }
                    Q = ~Q << (LB - G::DB);               // bit LB - DB + j <-> position j (bit 31 is junk; C never has it)
                } else
                {
#pragma unroll
                    for (int j = 0; j < CW - 1 + G::DB; j++) {
                        const int i0 = 4 - G::DB + j;                 // index of position k0 - DB + j in vals
                        const float half = __fmul_rn(vals[i0], 0.5f);
                        // none of the three taps may exceed half (NaN taps count as low, like `tap > half`)
                        const float m3 = fmaxf(fmaxf(vals[i0 + 3 * G::H], vals[i0 + 4 * G::H]), vals[i0 + 5 * G::H]);
                        Q |= !(m3 > half) ? (1u << (LB - G::DB + j)) : 0u;
                    }
                }
                unsigned Tm = C & Q;
                while (Tm) {
                    const int pc = 31 - __clz(Tm);
                    Tm &= ~(1u << pc);
                    if (preamble_match_rest<G::H>(sm.mag + k0 + pc - LB)) {
                        // key the candidate by its fall: first zero above the centre
                        const int fb = pc + __ffs(~(W >> (pc + 1)));
                        cm |= 1u << (fb - LB);
                        cent[ci] = k0 + pc - LB;
                    }
                }
            }
        } else {
            // generic per-edge path (window edges, halo'd shards, unspecialised sample rates)
            const long long i0 = w0 + k0;
            long long lo = P.own_lo - i0, hi = P.own_hi - i0;
            lo = lo < 0 ? 0 : (lo > CW ? CW : lo);
            hi = hi < 0 ? 0 : (hi > CW ? CW : hi);
            const unsigned own = (hi > lo) ? ((((1u << hi) - 1u) & ~((1u << lo) - 1u)) << LB) : 0u;
            F &= own;
            R &= own;
            if (R) tflags |= kFlagRise;
            if (F) tflags |= kFlagFall;
            unsigned fm = F >> LB;
            while (fm) {
                const int b = __ffs(fm) - 1;
                fm &= fm - 1;
                const int pos = LB + b;                         // bit of the fall in W
                const int run = __clz(~(W << (32 - pos)));      // ones just below it (pos if all of them)
                if (run >= pos) {
                    longf |= 1u << b;
                    continue;
                }
                const int kf = k0 + b;
                if (w0 + kf - run < 0) {
                    // the run reaches window index -1: it began before this window (framer.py:98-100)
                    if (!P.left_is_start) atomicOr(&P.hdr->flags, kFlagHalo);
                    continue;
                }
                const int kp = kf - ((run + 1) >> 1);           // framer.py:113
                tmax = kp > tmax ? kp : tmax;
                if (preamble_match<0>(sm.mag + kp, h)) {
                    cm |= 1u << b;
                    cent[ci] = kp;
                }
            }
        }
        if (longf) {
            have_long = true;
            unsigned lf = longf;
            while (lf) {
                const int b = __ffs(lf) - 1;
                lf &= lf - 1;
                const int slot = atomicAdd(&sm.scratch[kScrSlowN], 1);
                if (slot < kSlowCap) sm.slow[slot] = (unsigned short)(k0 + b);
            }
        }
        cmask[ci] = cm;
        lmask[ci] = lm;
    }

    // ---- per-warp candidate counts, edge flags, last pulse -> shared memory; the barrier that publishes them
    // also tells whether any run was deferred (rare), in which case the counts are redone after that pass ----
    int (*s_cnt)[NW] = reinterpret_cast<int (*)[NW]>(sm.scratch + kScrCnt);
    int *s_max = sm.scratch + scr_max(NW);
    unsigned *s_flag = reinterpret_cast<unsigned *>(sm.scratch + scr_flag(NW));
    int wtot[NCH];
#pragma unroll
    for (int ci = 0; ci < NCH; ci++) wtot[ci] = __reduce_add_sync(kFull, __popc(cmask[ci]));
    {
        const unsigned f = __reduce_or_sync(kFull, tflags);
        const int m = __reduce_max_sync(kFull, tmax);
        if (lane == 0) {
#pragma unroll
            for (int ci = 0; ci < NCH; ci++) s_cnt[ci][wid] = wtot[ci];
            s_flag[wid] = f;
            s_max[wid] = m;
        }
    }

    // ---- long runs (rare): one warp per deferred fall ----
    long long &s_max_long = *reinterpret_cast<long long *>(sm.scratch + scr_max_long(NW));
    if (gb.sync_or(have_long)) {
        for (int c = tid; c < T / CW; c += NT) {
            sm.xmask[c] = 0;
            sm.xlong[c] = 0;
        }
        if (tid == 0) s_max_long = -1;
        gb.sync();
        const int n_slow = sm.scratch[kScrSlowN];
        const Access<FMT> X{P, sm.mag, w0, S};
        if (n_slow <= kSlowCap) {
            for (int e = wid; e < n_slow; e += NW) {
                const int kf = sm.slow[e];
                long long p;
                if (!resolve_long<FMT>(P, sm, w0, kf, lane, p)) continue;
                const bool m = warp_match<FMT>(X, p, h, lane);
                if (lane == 0) {
                    atomicMax(&s_max_long, p);
                    if (m) {
                        const int c = (kf - kLeftHalo) / CW, b = (kf - kLeftHalo) % CW;
                        atomicOr(sm.xmask + c, 1u << b);
                        atomicOr(sm.xlong + c, 1u << b);
                    }
                }
            }
        } else {
            // more long runs than the deferral list holds: every warp rescans its own chunks
            for (int c = wid; c < T / CW; c += NW) {
                const int k0 = kLeftHalo + c * CW;
                const unsigned W = chunk_window<CW>(sm.bits, k0);
                unsigned F = ~W & (W << 1) & kChunkBits;
                const long long i0 = w0 + k0;
                long long lo = P.own_lo - i0, hi = P.own_hi - i0;
                lo = lo < 0 ? 0 : (lo > CW ? CW : lo);
                hi = hi < 0 ? 0 : (hi > CW ? CW : hi);
                F &= (hi > lo) ? ((((1u << hi) - 1u) & ~((1u << lo) - 1u)) << LB) : 0u;
                unsigned fm = F >> LB;
                while (fm) {
                    const int b = __ffs(fm) - 1;
                    fm &= fm - 1;
                    const int pos = LB + b;
                    const int run = __clz(~(W << (32 - pos)));
                    if (fast ? (run <= G::LMAX) : (run < pos)) continue;     // handled as a short run already
                    long long p;
                    if (!resolve_long<FMT>(P, sm, w0, k0 + b, lane, p)) continue;
                    const bool m = warp_match<FMT>(X, p, h, lane);
                    if (lane == 0) {
                        atomicMax(&s_max_long, p);
                        if (m) {
                            atomicOr(sm.xmask + c, 1u << b);
                            atomicOr(sm.xlong + c, 1u << b);
                        }
                    }
                }
            }
        }
        gb.sync();
#pragma unroll
        for (int ci = 0; ci < NCH; ci++) {
            const int c = tid + ci * NT;
            if (c < T / CW) {
                cmask[ci] |= sm.xmask[c];
                lmask[ci] |= sm.xlong[c];
            }
        }
#pragma unroll
        for (int ci = 0; ci < NCH; ci++) {
            wtot[ci] = __reduce_add_sync(kFull, __popc(cmask[ci]));
            if (lane == 0) s_cnt[ci][wid] = wtot[ci];
        }
        gb.sync();
    } else if (tid == 0) {
        s_max_long = -1;
    }

    // ---- ordered compaction of this tile's candidates into sm.list ----
    // candidates are rare: only warps that hold some compute their slots
    bool any_here = false;
#pragma unroll
    for (int ci = 0; ci < NCH; ci++) any_here |= wtot[ci] != 0;
    if (any_here) {
        int base = 0;
#pragma unroll
        for (int ci = 0; ci < NCH; ci++) {
            // lane q < 8 holds warp q's count: two warp reductions give this warp's offset and the row total
            const int v = (lane < NW) ? s_cnt[ci][lane] : 0;
            const int before = wtot[ci] ? __reduce_add_sync(kFull, lane < wid ? v : 0) : 0;
            const int total = (ci + 1 < NCH) ? __reduce_add_sync(kFull, v) : 0;
            if (wtot[ci]) {
                const int cnt = __popc(cmask[ci]);
                // rank inside the warp: a ballot when no lane holds more than one candidate (the usual case)
                const unsigned has = __ballot_sync(kFull, cnt != 0);
                int excl = __popc(has & ((1u << lane) - 1u));
                if (wtot[ci] != __popc(has)) {
                    int inc = cnt;
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        const int t = __shfl_up_sync(kFull, inc, d);
                        if (lane >= d) inc += t;
                    }
                    excl = inc - cnt;
                }
                int off = base + before + excl;
                unsigned cm = cmask[ci];
                const int k0 = kLeftHalo + (tid + ci * NT) * CW;
                if (cnt == 1 && lmask[ci] == 0) {
                    sm.list[off] = (unsigned short)cent[ci];        // centre kept since detection
                    cm = 0;
                }
                while (cm) {
                    const int b = __ffs(cm) - 1;
                    cm &= cm - 1;
                    const int kf = k0 + b;
                    unsigned short ent = (unsigned short)(kf | 0x8000);
                    if (!((lmask[ci] >> b) & 1u)) {
                        const int r = walk_back_smem(sm.bits, kf);       // >= 0: resolved in-thread above
                        ent = (unsigned short)((r + kf) >> 1);
                    }
                    sm.list[off++] = ent;
                }
            }
            base += total;
        }
    }
    int n_cand = 0;
    if (tid < 32) {
        int v = 0;
#pragma unroll
        for (int ci = 0; ci < NCH; ci++) v += (lane < NW) ? s_cnt[ci][lane] : 0;
        n_cand = __reduce_add_sync(kFull, v);
    }
    if (tid == 0) {
        unsigned ff = 0;
        int mi = -1;
#pragma unroll
        for (int q = 0; q < NW; q++) {
            ff |= s_flag[q];
            mi = s_max[q] > mi ? s_max[q] : mi;
        }
        long long mm = (mi >= 0) ? (w0 + mi) : -1;
        mm = s_max_long > mm ? s_max_long : mm;
        if (ff) atomicOr(&P.hdr->flags, ff);
        if (mm >= 0) atomicMax(&P.hdr->last_pulse, mm);
        // candidate records: a fixed block of kTileSlots per tile (no atomic round trip on the tile's critical
        // path); the rare tile with more takes space from the overflow area behind the blocks
        unsigned off = (unsigned)tile * kTileSlots;
        if (n_cand > kTileSlots)
            off = (unsigned)P.n_tiles_total * kTileSlots + (unsigned)atomicAdd(&P.hdr->n_cand, (unsigned long long)n_cand);
        P.tile_tab[tile] = make_uint2(off, (unsigned)n_cand);
        if (n_cand > 0) atomicAdd(&P.tile_chunk_tot[tile / kThreads], (unsigned)n_cand);   // gather_kernel's chunks
        sm.scratch[kScrOff] = (int)off;
        sm.scratch[kScrNCand] = n_cand;
        sm.scratch[kScrFast] = fast ? 1 : 0;
        // x[n_out - 1] for the carried prev_in0 (framer.py:87)
        const long long kl = (P.own_hi - 1) - w0;
        if (kl >= kLeftHalo && kl < kLeftHalo + T) P.hdr->x_last = sm.mag[kl];
    }
}

// ---- phase 3 of one tile, by a group of n_warps warps (tid = thread index inside the group) -------------------
template <int SPS, int FMT, int NT, class GB>
__device__ __forceinline__ void slice_tile(const ScanParams &P, const TileSmem &sm, const long long tile, const int tid,
                                           const int n_warps, const GB gb)
{
    using G = Geo<SPS, NT>;
    constexpr int T = G::T;
    const int rh = SPS ? G::RH : P.rh;
    const int lane = tid & 31, wid = tid >> 5;
    const int S = kLeftHalo + T + rh;
    const int h = SPS ? G::H : (P.sps >> 1);
    const long long w0 = tile * T + P.lead - kLeftHalo;
    const int n_cand = sm.scratch[kScrNCand];
    if (n_cand == 0) return;
    const bool fast = sm.scratch[kScrFast] != 0;
    const unsigned g_off = (unsigned)sm.scratch[kScrOff];

    // ---- phase 3: per-candidate SNR ratio + PPM slice + CRC ----
    const Access<FMT> X{P, sm.mag, w0, S};
    const int sps_rt = SPS ? SPS : P.sps;
    // records go straight to this tile's block of the candidate array (the two tasks of a pair fill
    // disjoint fields), so no staging buffer and no barrier after the slicing
    // Two task layouts, one per kernel (see kQuadMedians): noise medians four candidates per warp + PPM slices of pairs,
    // or both in pairs.
    if constexpr (kQuadMedians<SPS>) {
        // generic path of one pair (edge tiles, long runs, short noise windows): one warp per candidate, both roles
        auto generic_pair = [&](int c, bool two, unsigned e0, unsigned e1) {
            unsigned crc_t[4];
#pragma unroll
            for (int j = 0; j < 4; j++) crc_t[j] = (j * 32 + lane < kMsgBits) ? P.crc_tab[j * 32 + lane] : 0u;
            for (int q = 0; q < (two ? 2 : 1); q++) {
                const unsigned e = q ? e1 : e0;
                const int kf = e & 0x7fff;
                Rec &rec = P.cand[g_off + c + q];
                if (e >> 15) {
                    long long p;
                    resolve_long<FMT>(P, sm, w0, kf, lane, p);
                    slice_candidate(P, [&X, p](int d) { return X(p + d); }, p, sps_rt, h, lane, crc_t, false, rec);
                } else {
                    const float *mp = sm.mag + kf;              // list holds the centre slot
                    slice_candidate(P, [mp](int d) { return mp[d]; }, w0 + kf, sps_rt, h, lane, crc_t, fast, rec);
                }
            }
        };
        // first the noise medians of four candidates at a time (the long tasks), then the PPM slices of pairs (half-warp
        // each); the two kinds of task fill disjoint fields of the record.  A pair that needs the generic path (long run,
        // noise window leaving the staged span) is left to its slice task, which then does both roles; so is every pair of
        // an edge tile.
        const int n_quads = (SPS != 0 && fast) ? ((n_cand + 3) >> 2) : 0;
        const int n_tasks = n_quads + ((n_cand + 1) >> 1);
        for (int t = wid; t < n_tasks; t += n_warps) {
            if (t < n_quads) {
                const int c = 4 * t + (lane >> 3);
                const bool have = c < n_cand;
                const unsigned ec = sm.list[have ? c : 4 * t];
                const unsigned ep = (have && (c ^ 1) < n_cand) ? sm.list[c ^ 1] : ec;
                const bool wr = have && !((ec | ep) >> 15) && g_off + (unsigned)(c & ~1) + 1u < P.cand_cap;
                const int kp = wr ? (int)ec : kLeftHalo;          // any staged slot with a full window below it
                slice_quad_snr<FMT>(P, sm.mag + kp, w0 + kp, lane, wr, P.cand[g_off + (have ? c : 0)]);
                continue;
            }
            const int c = (t - n_quads) * 2;
            const bool two = c + 1 < n_cand;
            const unsigned e0 = sm.list[c];
            const unsigned e1 = two ? sm.list[c + 1] : e0;
            if (g_off + (unsigned)c + 1u >= P.cand_cap) {
                if (lane == 0) atomicOr(&P.hdr->flags, kFlagCandOverflow);
                continue;
            }
            if (SPS != 0 && fast && !((e0 | e1) >> 15)) {
                const int kp = (lane & 16) ? (int)e1 : (int)e0;
                const int cc = c + ((lane >> 4) & (two ? 1 : 0));
                const bool wr = two || !(lane & 16);
                slice_pair_bits<SPS ? SPS : 2>(P, sm.mag + kp, lane, wr, P.cand[g_off + cc]);
                continue;
            }
            generic_pair(c, two, e0, e1);
        }
    } else {
        // interior tiles: candidates go in pairs (half-warp each) and every pair is two tasks -- median/SNR and PPM
        // slice/CRC -- spread over the eight warps; edges / long runs: one warp per candidate, both roles
        const int n_tasks = ((n_cand + 1) >> 1) * 2;
        for (int t = wid; t < n_tasks; t += n_warps) {
            const int c = (t >> 1) * 2;
            const bool two = c + 1 < n_cand;
            const unsigned e0 = sm.list[c];
            const unsigned e1 = two ? sm.list[c + 1] : e0;
            if (g_off + (unsigned)c + 1u >= P.cand_cap) {
                if (lane == 0) atomicOr(&P.hdr->flags, kFlagCandOverflow);
                continue;
            }
            if (SPS != 0 && fast && !((e0 | e1) >> 15)) {
                const int kp = (lane & 16) ? (int)e1 : (int)e0;
                const int cc = c + ((lane >> 4) & (two ? 1 : 0));
                const bool wr = two || !(lane & 16);
                Rec &rec = P.cand[g_off + cc];
                if (t & 1)
                    slice_pair_bits<SPS ? SPS : 2>(P, sm.mag + kp, lane, wr, rec);
                else
                    slice_pair_snr<SPS ? SPS : 2, FMT>(P, sm.mag + kp, w0 + kp, lane, wr, rec);
                continue;
            }
            if (t & 1) continue;                            // generic path: the even task does both roles
            unsigned crc_t[4];
#pragma unroll
            for (int j = 0; j < 4; j++) crc_t[j] = (j * 32 + lane < kMsgBits) ? P.crc_tab[j * 32 + lane] : 0u;
            for (int q = 0; q < (two ? 2 : 1); q++) {
                const unsigned e = q ? e1 : e0;
                const int kf = e & 0x7fff;
                Rec &rec = P.cand[g_off + c + q];
                if (e >> 15) {
                    long long p;
                    resolve_long<FMT>(P, sm, w0, kf, lane, p);
                    slice_candidate(P, [&X, p](int d) { return X(p + d); }, p, sps_rt, h, lane, crc_t, false, rec);
                } else {
                    const float *mp = sm.mag + kf;              // list holds the centre slot
                    slice_candidate(P, [mp](int d) { return mp[d]; }, w0 + kf, sps_rt, h, lane, crc_t, fast, rec);
                }
            }
        }
    }
}

// ---- the kernel: one CTA (8 warps) per tile, phases in sequence ---------------------------------------------
template <int SPS, int FMT, int NT = kThreads>
__global__ void __launch_bounds__(NT, ((SPS == 2 || SPS == 4) ? ADSB_MIN_CTAS : ADSB_MIN_CTAS_BIG) * kThreads / NT)
scan_slice_kernel(const ScanParams P)
{
#if defined(ADSB_HOST_EMU)
    unsigned char *smem_raw = cuda_emu::dynamic_smem();
#else
    extern __shared__ __align__(16) unsigned char smem_raw[];
#endif
    const TileSmem sm = carve(smem_raw, Geo<SPS, NT>::T, SPS ? Geo<SPS, NT>::RH : P.rh, Geo<SPS, NT>::CW);
    // CTA 1 <-> the CTA of the launch's final tile (ScanParams::early_tile); all in 32-bit uniform arithmetic
    unsigned bx = blockIdx.x;
    const unsigned eb = P.early_tile;
    const unsigned sw = (bx == 1u) ? eb : ((bx == eb) ? 1u : bx);
    bx = eb ? sw : bx;
    const long long tile = (long long)(P.tile_first + bx);
    const CtaBar gb;
    scan_tile<SPS, FMT, NT>(P, sm, tile, gb);
    gb.sync();
    slice_tile<SPS, FMT, NT>(P, sm, tile, (int)threadIdx.x, NT / 32, gb);
}

}  // namespace adsb
