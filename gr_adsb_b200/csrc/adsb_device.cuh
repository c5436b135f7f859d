// Device-side building blocks of the B200 ADS-B burst path (sm_100a).
//
// K1  scan_slice_kernel   one pass over the sample buffer: |IQ|^2, threshold bitmap, pulse runs and
//                         centres, 16-tap preamble match, and -- for every match ("candidate") --
//                         median-noise SNR, 112-bit PPM slice and CRC-24, all out of one shared-memory
//                         tile.  Replaces the per-sample work of framer.py:83-159 and demod.py:87-95.
// K2  gather / chain / compact   tiny kernels over the sparse candidate list: order it, apply the
//                         63-symbol blanking rule of framer.py:121-123,165 and emit the accepted hits.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

// Kernel launches go through one macro so that the same sources also build for the GPU-less logic tests
// (tests/emu: every CUDA thread a host fiber; test infrastructure, never part of the shipped library).
#if defined(ADSB_HOST_EMU)
#define ADSB_LAUNCH(kern, grid, block, smem, stream, ...) \
    cuda_emu::launch((grid), (block), (smem), [=]() { kern(__VA_ARGS__); })
#else
#define ADSB_LAUNCH(kern, grid, block, smem, stream, ...) kern<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#endif


namespace adsb {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kLeftHalo = 128;      // samples staged left of a tile: 100 (noise window) + run look-back
constexpr int kNoise = 100;         // framer.py:31 NUM_NOISE_SAMPLES
constexpr int kMsgBits = 112;       // demod.py:29
constexpr unsigned kFull = 0xffffffffu;

// flags in Header.flags
constexpr unsigned kFlagRise = 1u, kFlagFall = 2u, kFlagCandOverflow = 4u, kFlagHalo = 8u,
                   kFlagHitOverflow = 16u;

// == adsb_hit (include/adsb_b200.h)
struct __align__(16) Rec {
    long long idx;
    float snr;
    unsigned flags;
    unsigned bits[4];
};
static_assert(sizeof(Rec) == 32, "record must stay 32 bytes");

constexpr unsigned kHitDemodValid = 1u, kHitCrcOk = 2u, kHitLate = 4u, kHitFecFixed = 8u;
constexpr long long kCarryNone = -(1LL << 62);   // "no blanking state" / "no candidate" in the carry exchange

struct Header {
    unsigned long long n_cand;       // candidates allocated by K1 (atomic)
    unsigned long long n_hits;       // accepted hits (K2 compact)
    unsigned long long n_demod;
    unsigned long long n_crc;
    long long last_pulse;            // max pulse centre, window coordinates (-1: none)
    long long last_accept;           // max accepted centre, window coordinates (-1: none)
    unsigned flags;
    unsigned n_fec;                  // hits repaired by the conservative error correction
    unsigned tiles_done;             // latency path: tiles finished so far (the last one resolves the call)
    unsigned pad0;
    float x_last;                    // x[n_out-1]  -> framer.py:87 prev_in0
    unsigned pad;
    long long first_cand;            // first candidate centre, window coordinates (-1: none)
    // multi-GPU blanking carry resolved on the device (global window coordinates)
    long long carry_ein;             // incoming state the next conditional re-resolve must use
    long long carry_used;            // incoming state the latest resolve assumed
    int carry_need;                  // 1: the conditional re-resolve kernels must run
    unsigned carry_unstable;         // rounds whose final check still found a rank needing a fix-up (sticky)
};

struct ScanParams {
    const void *in;        // float (|IQ|^2), float2 (IQ), short2 (sc16 IQ) or uchar2 (cu8 IQ)
    long long n_arr;       // items in `in`
    long long lead;        // items before in[0]: zeros (stream start), or the left neighbour's shard when in_left is set;
                           // window index = array index + lead
    long long n_loc;       // lead + n_arr
    // multi-GPU peer mode (adsb_shard_step): the neighbouring shards, read in place over NVLink by the edge tiles
    const void *in_left;   // window index i in [0, lead) is in_left[left_off + i] (null: zeros)
    long long left_off;
    const void *in_right;  // window index n_loc + j is in_right[j], j < n_right (null: zeros)
    long long n_right;
    long long win_hi;      // n_loc + n_right: end of the samples this call can see
    const unsigned long long *peer_ready;  // this rank's ShardCtrl: [0] left / [1] right neighbour's step counter, [2] own;
                                           // tiles that may touch a neighbour's samples wait until it has reached the step
    long long own_lo;      // falls owned: window index in [own_lo, own_hi)
    long long own_hi;
    long long base;        // added to window index in the emitted records
    long long n_tiles_total;
    float thr;
    int sps;
    int rh;                // right halo, samples (multiple of 32, >= 120*sps)
    int left_is_start;     // window index 0 is where this work() call's window starts
    int prev_bit;          // bitmap value at window index -1 (framer.py:83-84)
    unsigned tile_first;   // first tile of this launch (blockIdx.x is added)
    unsigned n_tiles_launch;  // tiles covered by this launch
    unsigned prefetch_tiles;  // L2 prefetch distance in tiles (0 = off)
    unsigned early_tile;      // the launch's CTA with this index swaps tiles with CTA 1 (0 = off): the stream's
                              // final tile takes the slow edge path (and, in multi-GPU peer mode, reads a neighbour's memory
                              // over NVLink) -- started first, its latency hides behind the interior tiles instead of
                              // forming the kernel's tail
    unsigned fast_lo;      // tiles [fast_lo, fast_lo + fast_span] lie fully inside the array and the owned range
    unsigned fast_span;    // (0xffffffff when there is none)
    int do_demod;
    int hist;              // 8*sps - 1
    Rec *cand;
    unsigned cand_cap;
    uint2 *tile_tab;       // per tile: (offset into cand, count)
    unsigned *tile_chunk_tot;  // candidates per chunk of kThreads tiles (ordered offsets in gather_kernel)
    Header *hdr;
    const unsigned *crc_tab;   // 112 words: x^(111-i) mod g
};

// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float2 ldg_stream_f2(const float2 *p)
{
    float2 v;
#if defined(ADSB_HOST_EMU)
    v = *p;
#else
    asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
#endif
    return v;
}
__device__ __forceinline__ float ldg_stream_f1(const float *p)
{
    float v;
#if defined(ADSB_HOST_EMU)
    v = *p;
#else
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
#endif
    return v;
}

__device__ __forceinline__ unsigned long long ldg_stream_b64(const void *p)
{
    unsigned long long v;
#if defined(ADSB_HOST_EMU)
    memcpy(&v, p, 8);
#else
    asm volatile("ld.global.nc.L1::no_allocate.b64 %0, [%1];" : "=l"(v) : "l"(p));
#endif
    return v;
}

// |IQ|^2 of one complex64 held in a 64-bit register: one packed multiply (Blackwell mul.f32x2, both
// lanes round-to-nearest like two FMULs) + one add -- same value as mag2() below
__device__ __forceinline__ float mag2_packed(unsigned long long v)
{
#if defined(ADSB_HOST_EMU)
    float2 c;
    memcpy(&c, &v, 8);
    return __fadd_rn(__fmul_rn(c.x, c.x), __fmul_rn(c.y, c.y));
#else
    unsigned long long sq;
    asm("mul.rn.f32x2 %0, %1, %1;" : "=l"(sq) : "l"(v));
    float a, b;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(sq));
    return __fadd_rn(a, b);
#endif
}

// |IQ|^2 exactly as the oracle: fl32(fl32(re*re) + fl32(im*im)), no FMA contraction (adsb_rx.py:180)
__device__ __forceinline__ float mag2(float2 v)
{
    return __fadd_rn(__fmul_rn(v.x, v.x), __fmul_rn(v.y, v.y));
}

// input sample formats of the scan kernel
enum { kFmtF32 = 0,     // float32 |IQ|^2 (what framer.work() is handed)
       kFmtC64 = 1,     // complex64 IQ (the USRP source's fc32 stream)
       kFmtSC16 = 2,    // interleaved int16 IQ, value = int * 2^-15 (the wire format of most SDRs)
       kFmtCU8 = 3 };   // interleaved uint8 IQ, value = (u - 127.5) / 128 (RTL-SDR dongles, the usual 1090 MHz receiver)
constexpr int kNumFmt = 4;

template <int FMT>
__host__ __device__ constexpr int item_bytes()
{
    return FMT == kFmtC64 ? 8 : FMT == kFmtCU8 ? 2 : 4;
}

// |IQ|^2 of one sc16 sample (re in the low half): widen, scale by 2^-15 (exact), then as mag2()
__device__ __forceinline__ float mag2_sc16(unsigned v)
{
    const float re = __fmul_rn(__int2float_rn((int)(short)(v & 0xffffu)), 3.0517578125e-05f);
    const float im = __fmul_rn(__int2float_rn((int)(short)(v >> 16)), 3.0517578125e-05f);
    return __fadd_rn(__fmul_rn(re, re), __fmul_rn(im, im));
}

// |IQ|^2 of one cu8 sample (re in the low byte): (u - 127.5) / 128 is exact in float32, then as mag2()
__device__ __forceinline__ float mag2_cu8(unsigned v)
{
    const float re = __fmul_rn(__fadd_rn(__int2float_rn((int)(v & 0xffu)), -127.5f), 0.0078125f);
    const float im = __fmul_rn(__fadd_rn(__int2float_rn((int)((v >> 8) & 0xffu)), -127.5f), 0.0078125f);
    return __fadd_rn(__fmul_rn(re, re), __fmul_rn(im, im));
}

__device__ __forceinline__ unsigned ldg_stream_b16(const void *p)
{
    unsigned short v;
#if defined(ADSB_HOST_EMU)
    memcpy(&v, p, 2);
#else
    asm volatile("ld.global.nc.L1::no_allocate.b16 %0, [%1];" : "=h"(v) : "l"(p));
#endif
    return v;
}

__device__ __forceinline__ unsigned ldg_stream_b32(const void *p)
{
    unsigned v;
#if defined(ADSB_HOST_EMU)
    memcpy(&v, p, 4);
#else
    asm volatile("ld.global.nc.L1::no_allocate.b32 %0, [%1];" : "=r"(v) : "l"(p));
#endif
    return v;
}

template <int FMT>
__device__ __forceinline__ float load_x(const void *in, long long j)
{
    if (FMT == kFmtC64) return mag2(ldg_stream_f2(reinterpret_cast<const float2 *>(in) + j));
    if (FMT == kFmtSC16) return mag2_sc16(ldg_stream_b32(reinterpret_cast<const unsigned *>(in) + j));
    if (FMT == kFmtCU8) return mag2_cu8(ldg_stream_b16(reinterpret_cast<const unsigned short *>(in) + j));
    return ldg_stream_f1(reinterpret_cast<const float *>(in) + j);
}

// raw bits of item j / their |IQ|^2: split so that a batch of loads can be issued before the first conversion waits for one
template <int FMT>
__device__ __forceinline__ unsigned long long load_raw(const void *in, long long j)
{
    if (FMT == kFmtC64) return ldg_stream_b64(reinterpret_cast<const float2 *>(in) + j);
    if (FMT == kFmtSC16) return ldg_stream_b32(reinterpret_cast<const unsigned *>(in) + j);
    if (FMT == kFmtCU8) return ldg_stream_b16(reinterpret_cast<const unsigned short *>(in) + j);
    return ldg_stream_b32(reinterpret_cast<const unsigned *>(in) + j);
}
template <int FMT>
__device__ __forceinline__ float raw_to_x(unsigned long long r)
{
    if (FMT == kFmtC64) return mag2_packed(r);
    if (FMT == kFmtSC16) return mag2_sc16((unsigned)r);
    if (FMT == kFmtCU8) return mag2_cu8((unsigned)r);
    return __uint_as_float((unsigned)r);
}

// where window sample i lives: the array, a neighbouring shard, or nowhere (zeros); returns false for "nowhere"
template <int FMT>
__device__ __forceinline__ bool window_raw(const ScanParams &P, long long i, unsigned long long &raw)
{
    const long long j = i - P.lead;
    if (j >= 0 && j < P.n_arr) {
        raw = load_raw<FMT>(P.in, j);
        return true;
    }
    if (j < 0) {
        if (!(P.in_left && i >= 0)) return false;
        raw = load_raw<FMT>(P.in_left, P.left_off + i);
        return true;
    }
    const long long jr = j - P.n_arr;
    if (!(P.in_right && jr < P.n_right)) return false;
    raw = load_raw<FMT>(P.in_right, jr);
    return true;
}

// window sample i (any i): zeros outside the array / inside the virtual lead
template <int FMT>
__device__ __forceinline__ float window_x(const ScanParams &P, long long i)
{
    const long long j = i - P.lead;
    if (j >= 0 && j < P.n_arr) return load_x<FMT>(P.in, j);
    if (j < 0) return (P.in_left && i >= 0) ? load_x<FMT>(P.in_left, P.left_off + i) : 0.0f;
    const long long jr = j - P.n_arr;
    return (P.in_right && jr < P.n_right) ? load_x<FMT>(P.in_right, jr) : 0.0f;
}

// ---- bounded waits on flags other GPUs write (multi-GPU peer mode) ---------------------------------------------
constexpr unsigned kFlagPeerTimeout = 32u;
constexpr unsigned long long kPeerTimeoutNs = 2000000000ULL;
__device__ __forceinline__ unsigned long long now_ns()
{
#if defined(ADSB_HOST_EMU)
    static unsigned long long t = 0;
    return t += 1000000ULL;            // every look at the clock is a millisecond: a wait that cannot succeed ends at once
#else
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
#endif
}
__device__ __forceinline__ void fence_system()
{
#if !defined(ADSB_HOST_EMU)
    __threadfence_system();
#endif
}
__device__ __forceinline__ unsigned long long vload64(const unsigned long long *p)
{
    return *reinterpret_cast<const volatile unsigned long long *>(p);
}
// wait until *p >= want (step counters only grow); false on timeout
__device__ __forceinline__ bool wait_ge(const unsigned long long *p, unsigned long long want)
{
    if (vload64(p) >= want) return true;
    const unsigned long long t0 = now_ns();
    while (vload64(p) < want) {
        if (now_ns() - t0 > kPeerTimeoutNs) return false;
#if !defined(ADSB_HOST_EMU)
        __nanosleep(100);
#endif
    }
    return true;
}

// total order on non-NaN floats as unsigned keys
__device__ __forceinline__ unsigned f2key(float x)
{
    unsigned u = __float_as_uint(x);
    return (u >> 31) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key2f(unsigned k)
{
    return __uint_as_float((k >> 31) ? (k & 0x7fffffffu) : ~k);
}

// 10*log10(ratio) + 1.6 with the reference's float32 steps (framer.py:157-159 under NumPy >= 2);
// ratio = fl32(x[p] / median)
__device__ __forceinline__ float snr_db_from_ratio(float ratio)
{
    float l = log10f(ratio);
    return __fadd_rn(__fmul_rn(10.0f, l), 1.6f);
}

// ---------------------------------------------------------------------------------------------
// Warp-wide bitonic sort of 128 keys, 4 per lane, logical position e = 4*lane + r, ascending.
// "Mirrored" formulation: the first step of every merge pairs e with e ^ (size-1), the rest pair e with
// e ^ j, and every comparator is ascending -- no direction flags, so an in-lane exchange is min + max.
__device__ __forceinline__ void cex(unsigned &a, unsigned &b)
{
    const unsigned lo = min(a, b), hi = max(a, b);
    a = lo;
    b = hi;
}

__device__ __forceinline__ void warp_sort128(unsigned (&k)[4], int lane)
{
#pragma unroll
    for (int size = 2; size <= 128; size <<= 1) {
        // mirror step
        if (size == 2) {
            cex(k[0], k[1]);
            cex(k[2], k[3]);
        } else if (size == 4) {
            cex(k[0], k[3]);
            cex(k[1], k[2]);
        } else {
            const int lm = (size >> 2) - 1;
            const bool lower = (lane & (size >> 3)) == 0;
            const unsigned o0 = __shfl_xor_sync(kFull, k[3], lm);
            const unsigned o1 = __shfl_xor_sync(kFull, k[2], lm);
            const unsigned o2 = __shfl_xor_sync(kFull, k[1], lm);
            const unsigned o3 = __shfl_xor_sync(kFull, k[0], lm);
            k[0] = lower ? min(k[0], o0) : max(k[0], o0);
            k[1] = lower ? min(k[1], o1) : max(k[1], o1);
            k[2] = lower ? min(k[2], o2) : max(k[2], o2);
            k[3] = lower ? min(k[3], o3) : max(k[3], o3);
        }
#pragma unroll
        for (int j = size >> 2; j > 0; j >>= 1) {
            if (j >= 4) {
                const int lm = j >> 2;
                const bool lower = (lane & lm) == 0;
#pragma unroll
                for (int r = 0; r < 4; r++) {
                    const unsigned o = __shfl_xor_sync(kFull, k[r], lm);
                    k[r] = lower ? min(k[r], o) : max(k[r], o);
                }
            } else if (j == 2) {
                cex(k[0], k[2]);
                cex(k[1], k[3]);
            } else {
                cex(k[0], k[1]);
                cex(k[2], k[3]);
            }
        }
    }
}

// Two independent 128-key sets per warp: lanes 0-15 hold one, lanes 16-31 the other; 8 keys per lane, logical
// position e = 8*(lane & 15) + r, same mirrored bitonic network as warp_sort128 (only the steps with a stride of 8
// or more cross lanes, which is what makes slicing candidates in pairs cheaper than one per warp).
// The two middle order statistics (ranks 63 and 64 of 128) of each set, without finishing the sort: sort the two
// 64-key halves A (lanes 0-7) and B (lanes 8-15) with sizes 2..64 of the network, then use that
// {min(A[i], B[63-i])} are exactly the 64 smallest keys and {max(A[i], B[63-i])} the 64 largest -- rank 63 is the
// largest of the former, rank 64 the smallest of the latter.  Replaces the last 7 of 28 network steps by one
// exchange and two reductions.  `top` = the largest key (last element of A or of B).
__device__ __forceinline__ void half_mid128(unsigned (&k)[8], int lane, unsigned &lo_mid, unsigned &hi_mid, unsigned &top)
{
    // each lane's 8 keys ascending
    // Batcher's odd-even merge sort of 8: 19 comparators instead of the bitonic network's 24
    cex(k[0], k[1]); cex(k[2], k[3]); cex(k[4], k[5]); cex(k[6], k[7]);
    cex(k[0], k[2]); cex(k[1], k[3]); cex(k[4], k[6]); cex(k[5], k[7]);
    cex(k[1], k[2]); cex(k[5], k[6]);
    cex(k[0], k[4]); cex(k[1], k[5]); cex(k[2], k[6]); cex(k[3], k[7]);
    cex(k[2], k[4]); cex(k[3], k[5]);
    cex(k[1], k[2]); cex(k[3], k[4]); cex(k[5], k[6]);
#pragma unroll
    for (int size = 16; size <= 64; size <<= 1) {
        {
            const int lm = (size >> 3) - 1;
            const bool lower = (lane & (size >> 4)) == 0;
            unsigned o[8];
#pragma unroll
            for (int r = 0; r < 8; r++) o[r] = __shfl_xor_sync(kFull, k[7 - r], lm);
#pragma unroll
            for (int r = 0; r < 8; r++) k[r] = lower ? min(k[r], o[r]) : max(k[r], o[r]);
        }
#pragma unroll
        for (int j = size >> 2; j >= 8; j >>= 1) {
            const int lm = j >> 3;
            const bool lower = (lane & lm) == 0;
#pragma unroll
            for (int r = 0; r < 8; r++) {
                const unsigned o = __shfl_xor_sync(kFull, k[r], lm);
                k[r] = lower ? min(k[r], o) : max(k[r], o);
            }
        }
        cex(k[0], k[4]); cex(k[1], k[5]); cex(k[2], k[6]); cex(k[3], k[7]);
        cex(k[0], k[2]); cex(k[1], k[3]); cex(k[4], k[6]); cex(k[5], k[7]);
#pragma unroll
        for (int r = 0; r < 8; r += 2) cex(k[r], k[r + 1]);
    }
    // A = positions 0..63 (lanes 0-7 of the half-warp), B = 64..127 (lanes 8-15), both ascending
    const unsigned ta = __shfl_sync(kFull, k[7], (lane & 16) | 7), tb = __shfl_sync(kFull, k[7], (lane & 16) | 15);
    top = max(ta, tb);
    const bool lower = (lane & 8) == 0;
    unsigned m = 0;
#pragma unroll
    for (int r = 0; r < 8; r++) {
        const unsigned o = __shfl_xor_sync(kFull, k[7 - r], 15);       // position e <-> 127 - e
        const unsigned v = lower ? min(k[r], o) : max(k[r], o);
        m = (r == 0) ? v : (lower ? max(m, v) : min(m, v));
    }
#pragma unroll
    for (int d = 1; d <= 4; d <<= 1) {
        const unsigned o = __shfl_xor_sync(kFull, m, d);
        m = lower ? max(m, o) : min(m, o);
    }
    lo_mid = __shfl_sync(kFull, m, lane & 16);
    hi_mid = __shfl_sync(kFull, m, (lane & 16) | 8);
}

// Four independent 128-key sets per warp: lanes 8c .. 8c+7 hold set c, 16 keys per lane, logical position
// e = 16*(lane & 7) + r, same mirrored bitonic network as warp_sort128.  Only the strides of 16 and more cross lanes:
// 64 shuffles for four sets (two sets per warp with 8 keys per lane took 63 for two -- and the shuffles were a quarter
// of the scan kernel's L1/shared data-pipe traffic).  In-lane: Batcher's odd-even merge sort of 16 (63 comparators),
// then 32-comparator bitonic merges.
// The two middle order statistics (ranks 63 and 64 of 128) of each set, without finishing the sort: sort the two
// 64-key halves A (positions 0..63, lanes 0-3 of the group) and B (64..127, lanes 4-7) with sizes 2..64 of the network,
// then use that {min(A[i], B[63-i])} are exactly the 64 smallest keys and {max(A[i], B[63-i])} the 64 largest -- rank 63
// is the largest of the former, rank 64 the smallest of the latter.  Replaces the last 7 of 28 network steps by one
// exchange and two reductions.  Lanes 0-3 of a group return rank 63, lanes 4-7 rank 64; lane_top = the lane's largest key.
__device__ __forceinline__ unsigned quarter_mid128(unsigned (&k)[16], int lane, unsigned &lane_top)
{
    cex(k[0], k[1]); cex(k[2], k[3]); cex(k[4], k[5]); cex(k[6], k[7]); cex(k[8], k[9]); cex(k[10], k[11]); cex(k[12], k[13]); cex(k[14], k[15]);
    cex(k[0], k[2]); cex(k[1], k[3]); cex(k[4], k[6]); cex(k[5], k[7]); cex(k[8], k[10]); cex(k[9], k[11]); cex(k[12], k[14]); cex(k[13], k[15]);
    cex(k[1], k[2]); cex(k[5], k[6]); cex(k[0], k[4]); cex(k[3], k[7]); cex(k[9], k[10]); cex(k[13], k[14]); cex(k[8], k[12]); cex(k[11], k[15]);
    cex(k[2], k[6]); cex(k[1], k[5]); cex(k[10], k[14]); cex(k[9], k[13]); cex(k[0], k[8]); cex(k[7], k[15]);
    cex(k[2], k[4]); cex(k[3], k[5]); cex(k[10], k[12]); cex(k[11], k[13]);
    cex(k[1], k[2]); cex(k[3], k[4]); cex(k[5], k[6]); cex(k[9], k[10]); cex(k[11], k[12]); cex(k[13], k[14]);
    cex(k[4], k[12]); cex(k[2], k[10]); cex(k[6], k[14]); cex(k[1], k[9]); cex(k[5], k[13]); cex(k[3], k[11]);
    cex(k[4], k[8]); cex(k[6], k[10]); cex(k[5], k[9]); cex(k[7], k[11]);
    cex(k[2], k[4]); cex(k[6], k[8]); cex(k[10], k[12]); cex(k[3], k[5]); cex(k[7], k[9]); cex(k[11], k[13]);
    cex(k[1], k[2]); cex(k[3], k[4]); cex(k[5], k[6]); cex(k[7], k[8]); cex(k[9], k[10]); cex(k[11], k[12]); cex(k[13], k[14]);
    lane_top = k[15];                                            // this lane's largest key
#pragma unroll
    for (int size = 32; size <= 64; size <<= 1) {
        {
            // mirror step: position e <-> e ^ (size - 1); registers go in pairs (r, 15 - r) so that two temporaries suffice
            const int lm = (size >> 4) - 1;
            const bool lower = (lane & (size >> 5)) == 0;
#pragma unroll
            for (int r = 0; r < 8; r++) {
                const unsigned oh = __shfl_xor_sync(kFull, k[15 - r], lm);
                const unsigned ol = __shfl_xor_sync(kFull, k[r], lm);
                k[r] = lower ? min(k[r], oh) : max(k[r], oh);
                k[15 - r] = lower ? min(k[15 - r], ol) : max(k[15 - r], ol);
            }
        }
        if (size == 64) {
            const bool lower = (lane & 1) == 0;                  // stride 16: the neighbouring lane, same register
#pragma unroll
            for (int r = 0; r < 16; r++) {
                const unsigned o = __shfl_xor_sync(kFull, k[r], 1);
                k[r] = lower ? min(k[r], o) : max(k[r], o);
            }
        }
#pragma unroll
        for (int j = 8; j > 0; j >>= 1)
#pragma unroll
            for (int r = 0; r < 16; r++)
                if (!(r & j)) cex(k[r], k[r | j]);
    }
    const bool lower = (lane & 4) == 0;
    unsigned m = 0;
#pragma unroll
    for (int r = 0; r < 16; r++) {
        const unsigned o = __shfl_xor_sync(kFull, k[15 - r], 7);       // position e <-> 127 - e
        const unsigned v = lower ? min(k[r], o) : max(k[r], o);
        m = (r == 0) ? v : (lower ? max(m, v) : min(m, v));
    }
#pragma unroll
    for (int d = 1; d <= 2; d <<= 1) {
        const unsigned o = __shfl_xor_sync(kFull, m, d);
        m = lower ? max(m, o) : min(m, o);
    }
    return m;
}

__device__ __forceinline__ unsigned pick4(const unsigned (&k)[4], int r)
{
    unsigned v = k[0];
    v = (r == 1) ? k[1] : v;
    v = (r == 2) ? k[2] : v;
    v = (r == 3) ? k[3] : v;
    return v;
}

// np.median of n <= 100 floats held 4 per lane (any order; slots >= n ignored); all lanes get the result
__device__ __forceinline__ float warp_median_regs(const float (&v)[4], int n, int lane)
{
    unsigned k[4];
    bool nan = false;
#pragma unroll
    for (int r = 0; r < 4; r++) {
        const int e = r * 32 + lane;
        nan |= (e < n) && (v[r] != v[r]);
        k[r] = (e < n) ? f2key(v[r]) : 0xffffffffu;
    }
    if (__any_sync(kFull, nan) || n <= 0)
        return __int_as_float(0x7fc00000);
    warp_sort128(k, lane);
    const int k1 = (n - 1) >> 1, k2 = n >> 1;
    unsigned a = __shfl_sync(kFull, pick4(k, k1 & 3), k1 >> 2);
    unsigned b = __shfl_sync(kFull, pick4(k, k2 & 3), k2 >> 2);
    float fa = key2f(a), fb = key2f(b);
    return (n & 1) ? fa : __fdiv_rn(__fadd_rn(fa, fb), 2.0f);
}

// message bits (ballot word: bit l = message bit 32*j + l) -> numpy.packbits byte order
__device__ __forceinline__ unsigned pack_msb_first(unsigned w)
{
    return __byte_perm(__brev(w), 0, 0x0123);
}

// block-wide exclusive scan of one int per thread (kThreads threads); returns exclusive prefix,
// total in `total`.  `scratch` = kWarps + 1 ints of shared memory.
__device__ __forceinline__ int block_excl_scan(int v, int *scratch, int &total)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int t = __shfl_up_sync(kFull, inc, d);
        if (lane >= d) inc += t;
    }
    if (lane == 31) scratch[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        int s = (lane < kWarps) ? scratch[lane] : 0;
#pragma unroll
        for (int d = 1; d < kWarps; d <<= 1) {
            int t = __shfl_up_sync(kFull, s, d);
            if (lane >= d) s += t;
        }
        if (lane < kWarps) scratch[lane] = s;
    }
    __syncthreads();
    total = scratch[kWarps - 1];
    int off = (wid ? scratch[wid - 1] : 0) + inc - v;
    __syncthreads();
    return off;
}

}  // namespace adsb
