// TEST INFRASTRUCTURE ONLY -- a functional stand-in for <cuda_runtime.h> so that the CUDA sources of
// gr_adsb_b200/csrc compile with g++ and run on a GPU-less machine under tests/emu (see tests/emu/README.md).
// Every CUDA thread of a block is a fiber; warp collectives and block barriers are rendezvous points between
// fibers, blocks run one after the other.  It checks kernel LOGIC (indexing, masks, orderings, barriers under
// a shuffled thread schedule); it says nothing about speed and is never loaded by the gr_adsb_b200 package.
#pragma once
#ifndef ADSB_HOST_EMU
#error "tests/emu/include/cuda_runtime.h is only for the host-emulation build (-DADSB_HOST_EMU)"
#endif
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <functional>

// ---- qualifiers ------------------------------------------------------------------------------------------
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))
// blocks run one after the other on one OS thread, so a function-local static is "shared by the block";
// dynamic shared memory (`extern __shared__`) is cuda_emu::dynamic_smem() instead
#define __shared__ static

// ---- vector types ------------------------------------------------------------------------------------------
struct float2 { float x, y; };
struct __attribute__((aligned(16))) float4 { float x, y, z, w; };
struct __attribute__((aligned(8))) uint2 { unsigned x, y; };
struct __attribute__((aligned(16))) uint4 { unsigned x, y, z, w; };
struct uint3 { unsigned x, y, z; };
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
static inline uint2 make_uint2(unsigned x, unsigned y) { uint2 v; v.x = x; v.y = y; return v; }
static inline float2 make_float2(float x, float y) { float2 v; v.x = x; v.y = y; return v; }

namespace cuda_emu {
extern uint3 g_threadIdx, g_blockIdx;
extern dim3 g_blockDim, g_gridDim;
void launch(dim3 grid, dim3 block, size_t smem_bytes, const std::function<void()> &body);
// rendezvous of the calling fiber's warp / block; see cuda_emu.cpp
uint64_t *warp_exchange(int op, unsigned mask, uint64_t v, unsigned *present);
int block_barrier(int pred, int mode);
int lane_id();
unsigned char *dynamic_smem();
void yield();                     // a spinning thread lets the others run (stands in for __nanosleep in wait loops)
}  // namespace cuda_emu

#define threadIdx (cuda_emu::g_threadIdx)
#define blockIdx (cuda_emu::g_blockIdx)
#define blockDim (cuda_emu::g_blockDim)
#define gridDim (cuda_emu::g_gridDim)

// ---- runtime API: "device" memory is host memory, streams run in program order -------------------------------
typedef int cudaError_t;
typedef struct emu_stream_ *cudaStream_t;
typedef struct emu_event_ *cudaEvent_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2, cudaErrorInvalidValue = 1 };
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2 };
enum { cudaFuncAttributeMaxDynamicSharedMemorySize = 8, cudaFuncAttributePreferredSharedMemoryCarveout = 9 };
enum { cudaSharedmemCarveoutMaxShared = 100 };
struct cudaDeviceProp { int major, minor, multiProcessorCount; };

static inline const char *cudaGetErrorString(cudaError_t e) { return e ? "emulated CUDA error" : "no error"; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaGetDeviceCount(int *n) { *n = 1; return cudaSuccess; }
static inline cudaError_t cudaGetDevice(int *d) { *d = 0; return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int)
{
    p->major = 10; p->minor = 0; p->multiProcessorCount = 1;
    return cudaSuccess;
}
// "device" allocations carry guard bands that cuda_emu.cpp checks after every kernel launch and at cudaFree:
// a kernel writing outside a buffer aborts the test with the buffer's size (a memcheck for out-of-bounds writes)
namespace cuda_emu {
void *guarded_alloc(size_t n);
void guarded_free(void *p);
}
static inline cudaError_t cudaMalloc(void **p, size_t n)
{
    *p = cuda_emu::guarded_alloc(n);
    return *p ? cudaSuccess : cudaErrorMemoryAllocation;
}
static inline cudaError_t cudaFree(void *p) { cuda_emu::guarded_free(p); return cudaSuccess; }
static inline cudaError_t cudaMallocHost(void **p, size_t n) { return cudaMalloc(p, n); }
static inline cudaError_t cudaFreeHost(void *p) { cuda_emu::guarded_free(p); return cudaSuccess; }
enum { cudaHostAllocMapped = 2, cudaErrorNotReady = 600 };
static inline cudaError_t cudaHostAlloc(void **p, size_t n, unsigned) { return cudaMalloc(p, n); }
static inline cudaError_t cudaStreamQuery(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) { memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t) { memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemset(void *d, int v, size_t n) { memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t) { memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { *s = (cudaStream_t)malloc(1); return cudaSuccess; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t s) { free(s); return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
static inline cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = (cudaEvent_t)malloc(1); return cudaSuccess; }
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { return cudaEventCreate(e); }
static inline cudaError_t cudaEventDestroy(cudaEvent_t e) { free(e); return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return cudaSuccess; }
template <class F> static inline cudaError_t cudaFuncSetAttribute(F *, int, int) { return cudaSuccess; }

// ---- scalar intrinsics ---------------------------------------------------------------------------------------
// compiled with -ffp-contract=off: a*b, a+b, a/b are the IEEE round-to-nearest single operations
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline float __int2float_rn(int v) { return (float)v; }
static inline unsigned __float_as_uint(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
static inline int __float_as_int(float f) { int u; memcpy(&u, &f, 4); return u; }
static inline float __uint_as_float(unsigned u) { float f; memcpy(&f, &u, 4); return f; }
static inline float __int_as_float(int u) { float f; memcpy(&f, &u, 4); return f; }
static inline int __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline unsigned __brev(unsigned x)
{
    x = ((x >> 1) & 0x55555555u) | ((x & 0x55555555u) << 1);
    x = ((x >> 2) & 0x33333333u) | ((x & 0x33333333u) << 2);
    x = ((x >> 4) & 0x0f0f0f0fu) | ((x & 0x0f0f0f0fu) << 4);
    return __builtin_bswap32(x);
}
static inline unsigned __byte_perm(unsigned a, unsigned b, unsigned s)
{
    const uint64_t v = ((uint64_t)b << 32) | a;
    unsigned r = 0;
    for (int i = 0; i < 4; i++) {
        const unsigned sel = (s >> (4 * i)) & 0xf;
        unsigned byte = (unsigned)(v >> (8 * (sel & 7))) & 0xffu;
        if (sel & 8) byte = (byte & 0x80u) ? 0xffu : 0u;
        r |= byte << (8 * i);
    }
    return r;
}
static inline unsigned __funnelshift_l(unsigned lo, unsigned hi, unsigned shift)
{
    const uint64_t v = ((uint64_t)hi << 32) | lo;
    return (unsigned)((v << (shift & 31)) >> 32);
}
static inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned shift)
{
    const uint64_t v = ((uint64_t)hi << 32) | lo;
    return (unsigned)(v >> (shift & 31));
}
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
static inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }
static inline long long min(long long a, long long b) { return a < b ? a : b; }
static inline long long max(long long a, long long b) { return a > b ? a : b; }

// ---- atomics (fibers never pre-empt each other) ------------------------------------------------------------------
template <class T> static inline T atomicAdd(T *p, T v) { T o = *p; *p = o + v; return o; }
static inline int atomicCAS(int *p, int cmp, int v) { int o = *p; if (o == cmp) *p = v; return o; }
static inline void __threadfence_block() {}
static inline void __threadfence() {}
static inline void __threadfence_system() {}
static inline unsigned atomicOr(unsigned *p, unsigned v) { unsigned o = *p; *p = o | v; return o; }
static inline int atomicMax(int *p, int v) { int o = *p; if (v > o) *p = v; return o; }
static inline long long atomicMax(long long *p, long long v) { long long o = *p; if (v > o) *p = v; return o; }

// ---- warp collectives -----------------------------------------------------------------------------------------------
namespace cuda_emu {
enum { OP_BALLOT = 1, OP_SHFL, OP_SHFL_XOR, OP_SHFL_UP, OP_ANY, OP_RED_ADD, OP_RED_OR, OP_RED_XOR, OP_RED_MAX,
       OP_SYNCWARP };
template <class T> static inline uint64_t to_bits(T v) { uint64_t b = 0; memcpy(&b, &v, sizeof(T)); return b; }
template <class T> static inline T from_bits(uint64_t b) { T v; memcpy(&v, &b, sizeof(T)); return v; }
}  // namespace cuda_emu

static inline unsigned __ballot_sync(unsigned mask, int pred)
{
    unsigned present;
    const uint64_t *v = cuda_emu::warp_exchange(cuda_emu::OP_BALLOT, mask, pred ? 1 : 0, &present);
    unsigned r = 0;
    for (int l = 0; l < 32; l++)
        if (((present >> l) & 1u) && v[l]) r |= 1u << l;
    return r;
}
static inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0; }
static inline void __syncwarp(unsigned mask = 0xffffffffu)
{
    unsigned present;
    cuda_emu::warp_exchange(cuda_emu::OP_SYNCWARP, mask, 0, &present);
}
template <class T> static inline T __shfl_sync(unsigned mask, T v, int src, int width = 32)
{
    unsigned present;
    const uint64_t *a = cuda_emu::warp_exchange(cuda_emu::OP_SHFL, mask, cuda_emu::to_bits(v), &present);
    const int lane = cuda_emu::lane_id();
    const int s = (lane & ~(width - 1)) | (src & (width - 1));
    return cuda_emu::from_bits<T>(a[s]);
}
template <class T> static inline T __shfl_xor_sync(unsigned mask, T v, int lm, int width = 32)
{
    unsigned present;
    const uint64_t *a = cuda_emu::warp_exchange(cuda_emu::OP_SHFL_XOR, mask, cuda_emu::to_bits(v), &present);
    const int lane = cuda_emu::lane_id();
    const int s = lane ^ lm;
    // a source outside the caller's width-segment returns the caller's own value
    return cuda_emu::from_bits<T>(a[((s & ~(width - 1)) == (lane & ~(width - 1))) ? s : lane]);
}
template <class T> static inline T __shfl_up_sync(unsigned mask, T v, unsigned d, int width = 32)
{
    unsigned present;
    const uint64_t *a = cuda_emu::warp_exchange(cuda_emu::OP_SHFL_UP, mask, cuda_emu::to_bits(v), &present);
    const int lane = cuda_emu::lane_id();
    const int s = lane - (int)d;
    return cuda_emu::from_bits<T>(a[(s >= (lane & ~(width - 1))) ? s : lane]);
}
template <class T> static inline T emu_reduce(int op, unsigned mask, T v)
{
    unsigned present;
    const uint64_t *a = cuda_emu::warp_exchange(op, mask, cuda_emu::to_bits(v), &present);
    bool first = true;
    T r = T();
    for (int l = 0; l < 32; l++) {
        if (!((present >> l) & 1u)) continue;
        const T x = cuda_emu::from_bits<T>(a[l]);
        if (first) { r = x; first = false; continue; }
        switch (op) {
        case cuda_emu::OP_RED_ADD: r = (T)(r + x); break;
        case cuda_emu::OP_RED_OR: r = (T)(r | x); break;
        case cuda_emu::OP_RED_XOR: r = (T)(r ^ x); break;
        default: r = x > r ? x : r; break;
        }
    }
    return r;
}
static inline int __reduce_add_sync(unsigned m, int v) { return emu_reduce(cuda_emu::OP_RED_ADD, m, v); }
static inline unsigned __reduce_add_sync(unsigned m, unsigned v) { return emu_reduce(cuda_emu::OP_RED_ADD, m, v); }
static inline unsigned __reduce_or_sync(unsigned m, unsigned v) { return emu_reduce(cuda_emu::OP_RED_OR, m, v); }
static inline unsigned __reduce_xor_sync(unsigned m, unsigned v) { return emu_reduce(cuda_emu::OP_RED_XOR, m, v); }
static inline int __reduce_max_sync(unsigned m, int v) { return emu_reduce(cuda_emu::OP_RED_MAX, m, v); }
static inline unsigned __reduce_max_sync(unsigned m, unsigned v) { return emu_reduce(cuda_emu::OP_RED_MAX, m, v); }

// ---- block barriers ------------------------------------------------------------------------------------------------
static inline void __syncthreads() { cuda_emu::block_barrier(0, 0); }
static inline int __syncthreads_or(int p) { return cuda_emu::block_barrier(p, 1); }
static inline int __syncthreads_and(int p) { return cuda_emu::block_barrier(p, 2); }
static inline int __syncthreads_count(int p) { return cuda_emu::block_barrier(p, 3); }
