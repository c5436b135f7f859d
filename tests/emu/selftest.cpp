// Self-test of the host emulation's collectives against their CUDA definitions (run by tests/test_emu_engine.py).
#include <stdio.h>

#include <vector>

#include "cuda_runtime.h"

static int g_fail = 0;
#define EXPECT(cond)                                                        \
    do {                                                                    \
        if (!(cond)) {                                                      \
            g_fail++;                                                       \
            fprintf(stderr, "FAILED line %d: %s\n", __LINE__, #cond);      \
        }                                                                   \
    } while (0)

__global__ void k_collectives(unsigned *out)
{
    const int tid = threadIdx.x, lane = tid & 31;
    const unsigned full = 0xffffffffu;
    unsigned *o = out + tid * 16;
    o[0] = __ballot_sync(full, lane % 3 == 0);
    o[1] = __shfl_sync(full, tid * 7u, 5);
    o[2] = __shfl_xor_sync(full, tid * 7u, 9);
    o[3] = __shfl_up_sync(full, tid * 7u, 3);
    o[4] = __shfl_sync(full, tid * 7u, 3, 16);            // width 16: source lane 3 of the caller's half
    o[5] = (unsigned)__reduce_add_sync(full, lane);
    o[6] = __reduce_or_sync(full, 1u << (lane & 7));
    o[7] = (unsigned)__reduce_max_sync(full, (int)((lane * 37) % 29));
    o[8] = __reduce_xor_sync(full, (unsigned)lane);
    o[9] = (unsigned)__any_sync(full, tid == 77);
    o[10] = (unsigned)__syncthreads_or(tid == 200);
    o[11] = (unsigned)__syncthreads_or(0);
    o[12] = (unsigned)__syncthreads_count(tid % 5 == 0);
    __shared__ int slot[8];
    if (lane == 0) slot[tid >> 5] = tid;
    __syncthreads();
    o[13] = (unsigned)slot[((tid >> 5) + 1) & 7];          // another warp's slot, visible after the barrier
}

__global__ void k_early_exit(int *out)
{
    if (threadIdx.x >= 40) return;                        // exited threads leave barriers and collectives
    __syncthreads();
    const unsigned m = __ballot_sync(0xffffffffu, 1);
    out[threadIdx.x] = (int)m;
}

int main()
{
    std::vector<unsigned> out(256 * 16);
    unsigned *p = out.data();
    cuda_emu::launch(1, 256, 0, [=]() { k_collectives(p); });
    for (int tid = 0; tid < 256; tid++) {
        const int lane = tid & 31, w0 = tid & ~31;
        const unsigned *o = p + tid * 16;
        unsigned bal = 0, orv = 0, xr = 0;
        int sum = 0, mx = -1;
        for (int l = 0; l < 32; l++) {
            if (l % 3 == 0) bal |= 1u << l;
            sum += l;
            orv |= 1u << (l & 7);
            xr ^= (unsigned)l;
            const int v = (l * 37) % 29;
            mx = v > mx ? v : mx;
        }
        EXPECT(o[0] == bal);
        EXPECT(o[1] == (unsigned)(w0 + 5) * 7u);
        EXPECT(o[2] == (unsigned)(w0 + (lane ^ 9)) * 7u);
        EXPECT(o[3] == (unsigned)(lane >= 3 ? tid - 3 : tid) * 7u);
        EXPECT(o[4] == (unsigned)(w0 + (lane & 16) + 3) * 7u);
        EXPECT(o[5] == (unsigned)sum && o[6] == orv && o[7] == (unsigned)mx && o[8] == xr);
        EXPECT(o[9] == (unsigned)(tid >= 64 && tid < 96));
        EXPECT(o[10] == 1u && o[11] == 0u && o[12] == 52u);
        EXPECT(o[13] == (unsigned)((((tid >> 5) + 1) & 7) * 32));
    }
    std::vector<int> e(64, -1);
    int *pe = e.data();
    cuda_emu::launch(1, 64, 0, [=]() { k_early_exit(pe); });
    for (int t = 0; t < 64; t++) EXPECT(e[t] == (t < 32 ? -1 : t < 40 ? 0xff : -1) || (t < 32 && e[t] == (int)0xffffffffu));
    if (g_fail == 0) printf("emu selftest ok\n");
    return g_fail ? 1 : 0;
}
