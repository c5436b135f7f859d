// TEST INFRASTRUCTURE ONLY -- fiber engine behind tests/emu/include/cuda_runtime.h.
//
// launch() runs the blocks of a grid one after the other; inside a block every CUDA thread is a fiber
// (ucontext) on the calling OS thread.  A fiber runs until it reaches a warp collective or a block barrier,
// where it parks until every live thread of its warp (block) has arrived.  The order in which runnable fibers
// are resumed is shuffled (EMU_SEED, default 1; 0 = thread order), so code that relies on an ordering the
// hardware does not promise -- a missing barrier, a read of another warp's shared-memory slot before it was
// written -- tends to show up as a parity failure.  Mismatched collectives inside a warp (different operation
// or mask at the same rendezvous), a partial mask, or a barrier that not all live threads reach abort the
// process with a message.
#include <stdio.h>
#include <sys/mman.h>
#if !defined(__x86_64__)
#include <ucontext.h>
#endif

#include <map>
#include <vector>

#include "cuda_runtime.h"

namespace cuda_emu {

uint3 g_threadIdx, g_blockIdx;
dim3 g_blockDim, g_gridDim;

namespace {

constexpr size_t kStackBytes = 256 * 1024;
unsigned char g_smem[232448 + 256] __attribute__((aligned(128)));     // the block's dynamic shared memory + canary

struct Warp {
    int expected, arrived;
    unsigned gen;
    unsigned alive;                 // lanes that have not returned from the kernel
    int op[2];
    unsigned mask[2], present[2];
    uint64_t vals[2][32];
};

struct Block {
    int expected, arrived;
    unsigned gen;
    int acc[2];                     // or / and / count accumulators of the barrier in flight
    int mode[2];
};

// ---- context switch: callee-saved registers + stack pointer on x86-64 (swapcontext makes a system call per
// switch, which dominated the run time); ucontext elsewhere
#if defined(__x86_64__)
extern "C" void emu_switch(void **save_sp, void *new_sp);
asm(R"(
    .text
    .globl emu_switch
    .type emu_switch,@function
emu_switch:
    pushq %rbp
    pushq %rbx
    pushq %r12
    pushq %r13
    pushq %r14
    pushq %r15
    movq %rsp, (%rdi)
    movq %rsi, %rsp
    popq %r15
    popq %r14
    popq %r13
    popq %r12
    popq %rbx
    popq %rbp
    ret
    .size emu_switch,.-emu_switch
)");
struct Context { void *sp; };
#else
struct Context { ucontext_t uc; };
#endif

struct Fiber {
    Context ctx;
    char *stack;
    uint3 tid;
    int warp, lane;
    bool done, started;
    const unsigned *wait_gen;       // parked while *wait_gen == wait_val
    unsigned wait_val;
};

Context g_sched;
std::vector<Fiber> g_fibers;
std::vector<Warp> g_warps;
Block g_block;
Fiber *g_cur = nullptr;
const std::function<void()> *g_body = nullptr;
unsigned long long g_rng = 0;

[[noreturn]] void die(const char *what)
{
    fprintf(stderr, "cuda_emu: %s (block %u, thread %u)\n", what, g_blockIdx.x, g_cur ? g_cur->tid.x : 0u);
    abort();
}

void switch_to(Context *from, Context *to)
{
#if defined(__x86_64__)
    emu_switch(&from->sp, to->sp);
#else
    swapcontext(&from->uc, &to->uc);
#endif
}

void park(const unsigned *gen, unsigned val)
{
    Fiber *f = g_cur;
    f->wait_gen = gen;
    f->wait_val = val;
    switch_to(&f->ctx, &g_sched);
}

void fiber_main();

void start_fiber(Fiber &f)
{
#if defined(__x86_64__)
    // frame emu_switch pops: r15 r14 r13 r12 rbx rbp, then `ret` into fiber_main with rsp = top - 8
    void **top = reinterpret_cast<void **>(f.stack + kStackBytes);
    top[-1] = nullptr;                                   // fiber_main never returns
    top[-2] = reinterpret_cast<void *>(&fiber_main);
    for (int i = 3; i <= 8; i++) top[-i] = nullptr;
    f.ctx.sp = top - 8;
#else
    getcontext(&f.ctx.uc);
    f.ctx.uc.uc_stack.ss_sp = f.stack;
    f.ctx.uc.uc_stack.ss_size = kStackBytes;
    f.ctx.uc.uc_link = &g_sched.uc;
    makecontext(&f.ctx.uc, fiber_main, 0);
#endif
}

void fiber_main()
{
    Fiber *f = g_cur;
    (*g_body)();
    f->done = true;
    // a thread that returned no longer takes part in collectives / barriers
    Warp &w = g_warps[f->warp];
    w.alive &= ~(1u << f->lane);
    w.expected--;
    if (w.arrived > 0 && w.arrived == w.expected) {
        w.arrived = 0;
        w.gen++;
    }
    g_block.expected--;
    if (g_block.arrived > 0 && g_block.arrived == g_block.expected) {
        g_block.arrived = 0;
        g_block.gen++;
    }
    switch_to(&f->ctx, &g_sched);
    die("resumed a finished thread");
}

unsigned next_random()
{
    g_rng = g_rng * 6364136223846793005ULL + 1442695040888963407ULL;
    return (unsigned)(g_rng >> 33);
}

}  // namespace

// ---- guarded "device" memory -----------------------------------------------------------------------------------
namespace {
constexpr size_t kGuard = 256;                 // bytes of canary on each side of an allocation
constexpr unsigned char kCanary = 0xC7;
std::map<void *, size_t> g_allocs;             // user pointer -> size

void check_guards(const char *when)
{
    for (const auto &a : g_allocs) {
        const unsigned char *u = static_cast<const unsigned char *>(a.first);
        for (size_t i = 0; i < kGuard; i++)
            if (u[-(long)kGuard + (long)i] != kCanary || u[a.second + i] != kCanary) {
                fprintf(stderr, "cuda_emu: write outside a %zu-byte device buffer (%s the buffer), found %s\n", a.second,
                        u[a.second + i] != kCanary ? "after" : "before", when);
                abort();
            }
    }
}
}  // namespace

void *guarded_alloc(size_t n)
{
    void *raw = nullptr;
    if (posix_memalign(&raw, 256, n + 2 * kGuard)) return nullptr;
    unsigned char *u = static_cast<unsigned char *>(raw);
    memset(u, kCanary, kGuard);
    memset(u + kGuard, 0xA5, n);               // fresh device memory is not zero
    memset(u + kGuard + n, kCanary, kGuard);
    g_allocs[u + kGuard] = n;
    return u + kGuard;
}

void guarded_free(void *p)
{
    if (!p) return;
    check_guards("at cudaFree");
    auto it = g_allocs.find(p);
    if (it == g_allocs.end()) {
        fprintf(stderr, "cuda_emu: cudaFree of a pointer cudaMalloc did not return\n");
        abort();
    }
    g_allocs.erase(it);
    free(static_cast<unsigned char *>(p) - kGuard);
}

// A thread in a wait loop gives way to the others.  A block whose threads only ever yield is a livelock.
static unsigned long long g_yields = 0;
void yield()
{
    if (++g_yields > 200000000ULL) die("livelock: threads keep spinning without reaching a collective or returning");
    Fiber *f = g_cur;
    f->wait_gen = nullptr;
    switch_to(&f->ctx, &g_sched);
}

int lane_id() { return g_cur->lane; }
unsigned char *dynamic_smem() { return g_smem; }

uint64_t *warp_exchange(int op, unsigned mask, uint64_t v, unsigned *present)
{
    Fiber *f = g_cur;
    Warp &w = g_warps[f->warp];
    const int g = (int)(w.gen & 1u);
    if ((mask & w.alive) != w.alive || !((mask >> f->lane) & 1u))
        die("warp collective with a partial mask (the emulation only supports masks covering the live warp)");
    if (w.arrived == 0) {
        w.op[g] = op;
        w.mask[g] = mask;
        w.present[g] = 0;
    } else if (w.op[g] != op || w.mask[g] != mask) {
        die("threads of one warp are at different collectives (divergent warp-synchronous code)");
    }
    w.vals[g][f->lane] = v;
    w.present[g] |= 1u << f->lane;
    if (++w.arrived == w.expected) {
        w.arrived = 0;
        w.gen++;
    } else {
        park(&w.gen, w.gen);
    }
    *present = w.present[g];
    return w.vals[g];
}

int block_barrier(int pred, int mode)
{
    Block &b = g_block;
    const int g = (int)(b.gen & 1u);
    if (b.arrived == 0) {
        b.mode[g] = mode;
        b.acc[g] = (mode == 2) ? 1 : 0;
    } else if (b.mode[g] != mode) {
        die("threads of one block are at different barriers");
    }
    if (mode == 1) b.acc[g] |= pred ? 1 : 0;
    if (mode == 2) b.acc[g] &= pred ? 1 : 0;
    if (mode == 3) b.acc[g] += pred ? 1 : 0;
    if (++b.arrived == b.expected) {
        b.arrived = 0;
        b.gen++;
    } else {
        park(&b.gen, b.gen);
    }
    return b.acc[g];
}

void launch(dim3 grid, dim3 block, size_t smem_bytes, const std::function<void()> &body)
{
    if (g_cur) die("nested kernel launch");
    const unsigned n_threads = block.x * block.y * block.z;
    if (n_threads == 0 || n_threads > 1024) die("bad block size");
    if (smem_bytes + 256 > sizeof(g_smem)) die("dynamic shared memory exceeds the emulated limit");
    static long seed = -1;
    if (seed < 0) {
        const char *e = getenv("EMU_SEED");
        seed = e ? atol(e) : 1;
    }
    if (g_fibers.size() < n_threads) {
        const size_t old = g_fibers.size();
        g_fibers.resize(n_threads);
        for (size_t i = old; i < n_threads; i++) {
            void *p = mmap(nullptr, kStackBytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_STACK, -1, 0);
            if (p == MAP_FAILED) die("mmap(stack)");
            g_fibers[i].stack = (char *)p;
        }
    }
    const unsigned n_warps = (n_threads + 31) / 32;
    g_warps.resize(n_warps);
    std::vector<unsigned> order(n_threads);
    g_blockDim = block;
    g_gridDim = grid;
    g_body = &body;
    for (unsigned bz = 0; bz < grid.z; bz++)
        for (unsigned by = 0; by < grid.y; by++)
            for (unsigned bx = 0; bx < grid.x; bx++) {
                g_blockIdx.x = bx;
                g_blockIdx.y = by;
                g_blockIdx.z = bz;
                g_yields = 0;
                memset(g_smem, 0xA5, smem_bytes);      // uninitialised shared memory is not zero
                memset(g_smem + smem_bytes, kCanary, kGuard);
                for (unsigned w = 0; w < n_warps; w++) {
                    Warp &W = g_warps[w];
                    memset(&W, 0, sizeof(W));
                    const unsigned n = (w + 1) * 32 <= n_threads ? 32u : n_threads - w * 32;
                    W.expected = (int)n;
                    W.alive = n == 32 ? 0xffffffffu : ((1u << n) - 1u);
                }
                memset(&g_block, 0, sizeof(g_block));
                g_block.expected = (int)n_threads;
                for (unsigned t = 0; t < n_threads; t++) {
                    Fiber &f = g_fibers[t];
                    f.tid.x = t % block.x;
                    f.tid.y = (t / block.x) % block.y;
                    f.tid.z = t / (block.x * block.y);
                    f.warp = (int)(t / 32);
                    f.lane = (int)(t % 32);
                    f.done = false;
                    f.started = false;
                    f.wait_gen = nullptr;
                    order[t] = t;
                }
                if (seed) {
                    g_rng = (unsigned long long)seed * 0x9E3779B97F4A7C15ULL + bx * 1315423911ULL;
                    for (unsigned t = n_threads - 1; t > 0; t--) {
                        const unsigned j = next_random() % (t + 1);
                        const unsigned tmp = order[t];
                        order[t] = order[j];
                        order[j] = tmp;
                    }
                }
                unsigned remaining = n_threads;
                while (remaining) {
                    bool progressed = false;
                    for (unsigned q = 0; q < n_threads; q++) {
                        Fiber &f = g_fibers[order[q]];
                        if (f.done) continue;
                        if (f.wait_gen && *f.wait_gen == f.wait_val) continue;
                        f.wait_gen = nullptr;
                        g_cur = &f;
                        g_threadIdx = f.tid;
                        if (!f.started) {
                            f.started = true;
                            start_fiber(f);
                        }
                        switch_to(&g_sched, &f.ctx);
                        progressed = true;
                        if (f.done) remaining--;
                    }
                    if (!progressed) die("deadlock: live threads wait at a collective / barrier the others never reach");
                }
                g_cur = nullptr;
                for (size_t i = 0; i < kGuard; i++)
                    if (g_smem[smem_bytes + i] != kCanary) die("write beyond the block's dynamic shared memory");
            }
    g_body = nullptr;
    check_guards("after a kernel");
}

}  // namespace cuda_emu

// Marks this shared library as the host emulation: gr_adsb_b200/_native.py refuses to load a library exporting
// this symbol unless the test switch ADSB_EMU=1 is set, so the package cannot end up computing on the CPU by accident.
extern "C" int adsb_host_emulation_marker(void) { return 1; }
