// cuda_runtime.h -- CPU EMULATION of the slice of CUDA that csrc/*.cu uses.  TEST INFRASTRUCTURE ONLY.
//
// tests/cuda_emu/build_emu.py rewrites the kernel sources mechanically (launch syntax, __shared__, the inline-PTX
// helpers) and compiles them with g++ against this header into tests/cuda_emu/_build/libfs2d_emu.so, so that the
// *kernel source code itself* -- tile indexing, clamp repair, slow-cell lists, barrier placement, TMA box geometry --
// can be executed and compared with the oracle on a machine without a GPU.  It is never loaded by the product
// (fs/_lib.py only ever opens lib/libfs2d.so) and proves nothing about performance or about hardware memory ordering.
//
// Execution model: one OS thread.  A launch runs its CTAs one after the other; the threads of a CTA are ucontext
// fibers scheduled round-robin; a fiber runs until it blocks in __syncthreads / a warp collective / an mbarrier wait.
// TMA box loads complete synchronously at issue time (the earliest moment the hardware could deliver them), so a
// load issued into a buffer that other threads have not finished reading corrupts the result and is detected.
#pragma once
#include <ucontext.h>

#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <functional>
#include <vector>

// ---- qualifiers ----------------------------------------------------------------------------------------------------
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __restrict__
#define __launch_bounds__(...)
#define __grid_constant__
#define __align__(n) __attribute__((aligned(n)))

// ---- vector types --------------------------------------------------------------------------------------------------
struct alignas(8) float2 { float x, y; };
struct alignas(16) float4 { float x, y, z, w; };
struct uchar4 { unsigned char x, y, z, w; };
struct uint3 { unsigned x, y, z; };
struct alignas(16) uint4 { unsigned x, y, z, w; };
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
inline float2 make_float2(float x, float y) { return float2{x, y}; }
inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }

// ---- host API stubs ------------------------------------------------------------------------------------------------
typedef int cudaError_t;
constexpr cudaError_t cudaSuccess = 0;
typedef void *cudaStream_t;
struct cudaDeviceProp { int major = 10, minor = 0; };
enum { cudaDevAttrMultiProcessorCount = 16 };
enum { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
enum { cudaEnableDefault = 0 };
enum cudaDriverEntryPointQueryResult { cudaDriverEntryPointSuccess = 0 };
inline const char *cudaGetErrorString(cudaError_t) { return "emulated"; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline cudaError_t cudaGetDevice(int *d) { *d = 0; return cudaSuccess; }
inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int) { *p = cudaDeviceProp(); return cudaSuccess; }
inline cudaError_t cudaDeviceGetAttribute(int *v, int, int) {
    const char *e = getenv("FS2D_EMU_SMS");      // "SM count": how many persistent CTAs a launch gets
    *v = e ? atoi(e) : 3;
    return cudaSuccess;
}
template <class F>
inline cudaError_t cudaFuncSetAttribute(F, int, int) { return cudaSuccess; }
template <class T>
inline cudaError_t cudaMalloc(T **p, size_t n) { *p = (T *)calloc(1, n); return cudaSuccess; }
inline cudaError_t cudaMemsetAsync(void *p, int v, size_t n, cudaStream_t) { memset(p, v, n); return cudaSuccess; }
inline cudaError_t cudaFree(void *p) { free(p); return cudaSuccess; }
enum cudaMemcpyKind { cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2 };
inline cudaError_t cudaMemcpyAsync(void *dst, const void *src, size_t n, cudaMemcpyKind, cudaStream_t) { memcpy(dst, src, n); return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaGetDriverEntryPoint(const char *name, void **fn, int, cudaDriverEntryPointQueryResult *q);   // cuda.h

// ---- the fiber machine ---------------------------------------------------------------------------------------------
namespace emu {

struct Warp {
    int active = 0, arrived = 0, gen = 0;
    uint32_t buf[2][32];
    int pred_and[2] = {1, 1};
};
struct Fiber {
    ucontext_t ctx;
    char *stack = nullptr;
    uint3 tid;
    int lin = 0;
    bool done = false;
};
struct Cta {
    int n = 0, exited = 0, arrived = 0, gen = 0;
    int acc_and[2] = {1, 1}, acc_or[2] = {0, 0}, res_and[2] = {1, 1}, res_or[2] = {0, 0};
    std::vector<Warp> warps;
    void *dyn = nullptr;
    uint3 bid;
    dim3 bdim, gdim;
};
struct Machine {
    ucontext_t sched;
    std::vector<Fiber> fibers;
    Fiber *cur = nullptr;
    Cta cta;
    std::function<void()> body;
    unsigned long progress = 0;
};
inline Machine &M() {
    static Machine m;
    return m;
}
constexpr size_t STACK = 192 * 1024;

inline void yield() { swapcontext(&M().cur->ctx, &M().sched); }
inline void cta_release() {
    Cta &c = M().cta;
    const int g = c.gen & 1;
    c.res_and[g] = c.acc_and[g];
    c.res_or[g] = c.acc_or[g];
    c.acc_and[g ^ 1] = 1;
    c.acc_or[g ^ 1] = 0;
    c.arrived = 0;
    ++c.gen;
    ++M().progress;
}
// __syncthreads with an optional predicate reduction; exited threads count as arrived
inline void barrier(int pred, int *r_and, int *r_or) {
    Cta &c = M().cta;
    const int g = c.gen & 1, my = c.gen;
    c.acc_and[g] &= pred != 0;
    c.acc_or[g] |= pred != 0;
    ++M().progress;
    if (++c.arrived + c.exited == c.n) cta_release();
    while (c.gen == my) yield();
    if (r_and) *r_and = c.res_and[g];
    if (r_or) *r_or = c.res_or[g];
}
// bar.sync id, count: a named barrier shared by exactly `count` threads (ids 1..15)
struct NamedBar { int arrived = 0, gen = 0; };
inline NamedBar *named_bars() {
    static NamedBar nb[16];
    return nb;
}
inline void named_barrier(int id, int count) {
    if (id < 1 || id > 15 || count % 32 != 0) { fprintf(stderr, "cuda_emu: bad named barrier (%d, %d)\n", id, count); abort(); }
    NamedBar &b = named_bars()[id];
    const int my = b.gen;
    ++M().progress;
    if (++b.arrived == count) {
        b.arrived = 0;
        ++b.gen;
    }
    while (b.gen == my) yield();
}
inline Warp &my_warp() { return M().cta.warps[M().cur->lin / 32]; }
inline void warp_release(Warp &w) {
    w.arrived = 0;
    ++w.gen;
    ++M().progress;
}
// all active lanes deposit `v`; returns the generation slot to read from
inline int warp_rendezvous(uint32_t v, int pred) {
    Warp &w = my_warp();
    const int g = w.gen & 1, my = w.gen, lane = M().cur->lin % 32;
    if (w.arrived == 0) w.pred_and[g] = 1;
    w.buf[g][lane] = v;
    w.pred_and[g] &= pred != 0;
    ++M().progress;
    if (++w.arrived == w.active) warp_release(w);
    while (w.gen == my) yield();
    return g;
}
inline void fiber_exit() {
    Machine &m = M();
    Cta &c = m.cta;
    m.cur->done = true;
    ++c.exited;
    ++m.progress;
    if (c.arrived > 0 && c.arrived + c.exited == c.n) cta_release();
    Warp &w = my_warp();
    --w.active;
    if (w.arrived > 0 && w.arrived == w.active) warp_release(w);
}
inline void trampoline() {
    M().body();
    fiber_exit();
    swapcontext(&M().cur->ctx, &M().sched);
}
inline void *dyn_smem() { return M().cta.dyn; }

template <class F>
void launch(dim3 grid, dim3 block, size_t smem, F f) {
    Machine &m = M();
    const int n = (int)(block.x * block.y * block.z);
    if ((int)m.fibers.size() < n) {
        const size_t old = m.fibers.size();
        m.fibers.resize(n);
        for (size_t i = old; i < (size_t)n; ++i) m.fibers[i].stack = (char *)malloc(STACK);
    }
    void *dyn = nullptr;
    if (posix_memalign(&dyn, 1024, smem ? smem : 1024)) abort();
    m.body = f;
    for (unsigned bz = 0; bz < grid.z; ++bz)
        for (unsigned by = 0; by < grid.y; ++by)
            for (unsigned bx = 0; bx < grid.x; ++bx) {
                Cta &c = m.cta;
                c = Cta();
                c.n = n;
                c.dyn = dyn;
                c.bid = uint3{bx, by, bz};
                c.bdim = block;
                c.gdim = grid;
                c.warps.assign((n + 31) / 32, Warp());
                for (int q = 0; q < 16; ++q) named_bars()[q] = NamedBar();
                memset(dyn, 0xCD, smem ? smem : 1024);     // shared memory starts as garbage
                for (int t = 0; t < n; ++t) {
                    Fiber &fb = m.fibers[t];
                    fb.done = false;
                    fb.lin = t;
                    fb.tid = uint3{t % block.x, (t / block.x) % block.y, t / (block.x * block.y)};
                    ++c.warps[t / 32].active;
                    getcontext(&fb.ctx);
                    fb.ctx.uc_stack.ss_sp = fb.stack;
                    fb.ctx.uc_stack.ss_size = STACK;
                    fb.ctx.uc_link = nullptr;
                    makecontext(&fb.ctx, (void (*)())trampoline, 0);
                }
                // Schedule.  Default: round-robin over all fibers (deterministic, warps advance in lock step).
                // FS2D_EMU_SCHED=<seed>: adversarial starvation -- one randomly chosen "victim" warp is not run at all while
                // any other warp can still make progress, then it runs one slice and a new victim is drawn.  Warps thereby
                // drift apart as far as the kernel's own synchronisation allows, so a missing barrier shows up as a wrong
                // result or a deadlock (tests/test_kernels_emulated.py runs the synchronisation-heavy kernels under several seeds).
                static const char *sched_env = getenv("FS2D_EMU_SCHED");
                static unsigned long long rng = sched_env ? 0x9E3779B97F4A7C15ull * (unsigned long long)(atoll(sched_env) + 1) : 0;
                auto next_rand = [&]() { rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17; return rng; };
                const int n_warps = (n + 31) / 32;
                std::vector<int> order(n_warps);
                for (int q = 0; q < n_warps; ++q) order[q] = q;
                auto run_warp = [&](int wq) {
                    for (int t = wq * 32; t < n && t < wq * 32 + 32; ++t) {
                        Fiber &fb = m.fibers[t];
                        if (fb.done) continue;
                        m.cur = &fb;
                        swapcontext(&m.sched, &fb.ctx);
                    }
                };
                // the leader warp (tile scheduler, TMA issue) is the most interesting victim: starve it half of the time
                auto draw_victim = [&]() { return (next_rand() & 1) ? 0 : (int)(next_rand() % n_warps); };
                int live = n, stalled = 0, victim = sched_env ? draw_victim() : -1;
                while (live > 0) {
                    const unsigned long before = m.progress;
                    if (sched_env)
                        for (int q = n_warps - 1; q > 0; --q) std::swap(order[q], order[next_rand() % (q + 1)]);
                    for (int q = 0; q < n_warps; ++q)
                        if (order[q] != victim) run_warp(order[q]);
                    if (sched_env && m.progress == before) {     // everybody else is blocked: let the victim move, redraw
                        run_warp(victim);
                        victim = draw_victim();
                    }
                    live = 0;
                    for (int t = 0; t < n; ++t) live += !m.fibers[t].done;
                    stalled = m.progress == before ? stalled + 1 : 0;
                    if (live > 0 && stalled > (sched_env ? 2 * n_warps : 0)) {
                        fprintf(stderr, "cuda_emu: deadlock in block (%u,%u,%u): %d threads blocked (barrier %d/%d arrived)\n", bx, by,
                                bz, live, c.arrived, c.n - c.exited);
                        abort();
                    }
                }
            }
    free(dyn);
}

// ---- mbarrier + TMA ----------------------------------------------------------------------------------------------
struct MBar {
    uint32_t tx;
    uint16_t pending;
    uint8_t phase, count;
};
static_assert(sizeof(MBar) == 8, "an mbarrier is a 64-bit shared-memory object");
inline void mbar_check(MBar *b) {
    if (b->pending == 0 && b->tx == 0) {
        b->phase ^= 1;
        b->pending = b->count;
        ++M().progress;
    }
}
inline void mbar_init(uint64_t *bar, uint32_t count) {
    MBar *b = reinterpret_cast<MBar *>(bar);
    b->tx = 0;
    b->pending = (uint16_t)count;
    b->phase = 0;
    b->count = (uint8_t)count;
}
inline void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {   // mbarrier.arrive.expect_tx
    MBar *b = reinterpret_cast<MBar *>(bar);
    b->tx += bytes;
    --b->pending;
    mbar_check(b);
}
inline void mbar_wait(uint64_t *bar, uint32_t parity) {       // try_wait.parity loop
    MBar *b = reinterpret_cast<MBar *>(bar);
    while (b->phase == (uint8_t)parity) yield();
}
// progress flags in shared memory (st.release / ld.acquire at CTA scope): a store is progress, a poll is a scheduling point --
// a polling loop that never sees its value is reported as a deadlock like any other blocked fiber
inline void flag_store(int *p, int v) {
    *p = v;
    ++M().progress;
}
inline int flag_load(const int *p) {
    yield();
    return *p;
}
// the polling loop itself: blocked while the counter is below `v`; getting past it is progress (a poll that finds its value
// at once is not a scheduling point, otherwise a round in which every fiber merely walks through satisfied polls would look
// like a deadlock)
inline void flag_wait_ge(const int *p, int v) {
    while (*p < v) yield();
    ++M().progress;
}
}  // namespace emu

// ---- device builtins -----------------------------------------------------------------------------------------------
#define threadIdx (emu::M().cur->tid)
#define blockIdx (emu::M().cta.bid)
#define blockDim (emu::M().cta.bdim)
#define gridDim (emu::M().cta.gdim)

template <class T>
inline T __ldg(const T *p) { return *p; }
inline int min(int a, int b) { return a < b ? a : b; }
inline int max(int a, int b) { return a > b ? a : b; }
inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }
inline float __int_as_float(int i) { float f; memcpy(&f, &i, 4); return f; }
inline int __float_as_int(float f) { int i; memcpy(&i, &f, 4); return i; }
inline void __syncthreads() { emu::barrier(1, nullptr, nullptr); }
inline int __syncthreads_and(int p) { int r; emu::barrier(p, &r, nullptr); return r; }
inline int __syncthreads_or(int p) { int r; emu::barrier(p, nullptr, &r); return r; }
inline float __shfl_up_sync(unsigned, float v, int delta) {
    const int lane = emu::M().cur->lin % 32;
    const int g = emu::warp_rendezvous((uint32_t)__float_as_int(v), 1);
    return lane - delta >= 0 ? __int_as_float((int)emu::my_warp().buf[g][lane - delta]) : v;
}
inline float __shfl_down_sync(unsigned, float v, int delta) {
    const int lane = emu::M().cur->lin % 32;
    const int g = emu::warp_rendezvous((uint32_t)__float_as_int(v), 1);
    return lane + delta < 32 ? __int_as_float((int)emu::my_warp().buf[g][lane + delta]) : v;
}
inline void __syncwarp(unsigned = 0xffffffffu) { emu::warp_rendezvous(0, 1); }
inline int __shfl_up_sync(unsigned, int v, int delta) {
    const int lane = emu::M().cur->lin % 32;
    const int g = emu::warp_rendezvous((uint32_t)v, 1);
    return lane - delta >= 0 ? (int)emu::my_warp().buf[g][lane - delta] : v;
}
inline int __shfl_sync(unsigned, int v, int src) {
    const int g = emu::warp_rendezvous((uint32_t)v, 1);
    return (int)emu::my_warp().buf[g][src & 31];
}
inline int __popc(unsigned x) { return __builtin_popcount(x); }
inline int __all_sync(unsigned, int pred) {
    const int g = emu::warp_rendezvous(0, pred);
    return emu::my_warp().pred_and[g];
}
inline unsigned __float_as_uint(float f) { unsigned i; memcpy(&i, &f, 4); return i; }
inline float __uint_as_float(unsigned i) { float f; memcpy(&f, &i, 4); return f; }
inline float __shfl_xor_sync(unsigned, float v, int lane_mask) {
    const int lane = emu::M().cur->lin % 32;
    const int g = emu::warp_rendezvous((uint32_t)__float_as_int(v), 1);
    return __int_as_float((int)emu::my_warp().buf[g][lane ^ lane_mask]);
}
inline unsigned atomicMax(unsigned *p, unsigned v) { const unsigned o = *p; if (v > o) *p = v; return o; }
inline int atomicAdd(int *p, int v) { const int o = *p; *p += v; return o; }
inline unsigned atomicAdd(unsigned *p, unsigned v) { const unsigned o = *p; *p += v; return o; }
