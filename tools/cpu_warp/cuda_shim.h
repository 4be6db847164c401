// cuda_shim.h -- just enough of the CUDA execution model to run ONE WARP of the compress kernels on the CPU.
//
// TEST INFRASTRUCTURE ONLY (tools/cpu_warp/run_window_kernel.cpp, tests/test_kernel_on_cpu_warp.py).  The kernel
// sources in snappy.jl_b200/csrc are compiled unchanged with -DSB200_CPU_EMU: their few inline-PTX statements have
// a plain C++ twin under that macro, everything else goes through the definitions below.
//
// Model: the 32 lanes of a warp are 32 coroutines (ucontext) on one OS thread.  A lane runs until it reaches a
// warp collective (__shfl_sync, __ballot_sync, __match_any_sync, __syncwarp, ...), parks there, and the scheduler
// resumes the next lane; when the 32nd lane arrives the collective completes for all of them.  That is the
// convergence contract the kernels are written against (full-mask collectives are reached by every lane), so a
// kernel that is correct here makes the same decisions lane for lane as on the GPU; what the model does not show is
// timing and memory-ordering races between lanes that are not separated by a collective.
#pragma once
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <ucontext.h>

#define __device__
#define __global__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __restrict__
#define __shared__
#define __align__(x) __attribute__((aligned(x)))
#define __launch_bounds__(...)

struct uint4 {
    uint32_t x, y, z, w;
};
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{x, y, z, w}; }
struct dim3_ {
    uint32_t x, y, z;
};

namespace cpu_warp {

constexpr int kLanes = 32;
constexpr size_t kStack = 256 * 1024;

struct Warp {
    ucontext_t sched, lane_ctx[kLanes];
    char* stacks[kLanes];
    bool done[kLanes];
    int cur = 0;
    uint32_t tid_base = 0, block = 0, block_dim = 32, grid_dim = 1;  // which warp of which CTA this is
    // collective state
    uint32_t vals[kLanes], snap[kLanes];
    int arrived = 0;
    uint64_t gen = 0, collectives = 0;
    void (*entry)(void*) = nullptr;
    void* arg = nullptr;
};
inline Warp& W() {
    static Warp w;
    return w;
}

inline void yield_lane() {
    Warp& w = W();
    swapcontext(&w.lane_ctx[w.cur], &w.sched);
}

// every lane contributes v; returns once all 32 have, with everyone's values in out[]
inline void exchange(uint32_t v, uint32_t out[kLanes]) {
    Warp& w = W();
    const int me = w.cur;
    w.vals[me] = v;
    const uint64_t g = w.gen;
    if (++w.arrived == kLanes) {
        memcpy(w.snap, w.vals, sizeof w.snap);
        w.arrived = 0;
        w.gen++;
        w.collectives++;
    } else {
        int spins = 0;
        while (w.gen == g) {
            yield_lane();
            if (++spins > 100000) {
                fprintf(stderr, "cpu_warp: lane %d waits at a collective the other lanes never reach\n", me);
                abort();
            }
        }
    }
    memcpy(out, w.snap, sizeof w.snap);
}

inline void trampoline() {
    Warp& w = W();
    w.entry(w.arg);
    w.done[w.cur] = true;
    for (;;) yield_lane();
}

// run entry(arg) once per lane, as one warp
inline void run_warp(void (*entry)(void*), void* arg) {
    Warp& w = W();
    w.entry = entry;
    w.arg = arg;
    w.arrived = 0;
    for (int l = 0; l < kLanes; l++) {
        if (!w.stacks[l]) w.stacks[l] = (char*)malloc(kStack);
        getcontext(&w.lane_ctx[l]);
        w.lane_ctx[l].uc_stack.ss_sp = w.stacks[l];
        w.lane_ctx[l].uc_stack.ss_size = kStack;
        w.lane_ctx[l].uc_link = &w.sched;
        makecontext(&w.lane_ctx[l], (void (*)())trampoline, 0);
        w.done[l] = false;
    }
    for (;;) {
        bool any = false;
        for (int l = 0; l < kLanes; l++) {
            if (w.done[l]) continue;
            any = true;
            w.cur = l;
            swapcontext(&w.sched, &w.lane_ctx[l]);
        }
        if (!any) break;
    }
    if (w.arrived != 0) {
        fprintf(stderr, "cpu_warp: %d lane(s) left parked at a collective\n", w.arrived);
        abort();
    }
}

struct Idx {
    uint32_t x, y, z;
};
inline Idx thread_idx() { return Idx{W().tid_base + (uint32_t)W().cur, 0, 0}; }
inline Idx block_idx() { return Idx{W().block, 0, 0}; }
inline Idx block_dim() { return Idx{W().block_dim, 1, 1}; }
inline Idx grid_dim() { return Idx{W().grid_dim, 1, 1}; }

// kernels whose threads never meet at a collective (one thread per item): every thread of every CTA, one after the other
template <typename K, typename... A>
inline void launch_independent_threads(uint32_t grid, uint32_t block, K kernel, A... args) {
    Warp& w = W();
    w.block_dim = block;
    w.grid_dim = grid;
    w.cur = 0;
    for (uint32_t b = 0; b < grid; b++)
        for (uint32_t t = 0; t < block; t++) {
            w.block = b;
            w.tid_base = t;
            kernel(args...);
        }
    w.block = w.tid_base = 0;
    w.block_dim = 32;
    w.grid_dim = 1;
}

}  // namespace cpu_warp

#define threadIdx (cpu_warp::thread_idx())
#define blockIdx (cpu_warp::block_idx())
#define blockDim (cpu_warp::block_dim())
#define gridDim (cpu_warp::grid_dim())

// ---- warp collectives (full mask only: the kernels never use another) -----------------------------------------
static inline void check_full(uint32_t mask) {
    if (mask != 0xffffffffu) {
        fprintf(stderr, "cpu_warp: partial-mask collective\n");
        abort();
    }
}
static inline uint32_t __shfl_sync(uint32_t mask, uint32_t v, uint32_t src) {
    check_full(mask);
    uint32_t all[32];
    cpu_warp::exchange(v, all);
    return all[src & 31];
}
static inline uint32_t __shfl_up_sync(uint32_t mask, uint32_t v, uint32_t delta) {
    check_full(mask);
    uint32_t all[32];
    cpu_warp::exchange(v, all);
    const uint32_t me = (uint32_t)cpu_warp::W().cur;
    return me >= delta ? all[me - delta] : v;
}
static inline uint32_t __ballot_sync(uint32_t mask, int pred) {
    check_full(mask);
    uint32_t all[32], r = 0;
    cpu_warp::exchange(pred ? 1u : 0u, all);
    for (int l = 0; l < 32; l++) r |= all[l] << l;
    return r;
}
static inline int __any_sync(uint32_t mask, int pred) { return __ballot_sync(mask, pred) != 0; }
static inline uint32_t __reduce_or_sync(uint32_t mask, uint32_t v) {
    check_full(mask);
    uint32_t all[32], r = 0;
    cpu_warp::exchange(v, all);
    for (int l = 0; l < 32; l++) r |= all[l];
    return r;
}
static inline uint32_t __match_any_sync(uint32_t mask, uint32_t v) {
    check_full(mask);
    uint32_t all[32], r = 0;
    cpu_warp::exchange(v, all);
    for (int l = 0; l < 32; l++) r |= (uint32_t)(all[l] == v) << l;
    return r;
}
static inline void __syncwarp(uint32_t mask = 0xffffffffu) {
    check_full(mask);
    uint32_t all[32];
    cpu_warp::exchange(0, all);
}

// ---- scalar intrinsics -------------------------------------------------------------------------------------------
static inline uint32_t __funnelshift_r(uint32_t lo, uint32_t hi, uint32_t sh) {
    return (uint32_t)((((uint64_t)hi << 32) | lo) >> (sh & 31u));
}
static inline uint32_t __funnelshift_rc(uint32_t lo, uint32_t hi, uint32_t sh) {
    return (uint32_t)((((uint64_t)hi << 32) | lo) >> (sh > 32u ? 32u : sh));
}
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline int __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }
template <typename T>
static inline T __ldg(const T* p) { return *p; }
template <typename T>
static inline T __ldcg(const T* p) { return *p; }
static inline void __syncthreads() {  // kernels with several cooperating warps per CTA are not modelled
    fprintf(stderr, "cpu_warp: __syncthreads() reached: this kernel needs more than one warp\n");
    abort();
}
static inline void __nanosleep(unsigned) { cpu_warp::yield_lane(); }
static inline void __threadfence() {}
static inline uint32_t atomicAdd(uint32_t* p, uint32_t v) {
    const uint32_t old = *p;
    *p = old + v;
    return old;
}
static inline uint32_t __reduce_max_sync(uint32_t mask, uint32_t v) {
    check_full(mask);
    uint32_t all[32], r = 0;
    cpu_warp::exchange(v, all);
    for (int l = 0; l < 32; l++) r = all[l] > r ? all[l] : r;
    return r;
}
static inline void __threadfence_system() {}
static inline uint32_t atomicMax(uint32_t* p, uint32_t v) {
    const uint32_t old = *p;
    if (v > old) *p = v;
    return old;
}
template <typename T>
static inline T atomicMin(T* p, T v) {
    const T old = *p;
    if (v < old) *p = v;
    return old;
}
static inline uint32_t atomicOr(uint32_t* p, uint32_t v) {
    const uint32_t old = *p;
    *p = old | v;
    return old;
}
template <typename T>
static inline T min(T a, T b) { return a < b ? a : b; }
template <typename T>
static inline T max(T a, T b) { return a > b ? a : b; }
