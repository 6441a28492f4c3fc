// tests/hostsim/cuda_emu.h -- TEST HARNESS, NOT PRODUCT CODE.
//
// A small CUDA execution emulator for the build container (which has no GPU): the kernels of
// pantas_b200/csrc/*.cuh are compiled with g++ and every CUDA thread of a block runs as a fibre
// (ucontext) on one OS thread.  __syncthreads() and the warp shuffles are yield points with the
// real semantics (all live threads of the block / all 32 lanes of the warp must arrive), atomics are
// plain read-modify-writes (fibres only switch at yield points), blocks run one after another.
// Shared-memory races and memory-model bugs are NOT found by this; logic errors are.
#pragma once
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <ucontext.h>

#include <functional>
#include <type_traits>
#include <vector>

#define PT_EMU 1
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__ inline
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))
#define __shared__ static
#define __restrict__

struct uint4 { uint32_t x, y, z, w; };
struct uint2 { uint32_t x, y; };
inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{x, y, z, w}; }
struct emu_dim3 { unsigned x, y, z; };

namespace emu {

constexpr size_t STACK_BYTES = 256 * 1024;

struct Warp {
    uint32_t val[32];
    uint32_t snap[2][32];
    uint32_t arrived = 0, gen = 0;
};
struct Fibre {
    ucontext_t ctx;
    char* stack = nullptr;
    bool done = false;
};
struct State {
    emu_dim3 tidx{0, 0, 0}, bidx{0, 0, 0}, bdim{1, 1, 1}, gdim{1, 1, 1};
    uint8_t* dyn = nullptr;
    ucontext_t sched;
    std::vector<Fibre> fib;
    std::vector<Warp> warps;
    uint32_t live = 0, bar_arrived = 0, bar_gen = 0;
    uint32_t named_arrived[16] = {0}, named_gen[16] = {0};
    uint64_t progress = 0;
    std::vector<int> where;        // what every fibre is waiting in (1 block barrier, 2 shuffle, 100 + id named barrier): deadlock report
    int cur = -1;
    const std::function<void()>* body = nullptr;
};
inline State& S() {
    static State s;
    return s;
}

inline void yield() {
    State& s = S();
    swapcontext(&s.fib[s.cur].ctx, &s.sched);
}
inline void trampoline() {
    State& s = S();
    (*s.body)();
    s.fib[s.cur].done = true;
    s.live--;
    swapcontext(&s.fib[s.cur].ctx, &s.sched);
}

// run `body` once per thread of a grid x block launch (1-D), with `smem_bytes` of dynamic shared memory
inline void launch(unsigned grid, unsigned block, size_t smem_bytes, const std::function<void()>& body) {
    State& s = S();
    std::vector<uint8_t> dyn(smem_bytes + 256);
    s.dyn = (uint8_t*)(((uintptr_t)dyn.data() + 127) & ~(uintptr_t)127);
    s.gdim = {grid, 1, 1};
    s.bdim = {block, 1, 1};
    s.body = &body;
    if (s.fib.size() < block) s.fib.resize(block);
    s.where.assign(block, 0);
    for (unsigned t = 0; t < block; t++)
        if (!s.fib[t].stack) s.fib[t].stack = (char*)malloc(STACK_BYTES);
    for (unsigned b = 0; b < grid; b++) {
        s.bidx = {b, 0, 0};
        memset(s.dyn, 0xAA, smem_bytes);      // a fresh CTA finds arbitrary bytes in shared memory
        s.warps.assign((block + 31) / 32, Warp());
        s.live = block;
        s.bar_arrived = 0;
        for (int k = 0; k < 16; k++) { s.named_arrived[k] = 0; s.named_gen[k] = 0; }
        for (unsigned t = 0; t < block; t++) {
            Fibre& f = s.fib[t];
            f.done = false;
            getcontext(&f.ctx);
            f.ctx.uc_stack.ss_sp = f.stack;
            f.ctx.uc_stack.ss_size = STACK_BYTES;
            f.ctx.uc_link = &s.sched;
            makecontext(&f.ctx, (void (*)())trampoline, 0);
        }
        uint64_t idle_passes = 0;
        while (s.live > 0) {
            const uint32_t live_before = s.live, gen_before = s.bar_gen;
            const uint64_t prog_before = s.progress;
            for (unsigned t = 0; t < block; t++) {
                if (s.fib[t].done) continue;
                s.cur = (int)t;
                s.tidx = {t, 0, 0};
                swapcontext(&s.sched, &s.fib[t].ctx);
            }
            if (s.live == live_before && s.bar_gen == gen_before && s.progress == prog_before) {
                if (++idle_passes > 1000000) {
                    fprintf(stderr, "cuda_emu: deadlock (barrier / shuffle that not every thread reaches)\n");
                    for (unsigned t = 0; t < block; t++)
                        if (!s.fib[t].done) fprintf(stderr, "  thread %u waits in %d\n", t, s.where[t]);
                    abort();
                }
            } else {
                idle_passes = 0;
            }
        }
    }
    s.dyn = nullptr;
}

inline uint32_t shfl(uint32_t v, int src_lane) {
    State& s = S();
    Warp& w = s.warps[s.cur >> 5];
    const uint32_t lane = (uint32_t)s.cur & 31u;
    const uint32_t width = s.bdim.x - ((uint32_t)s.cur & ~31u) < 32u ? s.bdim.x - ((uint32_t)s.cur & ~31u) : 32u;
    s.where[s.cur] = 2;
    w.val[lane] = v;
    w.arrived++;
    const uint32_t gen = w.gen;
    while (w.gen == gen) {
        if (w.arrived == width) {
            memcpy(w.snap[gen & 1u], w.val, sizeof w.val);
            w.arrived = 0;
            w.gen++;
            break;
        }
        yield();
    }
    return (src_lane < 0 || src_lane >= (int)width) ? v : w.snap[gen & 1u][src_lane];
}

}  // namespace emu

#define threadIdx (emu::S().tidx)
#define blockIdx (emu::S().bidx)
#define blockDim (emu::S().bdim)
#define gridDim (emu::S().gdim)

inline void __syncthreads() {
    emu::State& s = emu::S();
    s.where[s.cur] = 1;
    s.bar_arrived++;
    const uint32_t gen = s.bar_gen;
    while (s.bar_gen == gen) {
        if (s.bar_arrived >= s.live) {
            s.bar_arrived = 0;
            s.bar_gen++;
            break;
        }
        emu::yield();
    }
}
// bar.sync id, count: `count` threads of the block meet at named barrier `id` (1..15)
inline void emu_named_barrier(unsigned id, unsigned count) {
    emu::State& s = emu::S();
    s.where[s.cur] = 100 + (int)id;
    s.named_arrived[id]++;
    const uint32_t gen = s.named_gen[id];
    while (s.named_gen[id] == gen) {
        if (s.named_arrived[id] >= count) {
            s.named_arrived[id] = 0;
            s.named_gen[id]++;
            s.progress++;                    // (for the deadlock detector)
            break;
        }
        emu::yield();
    }
}
inline void __syncwarp(unsigned = 0xffffffffu) { (void)emu::shfl(0, 0); }

inline uint32_t __shfl_sync(unsigned, uint32_t v, int src) { return emu::shfl(v, src & 31); }
inline uint32_t __shfl_up_sync(unsigned, uint32_t v, unsigned d) { return emu::shfl(v, (int)(emu::S().cur & 31) - (int)d); }
inline uint32_t __shfl_down_sync(unsigned, uint32_t v, unsigned d) { return emu::shfl(v, (int)(emu::S().cur & 31) + (int)d); }
inline uint32_t __shfl_xor_sync(unsigned, uint32_t v, int m) { return emu::shfl(v, (int)(emu::S().cur & 31) ^ m); }
inline unsigned long long __shfl_up_sync(unsigned m, unsigned long long v, unsigned d) {
    const uint32_t lo = __shfl_up_sync(m, (uint32_t)v, d), hi = __shfl_up_sync(m, (uint32_t)(v >> 32), d);
    return ((unsigned long long)hi << 32) | lo;
}
inline unsigned long long __shfl_sync(unsigned m, unsigned long long v, int src) {
    const uint32_t lo = __shfl_sync(m, (uint32_t)v, src), hi = __shfl_sync(m, (uint32_t)(v >> 32), src);
    return ((unsigned long long)hi << 32) | lo;
}
inline int __shfl_xor_sync(unsigned m, int v, int x) { return (int)__shfl_xor_sync(m, (uint32_t)v, x); }
inline unsigned __ballot_sync(unsigned, int pred) {
    unsigned r = 0;
    for (int l = 0; l < 32; l++) r |= (emu::shfl(pred ? 1u : 0u, l) & 1u) << l;
    return r;
}

inline int __any_sync(unsigned m, int pred) { return __ballot_sync(m, pred) != 0u; }

// ---- intrinsics
inline long long clock64() { return 0; }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __popcll(unsigned long long v) { return __builtin_popcountll(v); }
inline int __ffs(int v) { return __builtin_ffs(v); }
inline int __ffsll(long long v) { return __builtin_ffsll(v); }
inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
inline int __clzll(long long v) { return v ? __builtin_clzll((unsigned long long)v) : 64; }
inline uint32_t __funnelshift_l(uint32_t lo, uint32_t hi, uint32_t sh) {
    const uint64_t v = ((uint64_t)hi << 32) | lo;
    return (uint32_t)((v << (sh & 31u)) >> 32);
}
inline uint32_t __funnelshift_r(uint32_t lo, uint32_t hi, uint32_t sh) {
    const uint64_t v = ((uint64_t)hi << 32) | lo;
    return (uint32_t)(v >> (sh & 31u));
}
inline uint32_t __byte_perm(uint32_t x, uint32_t y, uint32_t sel) {
    const uint64_t v = ((uint64_t)y << 32) | x;
    uint32_t r = 0;
    for (int i = 0; i < 4; i++) r |= (uint32_t)((v >> (8 * ((sel >> (4 * i)) & 7u))) & 0xFFu) << (8 * i);
    return r;
}
template <class T> inline T __ldg(const T* p) { return *p; }
template <class T> inline T __ldcg(const T* p) { return *p; }

template <class A, class B> inline typename std::common_type<A, B>::type min(A a, B b) {
    typedef typename std::common_type<A, B>::type C;
    return (C)a < (C)b ? (C)a : (C)b;
}
template <class A, class B> inline typename std::common_type<A, B>::type max(A a, B b) {
    typedef typename std::common_type<A, B>::type C;
    return (C)a > (C)b ? (C)a : (C)b;
}

// ---- atomics (fibres switch only at yield points)
template <class T, class V> inline T atomicAdd(T* p, V v) { T o = *p; *p = (T)(o + (T)v); return o; }
template <class T, class V> inline T atomicMin(T* p, V v) { T o = *p; if ((T)v < o) *p = (T)v; return o; }
template <class T, class V> inline T atomicMax(T* p, V v) { T o = *p; if ((T)v > o) *p = (T)v; return o; }
template <class T, class V> inline T atomicOr(T* p, V v) { T o = *p; *p = (T)(o | (T)v); return o; }
template <class T, class V> inline T atomicExch(T* p, V v) { T o = *p; *p = (T)v; return o; }
template <class T, class C, class V> inline T atomicCAS(T* p, C cmp, V v) { T o = *p; if (o == (T)cmp) *p = (T)v; return o; }
