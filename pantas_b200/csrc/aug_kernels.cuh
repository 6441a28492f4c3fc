// aug_kernels.cuh -- device code of the augment pass: tables + event sink (tables.cuh), TMA helpers,
// the fast path (team_tiles.cuh), the exact per-record kernel and the per-chunk epilogue.  Included by
// pantas_aug.cu inside its anonymous namespace (after line_core.cuh) -- and, with PT_EMU defined, by the
// CPU-only test harness tests/hostsim/fastsim.cpp, which runs the same kernels thread by thread.
#pragma once

#include "tables.cuh"

#ifndef PT_EMU
#define PT_DYNAMIC_SMEM(name) extern __shared__ __align__(128) uint8_t name[]

// ---------------------------------------------------------------- TMA helpers

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!ok);
}
// 1-D bulk copy global -> shared, completion signalled on the mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// the same for bytes that are read exactly once (the GAF stream): L2 evict-first, so that the stream does not push
// the node table's sectors out of the L2
__device__ __forceinline__ void tma_load_1d_stream(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    uint64_t policy;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
                 : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// named barrier `id` (1..15) for `count` threads of the CTA (SASS: BAR.SYNC id, count)
__device__ __forceinline__ void named_barrier(uint32_t id, uint32_t count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
#else   // ---- test harness: the bulk copy is a memcpy that has completed when it returns
#define PT_DYNAMIC_SMEM(name) uint8_t* const name = emu::S().dyn
// (*bar = number of completed phases; a waiter yields until the phase with its parity is complete)
inline void mbar_init(uint64_t* bar, uint32_t) { *bar = 0; }
inline void mbar_expect_tx(uint64_t*, uint32_t) {}
inline void mbar_wait(uint64_t* bar, uint32_t parity) { while ((*bar & 1u) == parity) emu::yield(); }
inline void tma_load_1d(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    memcpy(dst_smem, src_gmem, bytes);
    *bar += 1;
}
inline void tma_load_1d_stream(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) { tma_load_1d(dst_smem, src_gmem, bytes, bar); }
inline void fence_async_smem() {}
inline void named_barrier(uint32_t id, uint32_t count) { emu_named_barrier(id, count); }
#endif

struct ChunkArgs {
    const uint8_t* gaf;
    uint64_t nbytes;
    int64_t file_off;
    int64_t thr;
    uint32_t n_tiles;
    uint32_t tile_bytes;    // bytes per tile of the fast path: a multiple of 16, <= the geometry's TILE
    uint32_t stream_hint;   // bit 0: load the GAF with an L2 evict-first policy
    uint32_t loose;         // 1: team-level barriers inside a tile, one CTA-wide barrier per tile (PANTAS_LOOSE)
    uint32_t ablate;        // diagnostics (PANTAS_ABLATE): 0 = the whole pass, k = every tile stops after phase k (timing only)
};

// why a record is handed to the slow path (pt_debug_counters)
enum { WHY_LONG = 0, WHY_COLUMNS, WHY_INTS, WHY_TAGS, WHY_CS, WHY_PATH, WHY_STEPS_FULL, WHY_WALK, WHY_LINES_FULL, WHY_OTHER };

__device__ __noinline__ void defer_line_impl(unsigned long long* sc, uint32_t* deferred, uint64_t deferred_cap, uint64_t chunk_pos,
                                            int64_t file_off, int why) {
    atomicAdd(&sc[SC_WHY + why], 1ull);
    const unsigned long long j = atomicAdd(&sc[SC_NDEFER], 1ull);
    if (j < deferred_cap) deferred[j] = (uint32_t)chunk_pos;
    else report_error_sc(sc, pt::PT_X_DEFER_FULL, file_off + (int64_t)chunk_pos);
}
__device__ __forceinline__ void defer_line(const Tables& T, uint64_t chunk_pos, int64_t file_off, int why = WHY_OTHER) {
    defer_line_impl(T.sc, T.deferred, T.deferred_cap, chunk_pos, file_off, why);
}

#include "team_tiles.cuh"

// Records the fast path hands over (longer than a tile's look-ahead, or anything unusual): the exact per-record code.  A warp
// takes a record (the usual case: a handful per chunk, their latency is what counts): all lanes copy its first DEF_STAGE bytes into shared memory with 16-byte loads, lane 0 parses from there
// (byte-serial parsing straight from global memory pays one L2 round trip per byte); a record that is longer than the stage is
// parsed again from global memory (LINE_DEFER: nothing was emitted).
#ifdef PT_EMU
constexpr int DEF_STAGE = 512;          // (small in the CPU tests: both paths run)
#else
constexpr int DEF_STAGE = 2048;
#endif
__global__ void __launch_bounds__(128) augment_deferred_kernel(ChunkArgs A, Tables T) {
    __shared__ __align__(16) uint8_t stage[4][DEF_STAGE];
    DevSink sink(T);
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    const unsigned long long n = min(T.sc[SC_NDEFER], (unsigned long long)T.deferred_cap);
    if (n > 4ull * 4ull * gridDim.x) {
        // many records (an input the fast path does not like): one record per THREAD, latency hidden by numbers
        for (unsigned long long j = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; j < n;
             j += (unsigned long long)gridDim.x * blockDim.x) {
            const uint64_t a = T.deferred[j];
            const uint64_t a16 = a & ~15ull;
            pt::LineCtx cx;
            cx.s = A.gaf + a16;
            const uint64_t rest = A.nbytes - a16;
            cx.lim = rest > 0x7ffffff0ull ? 0x7ffffff0 : (int)rest;
            cx.lim_final = true;
            cx.base_off = A.file_off + (int64_t)a16;
            pt::process_line(cx, (int)(a - a16), A.thr, sink);
        }
    } else
    for (unsigned long long j = blockIdx.x * 4ull + warp; j < n; j += (unsigned long long)gridDim.x * 4ull) {
        const uint64_t a = T.deferred[j];
        const uint64_t a16 = a & ~15ull;                     // word loads need an aligned base
        const uint64_t rest = A.nbytes - a16;
        const uint32_t m = rest < (uint64_t)DEF_STAGE ? (uint32_t)rest : (uint32_t)DEF_STAGE;
        const uint4* src = reinterpret_cast<const uint4*>(A.gaf + a16);      // (the chunk is readable up to nbytes rounded up to 16)
        uint4* dst = reinterpret_cast<uint4*>(stage[warp]);
        for (uint32_t v = lane; v * 16u < m; v += 32u) dst[v] = src[v];
        __syncwarp();
        if (lane == 0) {
            pt::LineCtx cx;
            cx.s = stage[warp];
            cx.lim = (int)m;
            cx.lim_final = (uint64_t)m == rest;
            cx.base_off = A.file_off + (int64_t)a16;
            if (pt::process_line(cx, (int)(a - a16), A.thr, sink) == pt::LINE_DEFER) {
                cx.s = A.gaf + a16;
                cx.lim = rest > 0x7ffffff0ull ? 0x7ffffff0 : (int)rest;
                cx.lim_final = true;
                pt::process_line(cx, (int)(a - a16), A.thr, sink);
            }
        }
        __syncwarp();
    }
    uint32_t r = sink.rej;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
    if (lane == 0 && r) atomicAdd(&T.sc[SC_REJ], (unsigned long long)r);
}

// After both kernels of a chunk: fold the per-chunk scalars, reset the fast kernel's low-water mark.
__global__ void end_chunk_kernel(Tables T) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < T.team_cap; i += gridDim.x * blockDim.x) T.team_tile[i] = 0;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        T.sc[SC_DEFERRED_TOTAL] += min(T.sc[SC_NDEFER], (unsigned long long)T.deferred_cap);
        T.sc[SC_NDEFER] = 0;
        T.sc[SC_LWM] = 0;
    }
}

