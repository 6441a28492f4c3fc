// pantas_aug.cu -- sm_100a kernels + C ABI (include/pantas_aug.h) for pantas `augment`.
//
// Replaces the per-line loop of the reference,
//   /root/reference/scripts/alignments_augmentation_from_gaf.py:138-371   (REF:n)
// Data flow on the device (DESIGN.md has the full picture):
//
//   GAF bytes in HBM --cp.async.bulk (TMA 1-D, L2 evict-first)--> a CTA's shared-memory tile
//     fast_tiles.cuh   the fast path: persistent CTAs, barrier-separated phases over the tile (byte-parallel
//                      SWAR scan, records by role warps, one thread per path step for ids / walk / count)
//     line_core.cuh    exact thread-per-record path for every record the fast path declines
//                      (augment_deferred_kernel, bytes from global memory)
//     tables.cuh       NodeRec[idx] = one 32-byte sector per node: len, first-touch stamps, two
//                      inline out-links and the fused NC|RC counters (one RED.ADD.64 per path step);
//                      64-bit-key open-addressing tables for the remaining links (known: ovf,
//                      unknown: novel) and for deletion-derived IL/OL keys (sparse)
//   (aug_kernels.cuh holds the device code; this file adds the round-1a tile kernel kept for tests and the host side.)
//
// No tensor cores: nothing here is a contraction.  No CPU fallback: every entry
// point fails if the device is not sm_100.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/pantas_aug.h"
#include "line_core.cuh"

namespace {

#include "aug_kernels.cuh"

constexpr int N_BUCKETS = 32;       // walk order: perfect-match records by path length, then the rest

template <int THREADS, int MIN_CTAS>
__global__ void __launch_bounds__(THREADS, MIN_CTAS) augment_tiles_kernel(ChunkArgs A, Tables T) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t mbar;
    __shared__ uint32_t s_nlines;
    __shared__ uint32_t s_tile;
    __shared__ uint32_t s_bucket[N_BUCKETS];

    const int tid = threadIdx.x;
    uint8_t* buf = smem;                                      // [16 + tile + over + 16]
    const uint32_t buf_bytes = 32u + A.tile + A.over;
    uint32_t* list = reinterpret_cast<uint32_t*>(smem + ((buf_bytes + 127u) & ~127u));   // [list_cap]
    pt::LineRec* recs = reinterpret_cast<pt::LineRec*>(list + ((A.list_cap + 31u) & ~31u));   // [THREADS]

    if (tid == 0) mbar_init(&mbar, 1);
    __syncthreads();

    DevSink sink(T);
    const uint64_t nbytes16 = (A.nbytes + 15ull) & ~15ull;
    uint32_t parity = 0;
    unsigned long long my_lines = 0, my_tiles = 0;

    for (;;) {
        if (tid == 0) {
            s_tile = (uint32_t)atomicAdd(&T.sc[SC_TILE_NEXT], 1ull);
            s_nlines = 0;
        }
        if (tid < N_BUCKETS) s_bucket[tid] = 0;
        __syncthreads();
        const uint32_t tile = s_tile;
        if (tile >= A.n_tiles) break;

        const uint64_t t0 = (uint64_t)tile * A.tile;
        const uint64_t t1 = min(t0 + A.tile, A.nbytes);             // owned line starts are in [t0, t1)
        const uint64_t lo = tile ? t0 - 16 : 0;
        const uint64_t hi = min(t0 + A.tile + A.over, nbytes16);     // loaded bytes [lo, hi)
        const uint32_t skip = tile ? 0u : 16u;                       // buffer position 16 == byte t0
        if (tid == 0) {
            const uint32_t bytes = (uint32_t)(hi - lo);
            mbar_expect_tx(&mbar, bytes);
            tma_load_1d(buf + skip, A.gaf + lo, bytes, &mbar);
        }
        mbar_wait(&mbar, parity);
        parity ^= 1;

        // ---- phase 1: cooperative scan of [t0 - 1, t1) for line starts, CR, non-ASCII
        {
            const uint32_t owned = (uint32_t)(t1 - t0);
            const uint32_t v_end = (16u + owned + 15u) >> 4;
            if (tile == 0 && tid == 0 && A.nbytes > 0) {
                const uint32_t j = atomicAdd(&s_nlines, 1u);
                if (j < A.list_cap) list[j] = 16u; else defer_line(T, 0, A.file_off);
            }
            for (uint32_t v = tid + (tile ? 0u : 1u); v < v_end; v += THREADS) {
                const uint4 q4 = *reinterpret_cast<const uint4*>(buf + 16u * v);
                const uint32_t w[4] = {q4.x, q4.y, q4.z, q4.w};
                const uint64_t v_abs = t0 + 16ull * v - 16ull;      // file-chunk position of the vector
                const bool tail = v_abs + 16 > A.nbytes;            // bytes past the chunk end are garbage
                uint32_t any_hi = (w[0] | w[1] | w[2] | w[3]) & 0x80808080u;
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const uint32_t a = w[k] & 0x7F7F7F7Fu;
                    // 0x80 in every byte whose low 7 bits are in [0x0A, 0x0D]
                    uint32_t m = (a + 0x76767676u) & ~(a + 0x72727272u) & 0x80808080u;
                    while (m) {
                        const int byte = (__ffs(m) - 1) >> 3;
                        m &= m - 1;
                        const uint32_t x = 16u * v + 4u * k + byte;        // buffer position
                        const uint64_t abs_pos = v_abs + 4u * k + byte;
                        if (abs_pos >= A.nbytes) continue;
                        const uint32_t c = (w[k] >> (8 * byte)) & 0xFFu;
                        if (c == '\n') {
                            if (x + 1 >= 16u && abs_pos + 1 < t1) {
                                const uint32_t j = atomicAdd(&s_nlines, 1u);
                                if (j < A.list_cap) list[j] = x + 1;
                                else defer_line(T, abs_pos + 1, A.file_off);
                            }
                        } else if (c == '\r' && x >= 16u) {
                            if (abs_pos + 1 < A.nbytes && buf[x + 1] != '\n')
                                report_error(T, pt::PT_U_BARE_CR, A.file_off + (int64_t)abs_pos);
                        }
                    }
                }
                if (any_hi) {
                    if (!tail && v >= 1) report_error(T, pt::PT_U_NON_ASCII, A.file_off + (int64_t)v_abs);
                    else
                        for (int b = 0; b < 16; b++)
                            if (v_abs + b < A.nbytes && 16u * v + b >= 16u && buf[16u * v + b] >= 0x80)
                                report_error(T, pt::PT_U_NON_ASCII, A.file_off + (int64_t)(v_abs + b));
                }
            }
        }
        __syncthreads();

        const uint32_t total = s_nlines;
        const uint32_t nl = min(total, A.list_cap);
        if (tid == 0) { my_lines += total; my_tiles++; }
        pt::LineCtx cx;
        cx.s = buf;
        cx.lim = (int)(16u + (uint32_t)(min(hi, A.nbytes) - t0));
        cx.lim_final = (hi >= A.nbytes);
        cx.base_off = A.file_off + (int64_t)t0 - 16;

        for (uint32_t base = 0; base < nl; base += THREADS) {
            // ---- phase 2: front half, one thread per record; survivors pick a bucket
            pt::LineRec rec;
            int cls = pt::LINE_DONE;
            uint32_t bucket = 0, rank = 0;
            const uint32_t l = base + tid;
            if (l < nl) {
                const uint32_t p = list[l];
                cls = pt::front_line(cx, (int)p, A.thr, sink, rec);
                // records with a non-trivial cs string take the slow, divergent walk: they are
                // redone by augment_deferred_kernel so that no tile waits for them
                if (cls == pt::LINE_DEFER || cls == pt::LINE_GENERAL) {
                    defer_line(T, t0 + p - 16u, A.file_off);
                    cls = pt::LINE_DONE;
                } else if (cls == pt::LINE_SIMPLE) {
                    bucket = min((uint32_t)(rec.b5 - rec.a5) >> 4, (uint32_t)N_BUCKETS - 1u);
                    rank = atomicAdd(&s_bucket[bucket], 1u);
                }
            }
            __syncthreads();
            // ---- regroup: records of one class and similar path length sit next to each other
            uint32_t n_surv = 0;
            {
                uint32_t before = 0;
#pragma unroll
                for (int bkt = 0; bkt < N_BUCKETS; bkt++) {
                    const uint32_t c = s_bucket[bkt];
                    if ((uint32_t)bkt < bucket) before += c;
                    n_surv += c;
                }
                if (cls == pt::LINE_SIMPLE) recs[before + rank] = rec;
            }
            __syncthreads();
            if (tid < N_BUCKETS) s_bucket[tid] = 0;     // next round / next tile (ordered by the sync below)
            // ---- phase 3: walk, one thread per surviving record
            if ((uint32_t)tid < n_surv) {
                const pt::LineRec r = recs[tid];
                pt::walk_simple(cx, r, sink);
            }
            __syncthreads();   // every read of buf / list / recs is done before they are reused
        }
        if (nl == 0) __syncthreads();
    }

    // rejected-record count: warp reduce, one RED per warp
    uint32_t r = sink.rej;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
    if ((tid & 31) == 0 && r) atomicAdd(&T.sc[SC_REJ], (unsigned long long)r);
    if (tid == 0) {
        if (my_lines) atomicAdd(&T.sc[SC_LINES], my_lines);
        if (my_tiles) atomicAdd(&T.sc[SC_TILES], my_tiles);
    }
}

}  // namespace

// ================================================================== host side

struct pt_ctx {
    int device;
    int sm_count;
    cudaStream_t stream;
    bool own_stream;
    cudaStream_t copy_stream;
    cudaEvent_t ev_copied[2], ev_done[2], ev_t0, ev_t1;
    bool have_graph;
    Tables T;
    uint64_t n_edges;
    uint64_t novel_cap, sparse_cap, ovf_cap;
    bool epoch_open;             // 32-bit epoch state may be non-zero
    uint64_t epoch_end;          // highest file offset seen in the open epoch
    uint64_t folds;
    unsigned long long* cursor;         // 2 compaction cursors
    uint8_t* stage[2];
    uint64_t stage_bytes;
    int64_t next_ticket;
    uint64_t launches;
    uint32_t tile, over, list_cap, threads;
    int ctas_per_sm;
    uint32_t kernel_ver;         // 2: warp-autonomous fast path (fast_tiles.cuh) + slow path; 1: tile kernel of round 1a
    uint32_t fast_geo;           // mini-tile bytes of the fast path
    int fast_ctas_per_sm;
    bool profile;
    cudaEvent_t* prof_ev;        // pairs
    uint32_t prof_n, prof_cap;
    char err[512];
};

static int fail_cuda(pt_ctx* c, cudaError_t e, const char* what) {
    if (c) snprintf(c->err, sizeof c->err, "%s: %s", what, cudaGetErrorString(e));
    return PT_ERR_CUDA;
}
#define CK(call)                                                   \
    do {                                                           \
        cudaError_t e_ = (call);                                   \
        if (e_ != cudaSuccess) return fail_cuda(ctx, e_, #call);   \
    } while (0)

static int fail_msg(pt_ctx* c, int code, const char* msg) {
    if (c) snprintf(c->err, sizeof c->err, "%s", msg);
    return code;
}

static uint64_t pow2_at_least(uint64_t v) {
    uint64_t p = 1;
    while (p < v) p <<= 1;
    return p;
}
static uint32_t env_u32(const char* name, uint32_t dflt) {
    const char* v = getenv(name);
    if (!v || !*v) return dflt;
    long x = strtol(v, NULL, 10);
    return x > 0 ? (uint32_t)x : dflt;
}
static int grid_for(uint64_t n, int threads, int cap_blocks) {
    uint64_t b = (n + threads - 1) / threads;
    if (b < 1) b = 1;
    if (b > (uint64_t)cap_blocks) b = cap_blocks;
    return (int)b;
}

static void free_graph_tables(pt_ctx* ctx) {
    Tables& T = ctx->T;
    cudaFree(T.nodes); cudaFree(T.il_adj32); cudaFree(T.ol_adj32); cudaFree(T.ovf); cudaFree(T.ovf_edge);
    cudaFree(T.inl_edge); cudaFree(T.novel); cudaFree(T.sparse); cudaFree(T.nc64); cudaFree(T.il_adj64);
    cudaFree(T.ol_adj64); cudaFree(T.il_st64); cudaFree(T.ol_st64); cudaFree(T.rc64);
    T.nodes = NULL; T.il_adj32 = T.ol_adj32 = NULL; T.ovf = NULL; T.ovf_edge = T.inl_edge = NULL;
    T.novel = T.sparse = NULL; T.nc64 = T.il_adj64 = T.ol_adj64 = NULL; T.il_st64 = T.ol_st64 = NULL; T.rc64 = NULL;
}

// fold the open epoch's 32-bit state into the 64-bit totals (tables.cuh)
static int fold_epoch(pt_ctx* ctx) {
    if (!ctx->epoch_open) return 0;
    fold_epoch_kernel<<<grid_for(ctx->T.n_nodes, 256, ctx->sm_count * 8), 256, 0, ctx->stream>>>(ctx->T);
    if (cudaGetLastError() != cudaSuccess) return fail_msg(ctx, PT_ERR_CUDA, "fold_epoch_kernel launch failed");
    ctx->launches += 1;
    ctx->folds += 1;
    ctx->epoch_open = false;
    return 0;
}

extern "C" {

int pt_abi_version(void) { return PT_ABI_VERSION; }

const char* pt_strerror(int code) {
    switch (code) {
        case 0: return "ok";
        case PT_ERR_CUDA: return "CUDA runtime error";
        case PT_ERR_ARG: return "bad argument";
        case PT_ERR_STATE: return "call out of order";
        case PT_ERR_NOMEM: return "out of memory";
        case PT_ERR_NODEVICE: return "no sm_100 CUDA device";
        case pt::PT_E_COLUMNS: return "GAF record has fewer than 12 columns (reference: IndexError)";
        case pt::PT_E_MAPQ: return "MAPQ column is not an integer (reference: ValueError)";
        case pt::PT_E_COORD: return "path length/start/end is not an integer (reference: ValueError)";
        case pt::PT_E_NO_DV: return "record passes the filters but has no dv:f: tag (reference: ValueError)";
        case pt::PT_E_EMPTY_PATH: return "path column holds no step (reference: AssertionError)";
        case pt::PT_E_UNKNOWN_NODE: return "path step is not a node of the GFA (reference: KeyError)";
        case pt::PT_E_CS_SHORT: return "cs string ends before the path does (reference: IndexError)";
        case pt::PT_U_TILDE: return "unsupported: '~' op in cs string";
        case pt::PT_U_NON_ASCII: return "unsupported: non-ASCII byte in GAF";
        case pt::PT_U_BARE_CR: return "unsupported: lone carriage return in GAF";
        case pt::PT_U_BIG_INT: return "unsupported: integer field too large";
        case pt::PT_U_UNDERSCORE: return "unsupported: '_' inside an integer field";
        case pt::PT_U_POSITION: return "unsupported: IL/OL position out of range";
        case pt::PT_X_NOVEL_FULL: return "novel-link table full (raise PANTAS_NOVEL_CAP)";
        case pt::PT_X_SPARSE_FULL: return "sparse IL/OL table full (raise PANTAS_SPARSE_CAP)";
        case pt::PT_X_DEFER_FULL: return "long-record list full";
        default: return "unknown code";
    }
}

int pt_create(int device, pt_ctx** out) {
    if (!out) return PT_ERR_ARG;
    *out = NULL;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || device < 0 || device >= n) return PT_ERR_NODEVICE;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return PT_ERR_NODEVICE;
    if (prop.major != 10) return PT_ERR_NODEVICE;      // sm_100a cubin only; no fallback path exists
    pt_ctx* ctx = (pt_ctx*)calloc(1, sizeof(pt_ctx));
    if (!ctx) return PT_ERR_NOMEM;
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    if (cudaSetDevice(device) != cudaSuccess) { free(ctx); return PT_ERR_CUDA; }
    cudaError_t e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking);
    for (int k = 0; k < 2 && e == cudaSuccess; k++) {
        e = cudaEventCreateWithFlags(&ctx->ev_copied[k], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_done[k], cudaEventDisableTiming);
    }
    if (e == cudaSuccess) e = cudaEventCreate(&ctx->ev_t0);
    if (e == cudaSuccess) e = cudaEventCreate(&ctx->ev_t1);
    if (e == cudaSuccess) e = cudaMalloc(&ctx->T.sc, SC_COUNT * sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMalloc(&ctx->cursor, 2 * sizeof(unsigned long long));
    if (e != cudaSuccess) { free(ctx); return PT_ERR_CUDA; }
    ctx->own_stream = true;
    ctx->stage_bytes = (uint64_t)env_u32("PANTAS_STAGE_MB", 256) << 20;
    ctx->tile = env_u32("PANTAS_TILE_KB", 56) << 10;
    ctx->over = env_u32("PANTAS_OVER_KB", 4) << 10;
    // byte-granular overrides (tests drive tiny tiles through the deferral path)
    ctx->tile = (env_u32("PANTAS_TILE_BYTES", ctx->tile) + 15u) & ~15u;
    ctx->over = (env_u32("PANTAS_OVER_BYTES", ctx->over) + 15u) & ~15u;
    if (ctx->tile < 64) ctx->tile = 64;
    if (ctx->over < 16) ctx->over = 16;
    ctx->list_cap = env_u32("PANTAS_LIST_CAP", 512);
    ctx->threads = env_u32("PANTAS_THREADS", 256);
    if (ctx->threads != 32 && ctx->threads != 64 && ctx->threads != 128 && ctx->threads != 192 && ctx->threads != 256 && ctx->threads != 384 &&
        ctx->threads != 512)
        ctx->threads = 256;
    ctx->kernel_ver = env_u32("PANTAS_KERNEL", 2);
    ctx->fast_geo = env_u32("PANTAS_FAST_T", 32768);
    *out = ctx;
    return 0;
}

void pt_destroy(pt_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    free_graph_tables(ctx);
    cudaFree(ctx->T.sc); cudaFree(ctx->T.deferred); cudaFree(ctx->cursor); cudaFree(ctx->stage[0]); cudaFree(ctx->stage[1]);
    for (int k = 0; k < 2; k++) { cudaEventDestroy(ctx->ev_copied[k]); cudaEventDestroy(ctx->ev_done[k]); }
    cudaEventDestroy(ctx->ev_t0); cudaEventDestroy(ctx->ev_t1);
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    cudaStreamDestroy(ctx->copy_stream);
    for (uint32_t k = 0; k < ctx->prof_cap; k++) cudaEventDestroy(ctx->prof_ev[k]);
    free(ctx->prof_ev);
    free(ctx);
}

const char* pt_last_error(pt_ctx* ctx) { return ctx ? ctx->err : "null context"; }

int pt_set_stream(pt_ctx* ctx, void* cuda_stream) {
    if (!ctx) return PT_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    ctx->stream = (cudaStream_t)cuda_stream;
    ctx->own_stream = false;
    return 0;
}

static int reset_counts_impl(pt_ctx* ctx) {
    Tables& T = ctx->T;
    const int cap = ctx->sm_count * 8;
    const uint64_t N = T.n_nodes, E = ctx->n_edges;
    reset_nodes_kernel<<<grid_for(N, 256, cap), 256, 0, ctx->stream>>>(T);
    CK(cudaMemsetAsync(T.il_adj32, 0, N * sizeof(int32_t), ctx->stream));
    CK(cudaMemsetAsync(T.ol_adj32, 0, N * sizeof(int32_t), ctx->stream));
    CK(cudaMemsetAsync(T.nc64, 0, N * sizeof(long long), ctx->stream));
    CK(cudaMemsetAsync(T.il_adj64, 0, N * sizeof(long long), ctx->stream));
    CK(cudaMemsetAsync(T.ol_adj64, 0, N * sizeof(long long), ctx->stream));
    CK(cudaMemsetAsync(T.rc64, 0, (E ? E : 1) * sizeof(long long), ctx->stream));
    clear_ovf_kernel<<<grid_for(ctx->ovf_cap, 256, cap), 256, 0, ctx->stream>>>(T.ovf, T.ovf_edge, ctx->ovf_cap, 0);
    clear_side_kernel<<<grid_for(T.novel_mask + 1, 256, cap), 256, 0, ctx->stream>>>(T.novel, T.novel_mask + 1);
    clear_side_kernel<<<grid_for(T.sparse_mask + 1, 256, cap), 256, 0, ctx->stream>>>(T.sparse, T.sparse_mask + 1);
    unsigned long long sc[SC_COUNT];
    memset(sc, 0, sizeof sc);
    sc[SC_ERR] = ~0ull;
    CK(cudaMemcpyAsync(T.sc, sc, sizeof sc, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));       // sc[] is a stack buffer
    CK(cudaGetLastError());
    ctx->launches += 4;
    ctx->epoch_open = false;
    ctx->epoch_end = 0;
    T.epoch_base = 0;
    return 0;
}

int pt_set_graph(pt_ctx* ctx, const uint32_t* node_len, uint64_t n_nodes, uint32_t min_id, const uint64_t* edge_keys,
                 uint64_t n_edges, uint64_t novel_cap, uint64_t sparse_cap) {
    if (!ctx || !node_len || n_nodes == 0 || (n_edges && !edge_keys)) return fail_msg(ctx, PT_ERR_ARG, "pt_set_graph: bad argument");
    if (n_nodes >= 0xFFFFFFFFull || n_edges >= 0xFFFFFFFFull) return fail_msg(ctx, PT_ERR_ARG, "pt_set_graph: more than 2^32-2 nodes or links");
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    Tables& T = ctx->T;
    free_graph_tables(ctx);
    ctx->have_graph = false;

    T.n_nodes = n_nodes;
    T.min_id = min_id;
    T.epoch_base = 0;
    if (!novel_cap) novel_cap = env_u32("PANTAS_NOVEL_CAP", 0);
    if (!sparse_cap) sparse_cap = env_u32("PANTAS_SPARSE_CAP", 0);
    if (!novel_cap) novel_cap = (n_edges / 2 > (1u << 20)) ? n_edges / 2 : (1u << 20);
    if (!sparse_cap) sparse_cap = (n_nodes / 2 > (1u << 20)) ? n_nodes / 2 : (1u << 20);
    ctx->novel_cap = pow2_at_least(novel_cap);
    ctx->sparse_cap = pow2_at_least(sparse_cap);
    T.novel_mask = ctx->novel_cap - 1;
    T.sparse_mask = ctx->sparse_cap - 1;
    ctx->n_edges = n_edges;

    CK(cudaMalloc(&T.nodes, n_nodes * sizeof(NodeRec)));
    CK(cudaMalloc(&T.il_adj32, n_nodes * sizeof(int32_t)));
    CK(cudaMalloc(&T.ol_adj32, n_nodes * sizeof(int32_t)));
    CK(cudaMalloc(&T.inl_edge, 2 * n_nodes * sizeof(uint32_t)));
    CK(cudaMalloc(&T.nc64, n_nodes * sizeof(long long)));
    CK(cudaMalloc(&T.il_adj64, n_nodes * sizeof(long long)));
    CK(cudaMalloc(&T.ol_adj64, n_nodes * sizeof(long long)));
    CK(cudaMalloc(&T.il_st64, n_nodes * sizeof(unsigned long long)));
    CK(cudaMalloc(&T.ol_st64, n_nodes * sizeof(unsigned long long)));
    CK(cudaMalloc(&T.rc64, (n_edges ? n_edges : 1) * sizeof(long long)));
    CK(cudaMalloc(&T.novel, ctx->novel_cap * sizeof(SideSlot)));
    CK(cudaMalloc(&T.sparse, ctx->sparse_cap * sizeof(SideSlot)));

    uint32_t* d_len = NULL;
    uint64_t* d_keys = NULL;
    unsigned long long* d_stats = NULL;          // [0] bad / duplicate keys, [1] links that are not inline
    CK(cudaMalloc(&d_len, n_nodes * sizeof(uint32_t)));
    CK(cudaMalloc(&d_keys, (n_edges ? n_edges : 1) * sizeof(uint64_t)));
    CK(cudaMalloc(&d_stats, 2 * sizeof(unsigned long long)));
    CK(cudaMemcpyAsync(d_len, node_len, n_nodes * sizeof(uint32_t), cudaMemcpyDefault, ctx->stream));
    if (n_edges) CK(cudaMemcpyAsync(d_keys, edge_keys, n_edges * sizeof(uint64_t), cudaMemcpyDefault, ctx->stream));
    CK(cudaMemsetAsync(d_stats, 0, 2 * sizeof(unsigned long long), ctx->stream));
    const int cap = ctx->sm_count * 8;
    init_nodes_kernel<<<grid_for(n_nodes, 256, cap), 256, 0, ctx->stream>>>(T, d_len);
    if (n_edges) inline_edges_kernel<<<grid_for(n_edges, 256, cap), 256, 0, ctx->stream>>>(T, d_keys, n_edges, d_stats);
    unsigned long long stats[2] = {0, 0};
    CK(cudaMemcpyAsync(stats, d_stats, sizeof stats, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaGetLastError());
    // known links that are not inline: open addressing at load <= 0.5
    ctx->ovf_cap = pow2_at_least(stats[1] * 2 + 1024);
    T.ovf_mask = ctx->ovf_cap - 1;
    CK(cudaMalloc(&T.ovf, ctx->ovf_cap * sizeof(OvfSlot)));
    CK(cudaMalloc(&T.ovf_edge, ctx->ovf_cap * sizeof(uint32_t)));
    clear_ovf_kernel<<<grid_for(ctx->ovf_cap, 256, cap), 256, 0, ctx->stream>>>(T.ovf, T.ovf_edge, ctx->ovf_cap, 1);
    if (n_edges) ovf_edges_kernel<<<grid_for(n_edges, 256, cap), 256, 0, ctx->stream>>>(T, d_keys, n_edges, d_stats);
    CK(cudaMemcpyAsync(stats, d_stats, sizeof stats, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaGetLastError());
    cudaFree(d_len); cudaFree(d_keys); cudaFree(d_stats);
    ctx->launches += 4;
    if (stats[0]) return fail_msg(ctx, PT_ERR_ARG, "pt_set_graph: edge_keys hold duplicates or out-of-range node indices");
    ctx->have_graph = true;
    return reset_counts_impl(ctx);
}

int pt_reset_counts(pt_ctx* ctx) {
    if (!ctx || !ctx->have_graph) return fail_msg(ctx, PT_ERR_STATE, "pt_reset_counts: no graph");
    CK(cudaSetDevice(ctx->device));
    return reset_counts_impl(ctx);
}

static int launch_chunk(pt_ctx* ctx, const uint8_t* gaf_dev, uint64_t nbytes, uint64_t file_offset, int64_t thr) {
    if (nbytes == 0) return 0;
    if (((uintptr_t)gaf_dev & 15) != 0) return fail_msg(ctx, PT_ERR_ARG, "GAF chunk must be 16-byte aligned");
    Tables& T = ctx->T;
    if (nbytes > 0xF0000000ull) return fail_msg(ctx, PT_ERR_ARG, "chunk larger than 3.75 GiB: split it");
    // 32-bit stamps / counters are relative to an epoch of < 4 GiB of GAF: fold when this chunk does not fit
    if (ctx->epoch_open && (file_offset < (uint64_t)T.epoch_base || file_offset + nbytes - (uint64_t)T.epoch_base > EPOCH_SPAN)) {
        int rc = fold_epoch(ctx);
        if (rc) return rc;
    }
    if (!ctx->epoch_open) {
        T.epoch_base = (int64_t)file_offset;
        ctx->epoch_open = true;
        ctx->epoch_end = file_offset;
    }
    if (file_offset + nbytes > ctx->epoch_end) ctx->epoch_end = file_offset + nbytes;
    // second-pass list: records with a non-trivial cs string, records longer than the
    // look-ahead, list overflow.  A valid record is >= 40 bytes, so this holds all of them.
    const uint64_t want = nbytes / 40 + 4096;
    if (want > T.deferred_cap) {
        CK(cudaStreamSynchronize(ctx->stream));
        cudaFree(T.deferred);
        T.deferred = NULL;
        T.deferred_cap = 0;
        CK(cudaMalloc(&T.deferred, want * sizeof(uint32_t)));
        T.deferred_cap = want;
    }
    ChunkArgs A;
    A.gaf = gaf_dev;
    A.nbytes = nbytes;
    A.file_off = (int64_t)file_offset;
    A.thr = thr;
    A.tile = ctx->tile;
    A.over = ctx->over;
    A.list_cap = ctx->list_cap;
    {   // measured: -4 % kernel time, -25 % DRAM reads (PANTAS_STREAM_HINT=0 switches it off)
        const char* h = getenv("PANTAS_STREAM_HINT");
        A.stream_hint = (h && h[0] == '0') ? 0u : 1u;
        if (env_u32("PANTAS_PHASE_CLOCKS", 0)) A.stream_hint |= 2u;       // diagnostics: per-phase cycle counters (pt_debug_counters)
    }
    const uint64_t n_tiles = (nbytes + ctx->tile - 1) / ctx->tile;
    if (n_tiles > 0xFFFFFFF0ull) return fail_msg(ctx, PT_ERR_ARG, "chunk too large");
    A.n_tiles = (uint32_t)n_tiles;
    const size_t smem = ((32 + (size_t)ctx->tile + ctx->over + 127) & ~(size_t)127) +
                        (size_t)((ctx->list_cap + 31u) & ~31u) * 4 + (size_t)ctx->threads * sizeof(pt::LineRec);
    void (*kern)(ChunkArgs, Tables) = ctx->threads == 32    ? augment_tiles_kernel<32, 20>
                                      : ctx->threads == 64  ? augment_tiles_kernel<64, 12>
                                      : ctx->threads == 128 ? augment_tiles_kernel<128, 6>
                                      : ctx->threads == 192 ? augment_tiles_kernel<192, 4>
                                      : ctx->threads == 384 ? augment_tiles_kernel<384, 2>
                                      : ctx->threads == 512 ? augment_tiles_kernel<512, 1>
                                                            : augment_tiles_kernel<256, 3>;
    if (ctx->ctas_per_sm == 0) {
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int occ = 0;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, (int)ctx->threads, smem));
        if (occ < 1) return fail_msg(ctx, PT_ERR_ARG, "tile does not fit shared memory");
        ctx->ctas_per_sm = occ;
    }
    uint64_t grid = (uint64_t)ctx->sm_count * ctx->ctas_per_sm;
    if (grid > n_tiles) grid = n_tiles;
    // ---- fast path geometry (kernel_ver 2)
    void (*fkern)(ChunkArgs, Tables) = NULL;
    size_t fsmem = 0;
    uint32_t f_tiles = 0, f_grid = 0;
    uint32_t f_threads = 0;
    if (ctx->kernel_ver == 2) {
        typedef fastp::Geo<24576, 1024, 256> GA;
        typedef fastp::Geo<32768, 1024, 512> GB;
        typedef fastp::Geo<16384, 1024, 256> GC;
        typedef fastp::Geo<12288, 1024, 128> GD;
        typedef fastp::Geo<16384, 1024, 128> GE;
        typedef fastp::Geo<32768, 1024, 256> GF;
        typedef fastp::Geo<20480, 1024, 256> GG;
        typedef fastp::Geo<8192, 1024, 128> GH;
        typedef fastp::Geo<8192, 1024, 256> GI;
        typedef fastp::Geo<49152, 1024, 768> GJ;
        typedef fastp::Geo<40960, 1024, 512> GK;
        typedef fastp::Geo<49152, 1024, 1024> GL;
        typedef fastp::Geo<28672, 1024, 512> GM;
        typedef fastp::Geo<32768, 1024, 384> GN;
        typedef fastp::Geo<30720, 1024, 512> GO;
        typedef fastp::Geo<1024, 256, 64> GT;    // tests: many tile boundaries, records longer than the look-ahead
        uint32_t ft;
#define PT_PICK(Gx) { fkern = fastp::augment_fast_kernel<Gx>; fsmem = (size_t)Gx::SMEM_BYTES; ft = Gx::TILE; f_threads = Gx::THREADS; }
        if (ctx->fast_geo == 1024) PT_PICK(GT)
        else if (ctx->fast_geo == 32768) PT_PICK(GB)
        else if (ctx->fast_geo == 16384) PT_PICK(GC)
        else if (ctx->fast_geo == 12288) PT_PICK(GD)
        else if (ctx->fast_geo == 16385) PT_PICK(GE)
        else if (ctx->fast_geo == 32769) PT_PICK(GF)
        else if (ctx->fast_geo == 20480) PT_PICK(GG)
        else if (ctx->fast_geo == 8192) PT_PICK(GH)
        else if (ctx->fast_geo == 8193) PT_PICK(GI)
        else if (ctx->fast_geo == 49152) PT_PICK(GJ)
        else if (ctx->fast_geo == 40960) PT_PICK(GK)
        else if (ctx->fast_geo == 49153) PT_PICK(GL)
        else if (ctx->fast_geo == 28672) PT_PICK(GM)
        else if (ctx->fast_geo == 32770) PT_PICK(GN)
        else if (ctx->fast_geo == 30720) PT_PICK(GO)
        else if (ctx->fast_geo == 24576) PT_PICK(GA)
        else PT_PICK(GA)
#undef PT_PICK
        if (ctx->fast_ctas_per_sm == 0) {
            CK(cudaFuncSetAttribute(fkern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsmem));
            int occ = 0;
            CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fkern, (int)f_threads, fsmem));
            if (occ < 1) return fail_msg(ctx, PT_ERR_ARG, "fast path does not fit shared memory");
            ctx->fast_ctas_per_sm = occ;
        }
        const uint64_t nt = (nbytes + ft - 1) / ft;
        f_tiles = (uint32_t)nt;
        uint64_t g = (uint64_t)ctx->sm_count * ctx->fast_ctas_per_sm;
        if (g > nt) g = nt;
        f_grid = (uint32_t)g;
    }
    if (ctx->profile) {
        if (ctx->prof_n + 3 > ctx->prof_cap) {
            uint32_t nc = ctx->prof_cap ? ctx->prof_cap * 2 : 96;
            cudaEvent_t* ne = (cudaEvent_t*)realloc(ctx->prof_ev, nc * sizeof(cudaEvent_t));
            if (!ne) return fail_msg(ctx, PT_ERR_NOMEM, "profile events");
            ctx->prof_ev = ne;
            for (uint32_t k = ctx->prof_cap; k < nc; k++) CK(cudaEventCreate(&ctx->prof_ev[k]));
            ctx->prof_cap = nc;
        }
        CK(cudaEventRecord(ctx->prof_ev[ctx->prof_n], ctx->stream));
    }
    if (ctx->kernel_ver == 2) {
        ChunkArgs F = A;
        F.n_tiles = f_tiles;
        fkern<<<f_grid, f_threads, fsmem, ctx->stream>>>(F, T);
    } else {
        kern<<<(unsigned)grid, ctx->threads, smem, ctx->stream>>>(A, T);
    }
    if (ctx->profile) CK(cudaEventRecord(ctx->prof_ev[ctx->prof_n + 1], ctx->stream));
    augment_deferred_kernel<<<ctx->sm_count * 8, 128, 0, ctx->stream>>>(A, T);
    end_chunk_kernel<<<1, 1, 0, ctx->stream>>>(T);
    if (ctx->profile) {      // fast path + exact per-record path = "the augment pass" over this chunk
        CK(cudaEventRecord(ctx->prof_ev[ctx->prof_n + 2], ctx->stream));
        ctx->prof_n += 3;
    }
    CK(cudaGetLastError());
    ctx->launches += 3;
    return 0;
}

int pt_process_chunk(pt_ctx* ctx, const uint8_t* gaf_dev, uint64_t nbytes, uint64_t file_offset, int64_t mapq_thr) {
    if (!ctx || !ctx->have_graph) return fail_msg(ctx, PT_ERR_STATE, "pt_process_chunk: no graph");
    if (nbytes && !gaf_dev) return fail_msg(ctx, PT_ERR_ARG, "pt_process_chunk: null chunk");
    CK(cudaSetDevice(ctx->device));
    return launch_chunk(ctx, gaf_dev, nbytes, file_offset, mapq_thr);
}

int pt_set_stage_bytes(pt_ctx* ctx, uint64_t bytes) {
    if (!ctx || bytes < (1u << 16)) return PT_ERR_ARG;
    if (ctx->stage[0]) return fail_msg(ctx, PT_ERR_STATE, "staging buffers already allocated");
    ctx->stage_bytes = (bytes + 15) & ~15ull;
    return 0;
}
uint64_t pt_stage_bytes(pt_ctx* ctx) { return ctx ? ctx->stage_bytes : 0; }

int64_t pt_process_host(pt_ctx* ctx, const uint8_t* gaf_host, uint64_t nbytes, uint64_t file_offset, int64_t mapq_thr) {
    if (!ctx || !ctx->have_graph) return fail_msg(ctx, PT_ERR_STATE, "pt_process_host: no graph");
    if (nbytes > ctx->stage_bytes) return fail_msg(ctx, PT_ERR_ARG, "pt_process_host: chunk larger than pt_stage_bytes()");
    if (nbytes && !gaf_host) return fail_msg(ctx, PT_ERR_ARG, "pt_process_host: null chunk");
    CK(cudaSetDevice(ctx->device));
    if (!ctx->stage[0]) {
        CK(cudaMalloc(&ctx->stage[0], ctx->stage_bytes + 16));
        CK(cudaMalloc(&ctx->stage[1], ctx->stage_bytes + 16));
    }
    const int64_t ticket = ctx->next_ticket++;
    const int k = (int)(ticket & 1);
    if (ticket >= 2) CK(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_done[k], 0));   // stage k is free again
    if (nbytes) CK(cudaMemcpyAsync(ctx->stage[k], gaf_host, nbytes, cudaMemcpyHostToDevice, ctx->copy_stream));
    CK(cudaEventRecord(ctx->ev_copied[k], ctx->copy_stream));
    CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_copied[k], 0));
    int rc = launch_chunk(ctx, ctx->stage[k], nbytes, file_offset, mapq_thr);
    if (rc) return rc;
    CK(cudaEventRecord(ctx->ev_done[k], ctx->stream));
    return ticket;
}

int pt_wait_copy(pt_ctx* ctx, int64_t ticket) {
    if (!ctx || ticket < 0 || ticket >= ctx->next_ticket) return PT_ERR_ARG;
    if (ticket + 2 < ctx->next_ticket) return 0;          // its stage has been re-used: long done
    CK(cudaEventSynchronize(ctx->ev_copied[ticket & 1]));
    return 0;
}

int pt_sync(pt_ctx* ctx) {
    if (!ctx) return PT_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaGetLastError());
    return 0;
}

int pt_error(pt_ctx* ctx, uint64_t* bad_offset, int* code) {
    if (!ctx || !ctx->have_graph) return fail_msg(ctx, PT_ERR_STATE, "pt_error: no graph");
    CK(cudaSetDevice(ctx->device));
    unsigned long long w = 0;
    CK(cudaMemcpyAsync(&w, ctx->T.sc + SC_ERR, sizeof w, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (w == ~0ull) {
        if (code) *code = 0;
        if (bad_offset) *bad_offset = 0;
    } else {
        if (code) *code = (int)(w & 0xFF);
        if (bad_offset) *bad_offset = w >> 8;
    }
    return 0;
}

int pt_finalize(pt_ctx* ctx, uint64_t* n_novel, uint64_t* n_sparse) {
    if (!ctx || !ctx->have_graph) return fail_msg(ctx, PT_ERR_STATE, "pt_finalize: no graph");
    CK(cudaSetDevice(ctx->device));
    unsigned long long sc[SC_COUNT];
    CK(cudaMemcpyAsync(sc, ctx->T.sc, sizeof sc, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaGetLastError());
    if (n_novel) *n_novel = sc[SC_NOVEL_USED];
    if (n_sparse) *n_sparse = sc[SC_SPARSE_USED];
    return 0;
}

int pt_export_dense(pt_ctx* ctx, int64_t* sums_dev, uint64_t sums_len, int64_t* stamps_dev, uint64_t stamps_len) {
    if (!ctx || !ctx->have_graph) return fail_msg(ctx, PT_ERR_STATE, "pt_export_dense: no graph");
    const uint64_t N = ctx->T.n_nodes, E = ctx->n_edges;
    if (!sums_dev || !stamps_dev || sums_len < 3 * N + E + 4 || stamps_len < 2 * N)
        return fail_msg(ctx, PT_ERR_ARG, "pt_export_dense: buffers too small");
    CK(cudaSetDevice(ctx->device));
    int rc = fold_epoch(ctx);
    if (rc) return rc;
    const int cap = ctx->sm_count * 8;
    export_nodes_kernel<<<grid_for(N > E ? N : E, 256, cap), 256, 0, ctx->stream>>>(ctx->T, (long long*)sums_dev, (long long*)stamps_dev, E);
    export_ovf_kernel<<<grid_for(ctx->ovf_cap, 256, cap), 256, 0, ctx->stream>>>(ctx->T, (long long*)sums_dev + 3 * N);
    CK(cudaGetLastError());
    ctx->launches += 2;
    return 0;
}

int pt_export_side(pt_ctx* ctx, uint64_t* novel_dev, uint64_t novel_rows, uint64_t* sparse_dev, uint64_t sparse_rows) {
    if (!ctx || !ctx->have_graph) return fail_msg(ctx, PT_ERR_STATE, "pt_export_side: no graph");
    CK(cudaSetDevice(ctx->device));
    const int cap = ctx->sm_count * 8;
    CK(cudaMemsetAsync(ctx->cursor, 0, 2 * sizeof(unsigned long long), ctx->stream));
    if (novel_rows && novel_dev)
        compact_side_kernel<<<grid_for(ctx->novel_cap, 256, cap), 256, 0, ctx->stream>>>(
            ctx->T.novel, ctx->novel_cap, (unsigned long long*)novel_dev, novel_rows, ctx->cursor);
    if (sparse_rows && sparse_dev)
        compact_side_kernel<<<grid_for(ctx->sparse_cap, 256, cap), 256, 0, ctx->stream>>>(
            ctx->T.sparse, ctx->sparse_cap, (unsigned long long*)sparse_dev, sparse_rows, ctx->cursor + 1);
    CK(cudaGetLastError());
    ctx->launches += 2;
    return 0;
}

int pt_timer_start(pt_ctx* ctx) {
    if (!ctx) return PT_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    CK(cudaEventRecord(ctx->ev_t0, ctx->stream));
    return 0;
}
int pt_timer_stop(pt_ctx* ctx, float* ms) {
    if (!ctx || !ms) return PT_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    CK(cudaEventRecord(ctx->ev_t1, ctx->stream));
    CK(cudaEventSynchronize(ctx->ev_t1));
    CK(cudaEventElapsedTime(ms, ctx->ev_t0, ctx->ev_t1));
    return 0;
}

int pt_profile_enable(pt_ctx* ctx, int on) {
    if (!ctx) return PT_ERR_ARG;
    ctx->profile = on != 0;
    return 0;
}

int pt_kernel_time_split(pt_ctx* ctx, float* ms_fast, float* ms_slow, uint64_t* launches) {
    if (!ctx) return PT_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    float tf = 0, ts = 0;
    for (uint32_t k = 0; k + 2 < ctx->prof_n; k += 3) {
        float a = 0, b = 0;
        CK(cudaEventElapsedTime(&a, ctx->prof_ev[k], ctx->prof_ev[k + 1]));
        CK(cudaEventElapsedTime(&b, ctx->prof_ev[k + 1], ctx->prof_ev[k + 2]));
        tf += a;
        ts += b;
    }
    if (ms_fast) *ms_fast = tf;
    if (ms_slow) *ms_slow = ts;
    if (launches) *launches = ctx->prof_n / 3;
    ctx->prof_n = 0;
    return 0;
}

int pt_kernel_time(pt_ctx* ctx, float* ms_total, uint64_t* launches) {
    float a = 0, b = 0;
    const int rc = pt_kernel_time_split(ctx, &a, &b, launches);
    if (rc) return rc;
    if (ms_total) *ms_total = a + b;
    return 0;
}

int pt_debug_counters(pt_ctx* ctx, uint64_t* out, int n) {
    if (!ctx || !out || n < 0) return PT_ERR_ARG;
    if (!ctx->have_graph) return fail_msg(ctx, PT_ERR_STATE, "pt_debug_counters: no graph");
    CK(cudaSetDevice(ctx->device));
    unsigned long long sc[SC_COUNT];
    CK(cudaMemcpyAsync(sc, ctx->T.sc, sizeof sc, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    for (int k = 0; k < n; k++) out[k] = k < 32 ? sc[SC_WHY + k] : 0;      // 16 hand-over reasons, then 16 phase clocks
    return 0;
}

int pt_stats(pt_ctx* ctx, uint64_t* kernel_launches, uint64_t* deferred_lines, uint64_t* tiles) {
    if (!ctx) return PT_ERR_ARG;
    if (kernel_launches) *kernel_launches = ctx->launches;
    if (ctx->have_graph && (deferred_lines || tiles)) {
        CK(cudaSetDevice(ctx->device));
        unsigned long long sc[SC_COUNT];
        CK(cudaMemcpyAsync(sc, ctx->T.sc, sizeof sc, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        if (deferred_lines) *deferred_lines = sc[SC_DEFERRED_TOTAL];
        if (tiles) *tiles = sc[SC_TILES];
    } else {
        if (deferred_lines) *deferred_lines = 0;
        if (tiles) *tiles = 0;
    }
    return 0;
}

}  // extern "C"
