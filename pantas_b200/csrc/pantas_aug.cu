// pantas_aug.cu -- sm_100a kernels + C ABI (include/pantas_aug.h) for pantas `augment`.
//
// Replaces the per-line loop of the reference,
//   /root/reference/scripts/alignments_augmentation_from_gaf.py:138-371   (REF:n)
// Data flow on the device (DESIGN.md has the full picture):
//
//   GAF bytes in HBM --cp.async.bulk (TMA 1-D, L2 evict-first)--> a team's shared-memory tile
//     team_tiles.cuh   the fast path: persistent two-warp teams, ten per SM, phases over a tile of about 30 records, up to
//                      9 KiB (byte-parallel SWAR scan, records by role warps, one thread per path step for ids / fold / count)
//     line_core.cuh    exact thread-per-record path for every record the fast path declines
//                      (augment_deferred_kernel, bytes from global memory)
//     tables.cuh       NodeHot[idx] = 16 bytes per node (meta word + three counters): one 4-byte load and
//                      one 32-bit RED per path step, the table stays in the L2; 64-bit-key open-addressing
//                      tables for the remaining links (known: ovf, unknown: novel) and for deletion-derived
//                      IL/OL keys (sparse)
//   (aug_kernels.cuh holds the device code; this file is the host side.)
//
// No tensor cores: nothing here is a contraction.  No CPU fallback: every entry
// point fails if the device is not sm_100.
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <type_traits>

#include "../../include/pantas_aug.h"
#include "line_core.cuh"

namespace {

#include "aug_kernels.cuh"
#include "gfa_kernels.cuh"

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
    bool have_totals;            // an epoch has been folded into the 64-bit totals since the last reset
    uint64_t epoch_end;          // highest file offset seen in the open epoch
    uint64_t folds;
    uint8_t* stage[2];
    uint64_t stage_bytes;
    int64_t next_ticket;
    uint64_t launches;
    uint32_t geo;                // tile bytes of the fast path (PANTAS_TEAM_TILE)
    int teams_per_sm;
    uint32_t ablate;             // PANTAS_ABLATE (diagnostics)
    uint32_t tile_bytes_env;     // PANTAS_TILE_BYTES: fixed tile length (0: from the record length)
    double rec_len_hint;         // average bytes per GAF record seen so far (0: unknown)
    uint64_t bytes_since_reset;  // GAF bytes enqueued since the last reset
    bool l2_window;              // an access-policy window over the node table is set on `stream`
    size_t persist_max, window_max;
    bool profile;
    cudaEvent_t* prof_ev;        // triples
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
    return x >= 0 ? (uint32_t)x : dflt;
}
static int grid_for(uint64_t n, int threads, int cap_blocks) {
    uint64_t b = (n + threads - 1) / threads;
    if (b < 1) b = 1;
    if (b > (uint64_t)cap_blocks) b = cap_blocks;
    return (int)b;
}

static void free_graph_tables(pt_ctx* ctx) {
    Tables& T = ctx->T;
    cudaFree(T.nodes); cudaFree(T.st32); cudaFree(T.len_full); cudaFree(T.il_ex32); cudaFree(T.ol_ex32); cudaFree(T.ovf);
    cudaFree(T.ovf_edge); cudaFree(T.inl_edge); cudaFree(T.novel); cudaFree(T.sparse); cudaFree(T.novel_list); cudaFree(T.sparse_list); cudaFree(T.t64); cudaFree(T.il_ex64);
    cudaFree(T.ol_ex64); cudaFree(T.il_st64); cudaFree(T.ol_st64); cudaFree(T.rc64);
    T.nodes = NULL; T.st32 = NULL; T.len_full = NULL; T.il_ex32 = T.ol_ex32 = NULL; T.ovf = NULL; T.ovf_edge = T.inl_edge = NULL;
    T.novel = T.sparse = NULL; T.novel_list = T.sparse_list = NULL; T.t64 = T.il_ex64 = T.ol_ex64 = NULL; T.il_st64 = T.ol_st64 = NULL; T.rc64 = NULL;
}

// Keep the node table in the L2: persisting access-policy window on the context's stream (the GAF stream itself is
// loaded evict-first by the kernel).  Best effort: a device without the feature just runs without it.
static void set_l2_window(pt_ctx* ctx, bool on) {
    cudaStreamAttrValue v;
    memset(&v, 0, sizeof v);
    if (on && ctx->T.nodes && ctx->persist_max && ctx->window_max && env_u32("PANTAS_L2_PERSIST", 1)) {
        size_t bytes = (size_t)ctx->T.n_nodes * sizeof(NodeHot);
        size_t win = bytes < ctx->window_max ? bytes : ctx->window_max;
        size_t carve = win < ctx->persist_max ? win : ctx->persist_max;
        if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, carve) != cudaSuccess) { cudaGetLastError(); return; }
        v.accessPolicyWindow.base_ptr = ctx->T.nodes;
        v.accessPolicyWindow.num_bytes = win;
        v.accessPolicyWindow.hitRatio = win <= carve ? 1.0f : (float)((double)carve / (double)win);
        v.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        v.accessPolicyWindow.missProp = cudaAccessPropertyNormal;
        if (cudaStreamSetAttribute(ctx->stream, cudaStreamAttributeAccessPolicyWindow, &v) != cudaSuccess) { cudaGetLastError(); return; }
        ctx->l2_window = true;
    } else if (ctx->l2_window) {
        v.accessPolicyWindow.num_bytes = 0;
        cudaStreamSetAttribute(ctx->stream, cudaStreamAttributeAccessPolicyWindow, &v);
        cudaGetLastError();
        ctx->l2_window = false;
    }
}

// fold the open epoch's 32-bit state into the 64-bit totals (tables.cuh)
static int fold_epoch(pt_ctx* ctx) {
    if (!ctx->epoch_open) return 0;
    fold_epoch_kernel<<<grid_for(ctx->T.n_nodes, 256, ctx->sm_count * 8), 256, 0, ctx->stream>>>(ctx->T);
    if (cudaGetLastError() != cudaSuccess) return fail_msg(ctx, PT_ERR_CUDA, "fold_epoch_kernel launch failed");
    ctx->launches += 1;
    ctx->folds += 1;
    ctx->epoch_open = false;
    ctx->have_totals = true;
    return 0;
}

// the fast path's geometries: Geo<tile, look-ahead, step-list entries, teams per CTA, CTAs per SM>
typedef teamp::Geo<9216, 1024, 512, 10, 1> GeoP;     // production: one CTA of ten teams per SM, tiles of up to 9 KiB (measured best, profiles/r02_geometry_sweep.txt)
typedef teamp::Geo<8192, 1024, 512, 5, 2> GeoU;      // two CTAs of five teams per SM
typedef teamp::Geo<1024, 256, 96, 2, 1> GeoT;        // tests: many tile boundaries, records longer than the look-ahead

static const double RECORDS_PER_TILE = 30.0;     // measured optimum on two record lengths (274 B, 300 B): 29.9 - 30.9

template <class G>
static int launch_team(pt_ctx* ctx, ChunkArgs A, const Tables& T) {
    void (*kern)(ChunkArgs, Tables) = teamp::augment_team_kernel<G>;
    const size_t smem = (size_t)G::SMEM_BYTES * G::NT;
    const int threads = (int)teamp::THREADS * G::NT;
    if (ctx->teams_per_sm == 0) {
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int occ = 0;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, smem));
        if (occ < 1) return fail_msg(ctx, PT_ERR_ARG, "fast path does not fit shared memory");
        const uint32_t want = env_u32("PANTAS_CTAS_PER_SM", 0);
        if (want && (int)want < occ) occ = (int)want;
        ctx->teams_per_sm = occ;                           // CTAs of G::NT teams
    }
    // Tile length: `records` and `walk` give every record of a tile one lane of one warp, so the best tile carries just under
    // 32 records (measured: profiles/r02_geometry_sweep.txt).  The record length comes from the host buffer (pt_process_host)
    // or from the counters of what this context has processed so far; until either is known, 8 KiB.
    uint32_t tb = ctx->tile_bytes_env;
    if (!tb) tb = ctx->rec_len_hint > 0.0 ? (uint32_t)(ctx->rec_len_hint * RECORDS_PER_TILE) : 8192u;
    tb = (tb + 15u) & ~15u;
    if (tb > (uint32_t)G::TILE) tb = (uint32_t)G::TILE;
    if (tb < 1024u) tb = (uint32_t)G::TILE < 1024u ? (uint32_t)G::TILE : 1024u;
    A.tile_bytes = tb;
    const uint64_t nt = (A.nbytes + tb - 1) / tb;
    if (nt > 0xFFFFFFF0ull) return fail_msg(ctx, PT_ERR_ARG, "chunk too large");
    A.n_tiles = (uint32_t)nt;
    uint64_t g = (uint64_t)ctx->sm_count * ctx->teams_per_sm;
    if (g * G::NT > nt) g = (nt + G::NT - 1) / G::NT;
    if (g * G::NT > T.team_cap) g = T.team_cap / G::NT;
    kern<<<(unsigned)g, threads, smem, ctx->stream>>>(A, T);
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

static void destroy_ctx_objects(pt_ctx* ctx) {
    cudaFree(ctx->T.sc); cudaFree(ctx->T.deferred); cudaFree(ctx->T.team_tile);
    cudaFree(ctx->stage[0]); cudaFree(ctx->stage[1]);
    for (int k = 0; k < 2; k++) {
        if (ctx->ev_copied[k]) cudaEventDestroy(ctx->ev_copied[k]);
        if (ctx->ev_done[k]) cudaEventDestroy(ctx->ev_done[k]);
    }
    if (ctx->ev_t0) cudaEventDestroy(ctx->ev_t0);
    if (ctx->ev_t1) cudaEventDestroy(ctx->ev_t1);
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    for (uint32_t k = 0; k < ctx->prof_cap; k++) cudaEventDestroy(ctx->prof_ev[k]);
    free(ctx->prof_ev);
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
    ctx->persist_max = (size_t)prop.persistingL2CacheMaxSize;
    ctx->window_max = (size_t)prop.accessPolicyMaxWindowSize;
    if (cudaSetDevice(device) != cudaSuccess) { free(ctx); return PT_ERR_CUDA; }
    ctx->own_stream = true;
    ctx->T.team_cap = 16384;
    cudaError_t e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking);
    for (int k = 0; k < 2 && e == cudaSuccess; k++) {
        e = cudaEventCreateWithFlags(&ctx->ev_copied[k], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_done[k], cudaEventDisableTiming);
    }
    if (e == cudaSuccess) e = cudaEventCreate(&ctx->ev_t0);
    if (e == cudaSuccess) e = cudaEventCreate(&ctx->ev_t1);
    if (e == cudaSuccess) e = cudaMalloc(&ctx->T.sc, SC_COUNT * sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMalloc(&ctx->T.team_tile, ctx->T.team_cap * sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaMemset(ctx->T.team_tile, 0, ctx->T.team_cap * sizeof(uint32_t));
    if (e != cudaSuccess) {
        destroy_ctx_objects(ctx);
        free(ctx);
        return PT_ERR_CUDA;
    }
    ctx->stage_bytes = (uint64_t)env_u32("PANTAS_STAGE_MB", 64) << 20;
    ctx->geo = env_u32("PANTAS_TEAM_TILE", 8192);
    ctx->ablate = env_u32("PANTAS_ABLATE", 0);
    ctx->tile_bytes_env = env_u32("PANTAS_TILE_BYTES", 0);
    *out = ctx;
    return 0;
}

void pt_destroy(pt_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    set_l2_window(ctx, false);
    free_graph_tables(ctx);
    destroy_ctx_objects(ctx);
    free(ctx);
}

const char* pt_last_error(pt_ctx* ctx) { return ctx ? ctx->err : "null context"; }

int pt_set_stream(pt_ctx* ctx, void* cuda_stream) {
    if (!ctx) return PT_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    const bool had = ctx->l2_window;
    set_l2_window(ctx, false);
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    ctx->stream = (cudaStream_t)cuda_stream;
    ctx->own_stream = false;
    if (had) set_l2_window(ctx, true);
    return 0;
}

static int reset_counts_impl(pt_ctx* ctx) {
    Tables& T = ctx->T;
    const int cap = ctx->sm_count * 8;
    const uint64_t N = T.n_nodes, E = ctx->n_edges;
    reset_nodes_kernel<<<grid_for(N, 256, cap), 256, 0, ctx->stream>>>(T);
    CK(cudaMemsetAsync(T.il_ex32, 0, N * sizeof(int32_t), ctx->stream));
    CK(cudaMemsetAsync(T.ol_ex32, 0, N * sizeof(int32_t), ctx->stream));
    CK(cudaMemsetAsync(T.t64, 0, N * sizeof(long long), ctx->stream));
    CK(cudaMemsetAsync(T.il_ex64, 0, N * sizeof(long long), ctx->stream));
    CK(cudaMemsetAsync(T.ol_ex64, 0, N * sizeof(long long), ctx->stream));
    CK(cudaMemsetAsync(T.rc64, 0, (E ? E : 1) * sizeof(long long), ctx->stream));
    clear_ovf_kernel<<<grid_for(ctx->ovf_cap, 256, cap), 256, 0, ctx->stream>>>(T.ovf, T.ovf_edge, ctx->ovf_cap, 0);
    clear_side_kernel<<<grid_for(T.novel_mask + 1, 256, cap), 256, 0, ctx->stream>>>(T.novel, T.novel_mask + 1);
    clear_side_kernel<<<grid_for(T.sparse_mask + 1, 256, cap), 256, 0, ctx->stream>>>(T.sparse, T.sparse_mask + 1);
    reset_teams_kernel<<<4, 256, 0, ctx->stream>>>(T);
    unsigned long long sc[SC_COUNT];
    memset(sc, 0, sizeof sc);
    sc[SC_ERR] = ~0ull;
    CK(cudaMemcpyAsync(T.sc, sc, sizeof sc, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));       // sc[] is a stack buffer
    CK(cudaGetLastError());
    ctx->launches += 5;
    ctx->epoch_open = false;
    ctx->have_totals = false;
    ctx->bytes_since_reset = 0;
    ctx->epoch_end = 0;
    T.epoch_base = 0;
    return 0;
}

static int set_graph_impl(pt_ctx* ctx, const uint32_t* node_len, uint64_t n_nodes, uint32_t min_id, const uint64_t* edge_keys,
                          uint64_t n_edges, uint64_t novel_cap, uint64_t sparse_cap, uint32_t** d_len, uint64_t** d_keys,
                          unsigned long long** d_stats) {
    Tables& T = ctx->T;
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

    CK(cudaMalloc(&T.nodes, n_nodes * sizeof(NodeHot)));
    CK(cudaMalloc(&T.st32, n_nodes * sizeof(Stamp32)));
    CK(cudaMalloc(&T.len_full, n_nodes * sizeof(uint32_t)));
    CK(cudaMalloc(&T.il_ex32, n_nodes * sizeof(int32_t)));
    CK(cudaMalloc(&T.ol_ex32, n_nodes * sizeof(int32_t)));
    CK(cudaMalloc(&T.inl_edge, 2 * n_nodes * sizeof(uint32_t)));
    CK(cudaMalloc(&T.t64, n_nodes * sizeof(long long)));
    CK(cudaMalloc(&T.il_ex64, n_nodes * sizeof(long long)));
    CK(cudaMalloc(&T.ol_ex64, n_nodes * sizeof(long long)));
    CK(cudaMalloc(&T.il_st64, n_nodes * sizeof(unsigned long long)));
    CK(cudaMalloc(&T.ol_st64, n_nodes * sizeof(unsigned long long)));
    CK(cudaMalloc(&T.rc64, (n_edges ? n_edges : 1) * sizeof(long long)));
    CK(cudaMalloc(&T.novel, ctx->novel_cap * sizeof(SideSlot)));
    CK(cudaMalloc(&T.sparse, ctx->sparse_cap * sizeof(SideSlot)));
    CK(cudaMalloc(&T.novel_list, ctx->novel_cap * sizeof(uint32_t)));
    CK(cudaMalloc(&T.sparse_list, ctx->sparse_cap * sizeof(uint32_t)));

    CK(cudaMalloc(d_len, n_nodes * sizeof(uint32_t)));
    CK(cudaMalloc(d_keys, (n_edges ? n_edges : 1) * sizeof(uint64_t)));
    CK(cudaMalloc(d_stats, 2 * sizeof(unsigned long long)));   // [0] bad / duplicate keys, [1] links that are not inline
    CK(cudaMemcpyAsync(*d_len, node_len, n_nodes * sizeof(uint32_t), cudaMemcpyDefault, ctx->stream));
    if (n_edges) CK(cudaMemcpyAsync(*d_keys, edge_keys, n_edges * sizeof(uint64_t), cudaMemcpyDefault, ctx->stream));
    CK(cudaMemsetAsync(*d_stats, 0, 2 * sizeof(unsigned long long), ctx->stream));
    const int cap = ctx->sm_count * 8;
    init_nodes_kernel<<<grid_for(n_nodes, 256, cap), 256, 0, ctx->stream>>>(T, *d_len);
    if (n_edges) inline_edges_kernel<<<grid_for(n_edges, 256, cap), 256, 0, ctx->stream>>>(T, *d_keys, n_edges, *d_stats);
    unsigned long long stats[2] = {0, 0};
    CK(cudaMemcpyAsync(stats, *d_stats, sizeof stats, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaGetLastError());
    // known links that are not inline: open addressing at load <= 0.5
    ctx->ovf_cap = pow2_at_least(stats[1] * 2 + 1024);
    T.ovf_mask = ctx->ovf_cap - 1;
    CK(cudaMalloc(&T.ovf, ctx->ovf_cap * sizeof(OvfSlot)));
    CK(cudaMalloc(&T.ovf_edge, ctx->ovf_cap * sizeof(uint32_t)));
    clear_ovf_kernel<<<grid_for(ctx->ovf_cap, 256, cap), 256, 0, ctx->stream>>>(T.ovf, T.ovf_edge, ctx->ovf_cap, 1);
    if (n_edges) ovf_edges_kernel<<<grid_for(n_edges, 256, cap), 256, 0, ctx->stream>>>(T, *d_keys, n_edges, *d_stats);
    CK(cudaMemcpyAsync(stats, *d_stats, sizeof stats, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaGetLastError());
    ctx->launches += 4;
    if (stats[0]) return fail_msg(ctx, PT_ERR_ARG, "pt_set_graph: edge_keys hold duplicates or out-of-range node indices");
    return 0;
}

int pt_set_graph(pt_ctx* ctx, const uint32_t* node_len, uint64_t n_nodes, uint32_t min_id, const uint64_t* edge_keys,
                 uint64_t n_edges, uint64_t novel_cap, uint64_t sparse_cap) {
    if (!ctx || !node_len || n_nodes == 0 || (n_edges && !edge_keys)) return fail_msg(ctx, PT_ERR_ARG, "pt_set_graph: bad argument");
    if (n_nodes >= 0xFFFFFFFFull || n_edges >= 0xFFFFFFFFull) return fail_msg(ctx, PT_ERR_ARG, "pt_set_graph: more than 2^32-2 nodes or links");
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    set_l2_window(ctx, false);
    free_graph_tables(ctx);
    ctx->have_graph = false;
    uint32_t* d_len = NULL;
    uint64_t* d_keys = NULL;
    unsigned long long* d_stats = NULL;
    const int rc = set_graph_impl(ctx, node_len, n_nodes, min_id, edge_keys, n_edges, novel_cap, sparse_cap, &d_len, &d_keys, &d_stats);
    cudaFree(d_len); cudaFree(d_keys); cudaFree(d_stats);          // temporaries: on every path
    if (rc) {
        cudaStreamSynchronize(ctx->stream);
        free_graph_tables(ctx);
        return rc;
    }
    ctx->have_graph = true;
    set_l2_window(ctx, true);
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
    // 32-bit stamps / counters are relative to an epoch of < 4 GiB of GAF: fold when this chunk does not fit.  A chunk
    // that starts before the end of what the epoch has seen also folds: "settled" stamps assume offsets only grow.
    if (ctx->epoch_open && (file_offset < ctx->epoch_end || file_offset + nbytes - (uint64_t)T.epoch_base > EPOCH_SPAN)) {
        int rc = fold_epoch(ctx);
        if (rc) return rc;
    }
    if (!ctx->epoch_open) {
        T.epoch_base = (int64_t)file_offset;
        ctx->epoch_open = true;
        ctx->epoch_end = file_offset;
    }
    if (file_offset + nbytes > ctx->epoch_end) ctx->epoch_end = file_offset + nbytes;
    ctx->bytes_since_reset += nbytes;
    // second-pass list: records the fast path hands over.  The shortest record the reference accepts has 12 one-byte
    // columns, 11 separators and a line break (24 bytes), so this holds every record of the chunk.
    const uint64_t want = nbytes / 24 + 4096;
    if (want > T.deferred_cap) {
        CK(cudaStreamSynchronize(ctx->stream));
        cudaFree(T.deferred);
        T.deferred = NULL;
        T.deferred_cap = 0;
        CK(cudaMalloc(&T.deferred, want * sizeof(uint32_t)));
        T.deferred_cap = want;
    }
    ChunkArgs A;
    memset(&A, 0, sizeof A);
    A.gaf = gaf_dev;
    A.nbytes = nbytes;
    A.file_off = (int64_t)file_offset;
    A.thr = thr;
    A.ablate = ctx->ablate;
    A.loose = env_u32("PANTAS_LOOSE", 1);
    {   // measured: -4 % kernel time, -25 % DRAM reads (PANTAS_STREAM_HINT=0 switches it off)
        const char* h = getenv("PANTAS_STREAM_HINT");
        A.stream_hint = (h && h[0] == '0') ? 0u : 1u;
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
    int rc;
    switch (ctx->geo) {
        case 1024: rc = launch_team<GeoT>(ctx, A, T); break;
        case 8193: rc = launch_team<GeoU>(ctx, A, T); break;
        default: rc = launch_team<GeoP>(ctx, A, T); break;
    }
    if (rc) return rc;
    if (ctx->profile) CK(cudaEventRecord(ctx->prof_ev[ctx->prof_n + 1], ctx->stream));
    augment_deferred_kernel<<<ctx->sm_count * 8, 128, 0, ctx->stream>>>(A, T);
    end_chunk_kernel<<<4, 256, 0, ctx->stream>>>(T);
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
    if (ctx->rec_len_hint <= 0.0 && nbytes) {                   // record length from the head of the host buffer
        const uint64_t peek = nbytes < (256u << 10) ? nbytes : (256u << 10);
        uint64_t lines = 0, last = 0;
        for (const uint8_t* p = gaf_host; (p = (const uint8_t*)memchr(p, '\n', (size_t)(gaf_host + peek - p))) != NULL; p++) {
            lines++;
            last = (uint64_t)(p - gaf_host) + 1;
        }
        if (lines >= 8) ctx->rec_len_hint = (double)last / (double)lines;
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
    // ev_copied[k] belongs to the LATEST ticket of parity k; the copy stream is in order, so its completion
    // implies the completion of every earlier copy into the same stage
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
    if (sc[SC_LINES] >= 1000ull && ctx->bytes_since_reset)      // average record length: sizes the tiles of later chunks
        ctx->rec_len_hint = (double)ctx->bytes_since_reset / (double)sc[SC_LINES];
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
    // (no fold: the export kernels add the open epoch's 32-bit state to the 64-bit totals on the fly, the tables stay as they are)
    const uint32_t flags = (ctx->have_totals ? EXP_TOTALS : 0u) | (ctx->epoch_open ? EXP_LIVE : 0u);
    const int cap = ctx->sm_count * 8;
    export_nodes_kernel<<<grid_for(N > E ? N : E, 256, cap), 256, 0, ctx->stream>>>(ctx->T, (long long*)sums_dev, (long long*)stamps_dev, E, flags);
    export_inline_kernel<<<grid_for(N, 256, cap), 256, 0, ctx->stream>>>(ctx->T, (long long*)sums_dev, flags);
    export_ovf_kernel<<<grid_for(ctx->ovf_cap, 256, cap), 256, 0, ctx->stream>>>(ctx->T, (long long*)sums_dev);
    export_novel_ends_kernel<<<ctx->sm_count, 256, 0, ctx->stream>>>(ctx->T, (long long*)sums_dev);
    CK(cudaGetLastError());
    ctx->launches += 4;
    return 0;
}

int pt_export_side(pt_ctx* ctx, uint64_t* novel_dev, uint64_t novel_rows, uint64_t* sparse_dev, uint64_t sparse_rows) {
    if (!ctx || !ctx->have_graph) return fail_msg(ctx, PT_ERR_STATE, "pt_export_side: no graph");
    CK(cudaSetDevice(ctx->device));
    const int cap = ctx->sm_count * 8;
    if (novel_rows && novel_dev) {
        compact_side_kernel<<<grid_for(novel_rows, 256, cap), 256, 0, ctx->stream>>>(
            ctx->T.novel, ctx->T.novel_list, novel_rows, (unsigned long long*)novel_dev, novel_rows);
        ctx->launches += 1;
    }
    if (sparse_rows && sparse_dev) {
        compact_side_kernel<<<grid_for(sparse_rows, 256, cap), 256, 0, ctx->stream>>>(
            ctx->T.sparse, ctx->T.sparse_list, sparse_rows, (unsigned long long*)sparse_dev, sparse_rows);
        ctx->launches += 1;
    }
    CK(cudaGetLastError());
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
    for (int k = 0; k < n; k++) out[k] = k < 16 ? sc[SC_WHY + k] : 0;      // 16 hand-over reasons
    return 0;
}

// ---------------------------------------------------------------- the two GFA passes (gfa_kernels.cuh)

int pt_gfa_parse(pt_ctx* ctx, const uint8_t* gfa_dev, const int64_t* start_dev, const int64_t* end_dev, uint64_t n_lines,
                 uint32_t* a_rel_dev, uint32_t* slen_dev, uint32_t* kind_dev, uint32_t* v1_dev, uint32_t* v2_dev, uint64_t* err_dev) {
    if (!ctx || !gfa_dev || !start_dev || !end_dev || !a_rel_dev || !slen_dev || !kind_dev || !v1_dev || !v2_dev || !err_dev)
        return fail_msg(ctx, PT_ERR_ARG, "pt_gfa_parse: null argument");
    CK(cudaSetDevice(ctx->device));
    if (n_lines == 0) return 0;
    gfa::LineArrays L;
    L.start = (const long long*)start_dev;
    L.end = (const long long*)end_dev;
    L.a_rel = a_rel_dev;
    L.slen = slen_dev;
    L.kind = kind_dev;
    L.v1 = v1_dev;
    L.v2 = v2_dev;
    gfa::gfa_parse_kernel<<<grid_for(n_lines, 256, ctx->sm_count * 16), 256, 0, ctx->stream>>>(gfa_dev, L, n_lines, (unsigned long long*)err_dev);
    CK(cudaGetLastError());
    ctx->launches += 1;
    return 0;
}

static gfa::WriterArgs writer_args(const pt_gfa_writer* w) {
    gfa::WriterArgs W;
    W.s = w->gfa;
    W.start = (const long long*)w->start;
    W.a_rel = w->a_rel;
    W.slen = w->slen;
    W.kind = w->kind;
    W.v1 = w->v1;
    W.link_edge = w->link_edge;
    W.node_len = w->node_len;
    W.sums = (const long long*)w->sums;
    W.sp_slot = w->sp_slot;
    W.sp_off = (const long long*)w->sp_off;
    W.sp_text = w->sp_text;
    W.n_nodes = w->n_nodes;
    W.n_lines = w->n_lines;
    W.min_id = w->min_id;
    return W;
}

int pt_gfa_measure(pt_ctx* ctx, const pt_gfa_writer* w, int64_t* out_len_dev, uint64_t* err_dev) {
    if (!ctx || !w || !out_len_dev || !err_dev) return fail_msg(ctx, PT_ERR_ARG, "pt_gfa_measure: null argument");
    CK(cudaSetDevice(ctx->device));
    if (w->n_lines == 0) return 0;
    gfa::gfa_measure_kernel<<<grid_for(w->n_lines, 256, ctx->sm_count * 16), 256, 0, ctx->stream>>>(writer_args(w), (long long*)out_len_dev,
                                                                                                  (unsigned long long*)err_dev);
    CK(cudaGetLastError());
    ctx->launches += 1;
    return 0;
}

int pt_gfa_format(pt_ctx* ctx, const pt_gfa_writer* w, const int64_t* out_off_dev, uint8_t* out_dev, uint64_t* err_dev) {
    if (!ctx || !w || !out_off_dev || !out_dev || !err_dev) return fail_msg(ctx, PT_ERR_ARG, "pt_gfa_format: null argument");
    CK(cudaSetDevice(ctx->device));
    if (w->n_lines == 0) return 0;
    gfa::gfa_format_kernel<<<grid_for(w->n_lines, 256, ctx->sm_count * 16), 256, 0, ctx->stream>>>(writer_args(w), (const long long*)out_off_dev, out_dev,
                                                                                                 (unsigned long long*)err_dev);
    CK(cudaGetLastError());
    ctx->launches += 1;
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
