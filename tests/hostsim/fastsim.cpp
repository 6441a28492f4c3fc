// tests/hostsim/fastsim.cpp -- TEST HARNESS, NOT PRODUCT CODE.
//
// Runs the REAL device code of the augment pass (pantas_b200/csrc/aug_kernels.cuh: table build, the
// fast path of team_tiles.cuh, the exact per-record kernel, epoch fold, export) on the CPU through the
// fibre emulator of cuda_emu.h, following the launch sequence of pantas_aug.cu (pt_set_graph,
// pt_reset_counts, pt_process_chunk, pt_export_dense / pt_export_side).  The build container has no
// GPU: this is how CPU-only tests fuzz the fast path against the oracle.  Nothing under pantas_b200/
// loads this.
#include "cuda_emu.h"

#include <algorithm>

#include "../../pantas_b200/csrc/line_core.cuh"

namespace {
#include "../../pantas_b200/csrc/aug_kernels.cuh"

uint64_t pow2_at_least(uint64_t v) {
    uint64_t p = 1;
    while (p < v) p <<= 1;
    return p;
}
template <class T> T* zalloc(uint64_t n) { return (T*)calloc(n ? n : 1, sizeof(T)); }

template <class G>
void run_fast(unsigned grid, ChunkArgs A, const Tables& T) {
    // (the product picks the tile length from the record length; every fifth run here uses a shorter tile than the layout holds)
    A.tile_bytes = (uint32_t)G::TILE;
    if ((grid + A.loose) % 5u == 0u) A.tile_bytes = (uint32_t)(G::TILE - G::TILE / 8) & ~15u;
    if (getenv("FASTSIM_TILE_BYTES")) A.tile_bytes = (uint32_t)atoi(getenv("FASTSIM_TILE_BYTES"));
    A.n_tiles = (uint32_t)((A.nbytes + A.tile_bytes - 1) / A.tile_bytes);
    if (grid > A.n_tiles) grid = A.n_tiles;
    const unsigned ctas = (grid + G::NT - 1) / G::NT;           // `grid` counts teams
    emu::launch(ctas, teamp::THREADS * G::NT, (size_t)G::SMEM_BYTES * G::NT, [&] { teamp::augment_team_kernel<G>(A, T); });
}

}  // namespace

extern "C" {

struct fastsim_result {
    int64_t* sums;
    int64_t* stamps;
    uint64_t* novel;
    uint64_t n_novel;
    uint64_t* sparse;
    uint64_t n_sparse;
    uint64_t err_offset;
    int err_code;
    uint64_t n_deferred;
    uint64_t why[16];
};

// geo: 0 = tiny test tiles (1 KiB), 1 = 4 KiB tiles, 2 = the production geometry
int fastsim_run(const uint8_t* gaf, uint64_t nbytes, uint64_t file_off, int64_t thr, const uint32_t* node_len,
                uint64_t n_nodes, uint32_t min_id, const uint64_t* edge_keys, uint64_t n_edges, int geo, uint32_t grid,
                fastsim_result* out) {
    const uint64_t N = n_nodes, E = n_edges;
    Tables T;
    memset(&T, 0, sizeof T);
    T.n_nodes = N;
    T.min_id = min_id;
    const uint64_t novel_cap = 1u << 12, sparse_cap = 1u << 12;
    T.novel_mask = novel_cap - 1;
    T.sparse_mask = sparse_cap - 1;
    T.nodes = zalloc<NodeHot>(N);
    T.st32 = zalloc<Stamp32>(N);
    T.len_full = zalloc<uint32_t>(N);
    T.il_ex32 = zalloc<int32_t>(N);
    T.ol_ex32 = zalloc<int32_t>(N);
    T.inl_edge = zalloc<uint32_t>(2 * N);
    T.t64 = zalloc<long long>(N);
    T.il_ex64 = zalloc<long long>(N);
    T.ol_ex64 = zalloc<long long>(N);
    T.team_cap = 64;
    T.team_tile = zalloc<uint32_t>(T.team_cap);
    T.il_st64 = zalloc<unsigned long long>(N);
    T.ol_st64 = zalloc<unsigned long long>(N);
    T.rc64 = zalloc<long long>(E);
    T.novel = zalloc<SideSlot>(novel_cap);
    T.sparse = zalloc<SideSlot>(sparse_cap);
    T.novel_list = zalloc<uint32_t>(novel_cap);
    T.sparse_list = zalloc<uint32_t>(sparse_cap);
    T.sc = zalloc<unsigned long long>(SC_COUNT);
    T.deferred_cap = nbytes / 2 + 4096;
    T.deferred = zalloc<uint32_t>(T.deferred_cap);
    unsigned long long stats[2] = {0, 0};
    const unsigned KG = 2, KB = 64;

    // ---- pt_set_graph
    emu::launch(KG, KB, 0, [&] { init_nodes_kernel(T, node_len); });
    if (E) emu::launch(KG, KB, 0, [&] { inline_edges_kernel(T, edge_keys, E, stats); });
    const uint64_t ovf_cap = pow2_at_least(stats[1] * 2 + 16);
    T.ovf_mask = ovf_cap - 1;
    T.ovf = zalloc<OvfSlot>(ovf_cap);
    T.ovf_edge = zalloc<uint32_t>(ovf_cap);
    emu::launch(KG, KB, 0, [&] { clear_ovf_kernel(T.ovf, T.ovf_edge, ovf_cap, 1); });
    if (E) emu::launch(KG, KB, 0, [&] { ovf_edges_kernel(T, edge_keys, E, stats); });
    // ---- pt_reset_counts
    emu::launch(KG, KB, 0, [&] { reset_nodes_kernel(T); });
    emu::launch(KG, KB, 0, [&] { clear_side_kernel(T.novel, novel_cap); });
    emu::launch(KG, KB, 0, [&] { clear_side_kernel(T.sparse, sparse_cap); });
    T.sc[SC_ERR] = ~0ull;
    T.epoch_base = (int64_t)file_off;

    // ---- pt_process_chunk: the chunk must be 16-byte aligned and readable up to the next multiple of 16
    uint8_t* raw = (uint8_t*)malloc(nbytes + 96);
    uint8_t* dev = (uint8_t*)(((uintptr_t)raw + 15) & ~(uintptr_t)15);
    memset(dev, 0xEE, nbytes + 64);          // garbage after the data, like a reused staging buffer
    if (nbytes) memcpy(dev, gaf, nbytes);
    ChunkArgs A;
    memset(&A, 0, sizeof A);
    A.loose = (geo + grid) & 1u;      // both barrier modes get exercised
    if (getenv("FASTSIM_ABLATE")) A.ablate = (uint32_t)atoi(getenv("FASTSIM_ABLATE"));
    A.thr = thr;
    auto run_chunk = [&](const uint8_t* p, uint64_t n, uint64_t off) {
        if (!n) return;
        A.gaf = p;
        A.nbytes = n;
        A.file_off = (int64_t)off;
        typedef teamp::Geo<1024, 256, 96, 2, 1> G0;
        typedef teamp::Geo<4096, 512, 256, 3, 1> G1;
        typedef teamp::Geo<9216, 1024, 512, 10, 1> G2;
        if (geo == 0) run_fast<G0>(grid, A, T);
        else if (geo == 1) run_fast<G1>(grid, A, T);
        else run_fast<G2>(grid, A, T);
        emu::launch(2, 128, 0, [&] { augment_deferred_kernel(A, T); });
        emu::launch(2, 64, 0, [&] { end_chunk_kernel(T); });
    };
    // every third run is two chunks with an epoch fold between them (what launch_chunk does every < 4 GiB of GAF): the
    // export below then has to add the open epoch's 32-bit state to the folded 64-bit totals
    uint64_t mid = 0;
    if ((geo + grid) % 3u == 0u && nbytes > 2) {
        const void* nl = memchr(gaf + nbytes / 2, '\n', nbytes - nbytes / 2 - 1);
        if (nl) mid = (uint64_t)((const uint8_t*)nl - gaf) + 1;
    }
    bool folded = false;
    uint8_t* raw2 = NULL;
    if (mid) {
        run_chunk(dev, mid, file_off);
        emu::launch(KG, KB, 0, [&] { fold_epoch_kernel(T); });
        folded = true;
        T.epoch_base = (int64_t)(file_off + mid);
        raw2 = (uint8_t*)malloc(nbytes - mid + 96);
        uint8_t* dev2 = (uint8_t*)(((uintptr_t)raw2 + 15) & ~(uintptr_t)15);
        memset(dev2, 0xEE, nbytes - mid + 64);
        memcpy(dev2, gaf + mid, nbytes - mid);
        run_chunk(dev2, nbytes - mid, file_off + mid);
    } else {
        run_chunk(dev, nbytes, file_off);
    }
    // ---- pt_export_dense / pt_export_side
    // (every other run folds the epoch first, like a chunk that opens a new epoch would: the export then reads the 64-bit
    // totals instead of the open epoch's 32-bit state)
    uint32_t flags = EXP_LIVE | (folded ? EXP_TOTALS : 0u);
    if ((geo + grid) & 4u) {
        emu::launch(KG, KB, 0, [&] { fold_epoch_kernel(T); });
        flags = EXP_TOTALS;
    }
    out->sums = (int64_t*)calloc(3 * N + E + 4, sizeof(int64_t));
    out->stamps = (int64_t*)calloc(2 * N + 1, sizeof(int64_t));
    emu::launch(KG, KB, 0, [&] { export_nodes_kernel(T, (long long*)out->sums, (long long*)out->stamps, E, flags); });
    emu::launch(KG, KB, 0, [&] { export_inline_kernel(T, (long long*)out->sums, flags); });
    emu::launch(KG, KB, 0, [&] { export_ovf_kernel(T, (long long*)out->sums); });
    emu::launch(KG, KB, 0, [&] { export_novel_ends_kernel(T, (long long*)out->sums); });
    out->n_novel = T.sc[SC_NOVEL_USED] < novel_cap ? T.sc[SC_NOVEL_USED] : novel_cap;
    out->n_sparse = T.sc[SC_SPARSE_USED] < sparse_cap ? T.sc[SC_SPARSE_USED] : sparse_cap;
    out->novel = (uint64_t*)calloc(3 * novel_cap + 1, sizeof(uint64_t));
    out->sparse = (uint64_t*)calloc(3 * sparse_cap + 1, sizeof(uint64_t));
    emu::launch(KG, KB, 0, [&] { compact_side_kernel(T.novel, T.novel_list, out->n_novel, (unsigned long long*)out->novel, novel_cap); });
    emu::launch(KG, KB, 0, [&] { compact_side_kernel(T.sparse, T.sparse_list, out->n_sparse, (unsigned long long*)out->sparse, sparse_cap); });
    out->n_deferred = T.sc[SC_DEFERRED_TOTAL];
    for (int k = 0; k < 16; k++) out->why[k] = T.sc[SC_WHY + k];
    if (T.sc[SC_ERR] != ~0ull) {
        out->err_offset = T.sc[SC_ERR] >> 8;
        out->err_code = (int)(T.sc[SC_ERR] & 0xff);
    } else {
        out->err_offset = 0;
        out->err_code = 0;
    }
    free(raw);
    free(raw2);
    free(T.nodes); free(T.st32); free(T.len_full); free(T.il_ex32); free(T.ol_ex32); free(T.inl_edge); free(T.t64); free(T.il_ex64);
    free(T.ol_ex64); free(T.il_st64); free(T.ol_st64); free(T.rc64); free(T.novel); free(T.sparse); free(T.novel_list); free(T.sparse_list); free(T.sc); free(T.deferred);
    free(T.ovf); free(T.ovf_edge); free(T.team_tile);
    return 0;
}

void fastsim_free(fastsim_result* r) {
    free(r->sums);
    free(r->stamps);
    free(r->novel);
    free(r->sparse);
    memset(r, 0, sizeof *r);
}

}  // extern "C"
