// tests/hostsim/hostsim.cpp -- TEST HARNESS, NOT PRODUCT CODE.
//
// Compiles pantas_b200/csrc/line_core.cuh (the per-record logic the CUDA
// kernels run) with g++ and drives it with an std::map based Sink, emulating
// the kernel's tiling (owned line starts per tile, bounded look-ahead window,
// deferral of records that run past the window).  The build container has no
// GPU; this lets tests fuzz the record logic against the oracle there, and
// gives the world_size-2 gloo tests per-shard results in the exact layout the
// device library exports.  Nothing under pantas_b200/ loads this.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <map>
#include <unordered_map>
#include <vector>

#include "../../pantas_b200/csrc/line_core.cuh"

namespace {

struct Side { int64_t count; uint64_t stamp; };

struct HostSink {
    const uint32_t* node_len;
    uint64_t n_nodes;
    uint32_t min_id;
    std::vector<int64_t> nc, il_adj, ol_adj, rc;
    std::vector<uint64_t> il_stamp, ol_stamp;
    std::unordered_map<uint64_t, uint64_t> known;   // key -> edge index
    std::map<uint64_t, Side> novel_tab, sparse_tab;
    int64_t rej = 0;
    uint64_t err = UINT64_MAX;                      // (offset << 8) | code, min wins

    struct Stamps { uint64_t il, ol; };
    struct EdgePf { int unused; };
    bool id_to_idx(uint64_t id, uint32_t& idx) {
        if (id < min_id) return false;
        uint64_t d = id - min_id;
        if (d >= n_nodes) return false;
        idx = (uint32_t)d;
        return true;
    }
    uint32_t load_len(uint32_t idx) { return node_len[idx]; }
    Stamps load_stamps(uint32_t idx) { return Stamps{il_stamp[idx], ol_stamp[idx]}; }
    void edge_pf_init(EdgePf& pf) { pf.unused = 0; }
    void prefetch_edge(EdgePf&, uint32_t) {}
    void count_node(uint32_t idx) { nc[idx]++; }
    void dense(uint32_t idx, int64_t il, int64_t ol, uint64_t stamp, const Stamps&, bool, bool) {
        il_adj[idx] += il - 1;
        ol_adj[idx] += ol - 1;
        if (il > 0) il_stamp[idx] = std::min(il_stamp[idx], stamp);
        if (ol > 0) ol_stamp[idx] = std::min(ol_stamp[idx], stamp);
    }
    void sparse_ev(uint64_t key, uint64_t stamp, std::map<uint64_t, Side>& m) {
        auto it = m.find(key);
        if (it == m.end()) m[key] = Side{1, stamp};
        else { it->second.count++; it->second.stamp = std::min(it->second.stamp, stamp); }
    }
    void sparse(uint32_t idx, int dir, int64_t pos, uint64_t stamp) {
        const int64_t bias = 1ll << 30;
        if (pos < -bias || pos >= bias) { error(pt::PT_U_POSITION, (int64_t)(stamp >> 2)); return; }
        uint64_t key = ((uint64_t)idx << 32) | ((uint64_t)dir << 31) | (uint64_t)(pos + bias);
        sparse_ev(key, stamp, sparse_tab);
    }
    void edge(uint32_t a, uint32_t b, uint64_t stamp, const EdgePf&) {
        uint64_t key = ((uint64_t)a << 32) | b;
        auto it = known.find(key);
        if (it != known.end()) rc[it->second]++;
        else sparse_ev(key, stamp, novel_tab);
    }
    void reject() { rej++; }
    void error(int code, int64_t off) { err = std::min(err, ((uint64_t)off << 8) | (uint64_t)code); }
};

}  // namespace

extern "C" {

// Results use the same flat layout as pt_export_dense / pt_export_side
// (include/pantas_aug.h): sums = [nc | il_adj | ol_adj | rc | rej, n_lines, 0, 0],
// stamps = [il_stamp | ol_stamp] (INT64_MAX when never touched),
// side tables = rows of {key, count, stamp}.
struct hostsim_result {
    int64_t* sums;
    int64_t* stamps;
    uint64_t* novel;
    uint64_t n_novel;
    uint64_t* sparse;
    uint64_t n_sparse;
    uint64_t err_offset;
    int err_code;
    uint64_t n_deferred;
};

int hostsim_run(const uint8_t* gaf, uint64_t nbytes, uint64_t file_off, int64_t thr, const uint32_t* node_len,
                uint64_t n_nodes, uint32_t min_id, const uint64_t* edge_keys, uint64_t n_edges, uint32_t tile,
                uint32_t over, hostsim_result* out) {
    HostSink sink;
    sink.node_len = node_len;
    sink.n_nodes = n_nodes;
    sink.min_id = min_id;
    sink.nc.assign(n_nodes, 0);
    sink.il_adj.assign(n_nodes, 0);
    sink.ol_adj.assign(n_nodes, 0);
    sink.rc.assign(n_edges, 0);
    sink.il_stamp.assign(n_nodes, (uint64_t)INT64_MAX);
    sink.ol_stamp.assign(n_nodes, (uint64_t)INT64_MAX);
    for (uint64_t e = 0; e < n_edges; e++) sink.known.emplace(edge_keys[e], e);

    // the device buffers are readable up to the next multiple of 16 past the data; give the
    // host copy the same slack (SWAR loads touch whole aligned words)
    std::vector<uint8_t> padded(nbytes + 64, (uint8_t)'\n');
    if (nbytes) memcpy(padded.data(), gaf, nbytes);
    gaf = padded.data();
    uint64_t n_lines = 0, n_deferred = 0;
    // cooperative-scan equivalents: non-ASCII, bare CR
    for (uint64_t i = 0; i < nbytes; i++) {
        if (gaf[i] >= 0x80) sink.error(pt::PT_U_NON_ASCII, (int64_t)(file_off + i));
        if (gaf[i] == '\r' && i + 1 < nbytes && gaf[i + 1] != '\n') sink.error(pt::PT_U_BARE_CR, (int64_t)(file_off + i));
    }
    if (tile == 0) tile = 1u << 30;
    for (uint64_t t0 = 0; t0 < nbytes; t0 += tile) {
        const uint64_t t1 = std::min<uint64_t>(t0 + tile, nbytes);
        const uint64_t lo = t0 >= 16 ? t0 - 16 : 0;
        const uint64_t hi = std::min<uint64_t>(t0 + tile + over, nbytes);
        pt::LineCtx cx;
        cx.s = gaf + lo;
        cx.lim = (int)(hi - lo);
        cx.lim_final = (hi == nbytes);
        cx.base_off = (int64_t)(file_off + lo);
        for (uint64_t a = t0; a < t1; a++) {
            if (!(a == 0 || gaf[a - 1] == '\n')) continue;
            n_lines++;
            int r = pt::process_line(cx, (int)(a - lo), thr, sink);
            if (r == pt::LINE_DEFER) {
                n_deferred++;
                pt::LineCtx full;
                full.s = gaf + a;
                full.lim = (int)std::min<uint64_t>(nbytes - a, 0x7fffffff);
                full.lim_final = true;
                full.base_off = (int64_t)(file_off + a);
                pt::process_line(full, 0, thr, sink);
            }
        }
    }

    const uint64_t N = n_nodes, E = n_edges;
    out->sums = (int64_t*)calloc(3 * N + E + 4, sizeof(int64_t));
    out->stamps = (int64_t*)calloc(2 * N + 1, sizeof(int64_t));
    for (uint64_t i = 0; i < N; i++) {
        out->sums[i] = sink.nc[i];
        out->sums[N + i] = sink.il_adj[i];
        out->sums[2 * N + i] = sink.ol_adj[i];
        out->stamps[i] = (int64_t)sink.il_stamp[i];
        out->stamps[N + i] = (int64_t)sink.ol_stamp[i];
    }
    for (uint64_t e = 0; e < E; e++) out->sums[3 * N + e] = sink.rc[e];
    out->sums[3 * N + E] = sink.rej;
    out->sums[3 * N + E + 1] = (int64_t)n_lines;
    auto dump = [](const std::map<uint64_t, Side>& m, uint64_t*& arr, uint64_t& n) {
        n = m.size();
        arr = (uint64_t*)calloc(3 * n + 1, sizeof(uint64_t));
        uint64_t k = 0;
        for (auto& kv : m) {
            arr[3 * k] = kv.first;
            arr[3 * k + 1] = (uint64_t)kv.second.count;
            arr[3 * k + 2] = kv.second.stamp;
            k++;
        }
    };
    dump(sink.novel_tab, out->novel, out->n_novel);
    dump(sink.sparse_tab, out->sparse, out->n_sparse);
    out->n_deferred = n_deferred;
    if (sink.err != UINT64_MAX) {
        out->err_offset = sink.err >> 8;
        out->err_code = (int)(sink.err & 0xff);
    } else {
        out->err_offset = 0;
        out->err_code = 0;
    }
    return 0;
}

void hostsim_free(hostsim_result* r) {
    free(r->sums);
    free(r->stamps);
    free(r->novel);
    free(r->sparse);
    memset(r, 0, sizeof *r);
}

}  // extern "C"
