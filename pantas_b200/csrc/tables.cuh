// tables.cuh -- device-resident graph / counter tables of the augment pass and the event sink the
// parsers write into (included by pantas_aug.cu inside its anonymous namespace).
//
// Replaces the reference's dictionaries (REF:n = /root/reference/scripts/alignments_augmentation_from_gaf.py:n)
//   nodes_info     REF:118,121-126   id -> (len, [IL dict, OL dict])
//   nodes_weights  REF:116,263-269   id -> NC
//   weights        REF:115,357-363   (from, to) -> RC
//
// Layout goal: one path step touches ONE 32-byte sector.
//
//   NodeRec[idx]  (32 B, idx = id - min_id)
//     +0  len        u32   sequence length, NODE_LEN_ABSENT if the GFA has no such S line
//     +4  il_stamp   u32   first-touch stamp of IL[v][0]      (file offset - epoch base; UNSET32 = never)
//     +8  ol_stamp   u32   first-touch stamp of OL[v][len(v)]
//     +12 d0, d1     i16   to_idx - idx of up to two out-links held inline (0 = none)
//     +16 c0         u64   low half: NC (part A), high half: RC of inline link 0
//     +24 c1         u64   low half: NC (part B), high half: RC of inline link 1
//   A step of a read adds 1 to NC of its node and 1 to RC of the link it leaves the node by: when
//   that link is inline both happen in ONE RED.ADD.64 of (1 | 1 << 32) on c0 or c1 (NC = lo(c0) + lo(c1)).
//   Links that are not inline (third out-link of a node, |to - from| >= 2^15, self loops) live in a
//   64-bit-key open-addressing table `ovf` (read-only keys, RED.ADD.64 counts); links that are not in
//   the GFA at all go to the CAS-insert `novel` table; deletion-derived IL/OL keys to `sparse`.
//
//   IL[v][0] and OL[v][len] are stored as NC[v] + adj[v] (il_adj32 / ol_adj32): interior nodes of a
//   read contribute adj = 0, so only the two ends of a read touch those arrays.
//
// The 32-bit halves and stamps are relative to an EPOCH (a run of chunks spanning < 4 GiB of GAF).
// fold_epoch_kernel adds them into the 64-bit totals (nc64, rc64, adj64, stamp64) and clears them;
// the host folds before an epoch could overflow and before every export, so results are exact
// 64-bit counts and global (file offset << 2 | e) stamps.
#pragma once

constexpr uint64_t KEY_EMPTY = 0xFFFFFFFFFFFFFFFFull;
constexpr uint64_t STAMP_UNSET = 0x7FFFFFFFFFFFFFFFull;   // INT64_MAX: exported as int64, reduced with MIN
constexpr uint32_t UNSET32 = 0xFFFFFFFFu;
constexpr uint32_t NO_EDGE = 0xFFFFFFFFu;
constexpr uint64_t EPOCH_SPAN = 0xFFFFFFF0ull;            // an epoch covers file offsets [base, base + EPOCH_SPAN)

struct __align__(32) NodeRec {
    uint32_t len;
    uint32_t il_stamp;
    uint32_t ol_stamp;
    uint32_t d01;                 // d0 | d1 << 16 (two's complement i16 each)
    unsigned long long c0;
    unsigned long long c1;
};
struct __align__(16) OvfSlot {
    unsigned long long key;
    unsigned long long count;
};
struct __align__(32) SideSlot {
    unsigned long long key;
    unsigned long long count;
    unsigned long long stamp;
    unsigned long long pad;
};

// device scalars (unsigned long long each)
enum { SC_REJ = 0, SC_LINES, SC_ERR, SC_NOVEL_USED, SC_SPARSE_USED, SC_DEFERRED_TOTAL, SC_TILES, SC_TILE_NEXT, SC_NDEFER,
       SC_WHY = 16 /* 16 hand-over reasons */, SC_PHASE = 32 /* 16 phase clocks (diagnostics) */, SC_COUNT = 48 };

struct Tables {
    NodeRec* nodes;
    int32_t* il_adj32;
    int32_t* ol_adj32;
    OvfSlot* ovf;
    uint32_t* ovf_edge;            // slot -> L-line edge index (export only)
    uint32_t* inl_edge;            // [2N] inline slot -> edge index (fold only)
    SideSlot* novel;
    SideSlot* sparse;
    unsigned long long* sc;
    uint32_t* deferred;            // chunk-relative starts of records redone from global memory
    // 64-bit totals (cold: touched by fold / export only)
    long long* nc64;
    long long* il_adj64;
    long long* ol_adj64;
    unsigned long long* il_st64;
    unsigned long long* ol_st64;
    long long* rc64;               // [E]
    uint64_t n_nodes;
    uint64_t ovf_mask;
    uint64_t novel_mask;
    uint64_t sparse_mask;
    uint64_t deferred_cap;
    int64_t epoch_base;            // file offset the 32-bit stamps are relative to
    uint32_t min_id;
};

__device__ __forceinline__ uint64_t mix64(uint64_t h) {
    h ^= h >> 33; h *= 0xff51afd7ed558ccdull; h ^= h >> 33; h *= 0xc4ceb9fe1a85ec53ull; h ^= h >> 33;
    return h;
}

__device__ __forceinline__ void report_error(const Tables& T, int code, int64_t off) {
    atomicMin(&T.sc[SC_ERR], ((unsigned long long)off << 8) | (unsigned long long)code);
}

// insert-or-increment in a 64-bit-key open-addressing table (linear probing, CAS claim)
__device__ __forceinline__ void side_add(SideSlot* tab, uint64_t mask, unsigned long long* used, uint64_t key,
                                         uint64_t stamp, const Tables& T, int full_code) {
    uint64_t h = mix64(key) & mask;
    for (uint64_t probes = 0; probes <= mask; probes++) {
        unsigned long long k = *(volatile unsigned long long*)&tab[h].key;
        if (k == KEY_EMPTY) {
            k = atomicCAS(&tab[h].key, KEY_EMPTY, (unsigned long long)key);
            if (k == KEY_EMPTY) {
                unsigned long long n = atomicAdd(used, 1ull);
                if (n * 4 >= (mask + 1) * 3) report_error(T, full_code, (int64_t)(stamp >> 2));
                k = key;
            }
        }
        if (k == key) {
            atomicAdd(&tab[h].count, 1ull);
            atomicMin(&tab[h].stamp, (unsigned long long)stamp);
            return;
        }
        h = (h + 1) & mask;
    }
    report_error(T, full_code, (int64_t)(stamp >> 2));
}

__device__ __forceinline__ int32_t sext16(uint32_t v) { return (int32_t)(int16_t)(uint16_t)v; }

struct DevSink {
    const Tables& T;
    uint32_t rej;
    __device__ __forceinline__ explicit DevSink(const Tables& t) : T(t), rej(0) {}

    struct Stamps { uint32_t il, ol; };
    struct EdgePf { uint32_t from; uint32_t d01; };          // inline link deltas of node `from`
    struct Hot { uint32_t len, il, ol, d01; };               // the read half of a NodeRec

    __device__ __forceinline__ bool id_to_idx(uint64_t id, uint32_t& idx) {
        const uint64_t d = id - T.min_id;                          // wraps to huge when id < min_id
        if (d >= T.n_nodes) return false;
        idx = (uint32_t)d;
        return true;
    }
    __device__ __forceinline__ void prefetch_node(uint32_t idx) {
#ifndef PT_EMU
        asm volatile("prefetch.global.L2 [%0];" ::"l"(&T.nodes[idx]));
#else
        (void)idx;
#endif
    }
    // L2 (always current): first-touch stamps only ever decrease, len / d01 never change
    __device__ __forceinline__ Hot load_hot(uint32_t idx) {
        const uint4 v = __ldcg(reinterpret_cast<const uint4*>(&T.nodes[idx]));
        Hot h;
        h.len = v.x; h.il = v.y; h.ol = v.z; h.d01 = v.w;
        return h;
    }
    __device__ __forceinline__ uint32_t load_len(uint32_t idx) { return __ldg(&T.nodes[idx].len); }
    __device__ __forceinline__ Stamps load_stamps(uint32_t idx) {
        Stamps r;
        r.il = __ldcg(&T.nodes[idx].il_stamp);
        r.ol = __ldcg(&T.nodes[idx].ol_stamp);
        return r;
    }
    __device__ __forceinline__ void edge_pf_init(EdgePf& pf) { pf.from = 0xFFFFFFFFu; pf.d01 = 0; }
    __device__ __forceinline__ void prefetch_edge(EdgePf& pf, uint32_t from) {
        pf.d01 = __ldg(&T.nodes[from].d01);
        pf.from = from;
    }
    // which inline slot of `from` (deltas d01) holds the link to `to`: 0, 1 or -1
    static __device__ __forceinline__ int inline_slot(uint32_t d01, uint32_t from, uint32_t to) {
        const int64_t delta = (int64_t)to - (int64_t)from;
        if (delta == 0 || delta < -32768 || delta > 32767) return -1;
        if ((int32_t)delta == sext16(d01 & 0xFFFFu)) return 0;
        if ((int32_t)delta == sext16(d01 >> 16)) return 1;
        return -1;
    }
    // NC[idx] += 1 and, if slot >= 0, RC of that inline link += 1: one RED.ADD.64
    __device__ __forceinline__ void bump(uint32_t idx, int slot) {
        unsigned long long* c = &T.nodes[idx].c0 + (slot > 0 ? 1 : 0);       // c1 follows c0
        atomicAdd(c, slot >= 0 ? 0x100000001ull : 1ull);
    }
    // RC of a link that is not inline: known link -> ovf table, otherwise novel (REF:426-427)
    __device__ __forceinline__ void edge_far(uint32_t a, uint32_t b, uint64_t stamp) {
        const uint64_t key = ((uint64_t)a << 32) | b;
        uint64_t h = mix64(key) & T.ovf_mask;
        for (;;) {
            const unsigned long long k = __ldg(&T.ovf[h].key);
            if (k == key) { atomicAdd(&T.ovf[h].count, 1ull); return; }
            if (k == KEY_EMPTY) break;
            h = (h + 1) & T.ovf_mask;
        }
        side_add(T.novel, T.novel_mask, &T.sc[SC_NOVEL_USED], key, stamp, T, pt::PT_X_NOVEL_FULL);
    }
    // IL[idx][0] += il, OL[idx][len] += ol relative to the NC increment of the same step; stamp = (offset << 2) | 1
    __device__ __forceinline__ void dense(uint32_t idx, int64_t il, int64_t ol, uint64_t stamp, const Stamps& st) {
        if (il != 1) atomicAdd(&T.il_adj32[idx], (int32_t)(il - 1));
        if (ol != 1) atomicAdd(&T.ol_adj32[idx], (int32_t)(ol - 1));
        // first-touch stamps: write only when we are earlier than what was there a moment ago
        const uint32_t rel = (uint32_t)((int64_t)(stamp >> 2) - T.epoch_base);
        if (il > 0 && rel < st.il) atomicMin(&T.nodes[idx].il_stamp, rel);
        if (ol > 0 && rel < st.ol) atomicMin(&T.nodes[idx].ol_stamp, rel);
    }

    // the same with the stamp test made by the caller (per tile: could this tile lower the stamp at all?)
    __device__ __forceinline__ void dense_flagged(uint32_t idx, int64_t il, int64_t ol, uint64_t stamp, bool need_il, bool need_ol) {
        if (il != 1) atomicAdd(&T.il_adj32[idx], (int32_t)(il - 1));
        if (ol != 1) atomicAdd(&T.ol_adj32[idx], (int32_t)(ol - 1));
        const uint32_t rel = (uint32_t)((int64_t)(stamp >> 2) - T.epoch_base);
        if (il > 0 && need_il) atomicMin(&T.nodes[idx].il_stamp, rel);
        if (ol > 0 && need_ol) atomicMin(&T.nodes[idx].ol_stamp, rel);
    }

    // ---- the per-record interface of line_core.cuh (slow path)
    __device__ __forceinline__ void count_node(uint32_t idx) { atomicAdd(&T.nodes[idx].c0, 1ull); }
    __device__ __forceinline__ void sparse(uint32_t idx, int dir, int64_t pos, uint64_t stamp) {
        const int64_t bias = 1ll << 30;
        if (pos < -bias || pos >= bias) { report_error(T, pt::PT_U_POSITION, (int64_t)(stamp >> 2)); return; }
        const uint64_t key = ((uint64_t)idx << 32) | ((uint64_t)dir << 31) | (uint64_t)(pos + bias);
        side_add(T.sparse, T.sparse_mask, &T.sc[SC_SPARSE_USED], key, stamp, T, pt::PT_X_SPARSE_FULL);
    }
    __device__ __forceinline__ void edge(uint32_t a, uint32_t b, uint64_t stamp, const EdgePf& pf) {
        const uint32_t d01 = pf.from == a ? pf.d01 : __ldg(&T.nodes[a].d01);
        const int slot = inline_slot(d01, a, b);
        if (slot >= 0) atomicAdd(slot ? &T.nodes[a].c1 : &T.nodes[a].c0, 0x100000000ull);
        else edge_far(a, b, stamp);
    }
    __device__ __forceinline__ void reject() { rej++; }
    __device__ __forceinline__ void error(int code, int64_t off) { report_error(T, code, off); }
};

// ---------------------------------------------------------------- graph build / fold / reset / export

__global__ void init_nodes_kernel(Tables T, const uint32_t* len) {
    const uint64_t n = T.n_nodes;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        NodeRec r;
        r.len = len[i];
        r.il_stamp = UNSET32;
        r.ol_stamp = UNSET32;
        r.d01 = 0;
        r.c0 = 0;
        r.c1 = 0;
        T.nodes[i] = r;
        T.inl_edge[2 * i] = NO_EDGE;
        T.inl_edge[2 * i + 1] = NO_EDGE;
    }
}
__global__ void clear_ovf_kernel(OvfSlot* e, uint32_t* ovf_edge, uint64_t cap, int keys_too) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < cap; i += (uint64_t)gridDim.x * blockDim.x) {
        if (keys_too) { e[i].key = KEY_EMPTY; ovf_edge[i] = NO_EDGE; }
        e[i].count = 0;
    }
}
// pass 1: claim inline slots (first come first served; any assignment gives the same counts)
__global__ void inline_edges_kernel(Tables T, const uint64_t* keys, uint64_t n_edges, unsigned long long* stats) {
    for (uint64_t e = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; e < n_edges; e += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t key = keys[e];
        const uint64_t from = key >> 32, to = key & 0xFFFFFFFFull;
        if (from >= T.n_nodes || to >= T.n_nodes) { atomicAdd(&stats[0], 1ull); continue; }
        const int64_t delta = (int64_t)to - (int64_t)from;
        bool placed = false;
        if (delta != 0 && delta >= -32768 && delta <= 32767) {
            const uint32_t d16 = (uint32_t)delta & 0xFFFFu;
            uint32_t* w = &T.nodes[from].d01;
            uint32_t old = *w;
            for (;;) {
                uint32_t neu;
                int slot;
                if ((old & 0xFFFFu) == d16 || (old >> 16) == d16) { atomicAdd(&stats[0], 1ull); placed = true; break; }   // duplicate key
                if ((old & 0xFFFFu) == 0) { neu = old | d16; slot = 0; }
                else if ((old >> 16) == 0) { neu = old | (d16 << 16); slot = 1; }
                else break;
                const uint32_t seen = atomicCAS(w, old, neu);
                if (seen == old) { T.inl_edge[2 * from + slot] = (uint32_t)e; placed = true; break; }
                old = seen;
            }
        }
        if (!placed) atomicAdd(&stats[1], 1ull);           // goes to the ovf table in pass 2
    }
}
// pass 2: everything that is not inline goes to the ovf hash table
__global__ void ovf_edges_kernel(Tables T, const uint64_t* keys, uint64_t n_edges, unsigned long long* stats) {
    for (uint64_t e = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; e < n_edges; e += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t key = keys[e];
        const uint64_t from = key >> 32, to = key & 0xFFFFFFFFull;
        if (from >= T.n_nodes || to >= T.n_nodes) continue;
        if (T.inl_edge[2 * from] == (uint32_t)e || T.inl_edge[2 * from + 1] == (uint32_t)e) continue;
        if (DevSink::inline_slot(T.nodes[from].d01, (uint32_t)from, (uint32_t)to) >= 0) continue;    // duplicate of an inline key (counted in pass 1)
        uint64_t h = mix64(key) & T.ovf_mask;
        for (uint64_t probes = 0; probes <= T.ovf_mask; probes++) {
            const unsigned long long k = atomicCAS(&T.ovf[h].key, KEY_EMPTY, (unsigned long long)key);
            if (k == KEY_EMPTY) { T.ovf_edge[h] = (uint32_t)e; break; }
            if (k == key) { atomicAdd(&stats[0], 1ull); break; }          // duplicate key: caller must de-duplicate
            h = (h + 1) & T.ovf_mask;
        }
    }
}

// 32-bit epoch state -> 64-bit totals; clears the epoch state
__global__ void fold_epoch_kernel(Tables T) {
    const uint64_t N = T.n_nodes;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < N; i += (uint64_t)gridDim.x * blockDim.x) {
        NodeRec r = T.nodes[i];
        const int32_t ia = T.il_adj32[i], oa = T.ol_adj32[i];
        if ((r.c0 | r.c1) != 0) {
            T.nc64[i] += (long long)((r.c0 & 0xFFFFFFFFull) + (r.c1 & 0xFFFFFFFFull));
            const uint32_t e0 = T.inl_edge[2 * i], e1 = T.inl_edge[2 * i + 1];
            if ((r.c0 >> 32) != 0 && e0 != NO_EDGE) T.rc64[e0] += (long long)(r.c0 >> 32);
            if ((r.c1 >> 32) != 0 && e1 != NO_EDGE) T.rc64[e1] += (long long)(r.c1 >> 32);
            T.nodes[i].c0 = 0;
            T.nodes[i].c1 = 0;
        }
        if (ia) { T.il_adj64[i] += ia; T.il_adj32[i] = 0; }
        if (oa) { T.ol_adj64[i] += oa; T.ol_adj32[i] = 0; }
        if (r.il_stamp != UNSET32) {
            const unsigned long long s = ((unsigned long long)(T.epoch_base + (int64_t)r.il_stamp) << 2) | 1ull;
            if (s < T.il_st64[i]) T.il_st64[i] = s;
            T.nodes[i].il_stamp = UNSET32;
        }
        if (r.ol_stamp != UNSET32) {
            const unsigned long long s = ((unsigned long long)(T.epoch_base + (int64_t)r.ol_stamp) << 2) | 1ull;
            if (s < T.ol_st64[i]) T.ol_st64[i] = s;
            T.nodes[i].ol_stamp = UNSET32;
        }
    }
}
__global__ void reset_nodes_kernel(Tables T) {
    const uint64_t N = T.n_nodes;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < N; i += (uint64_t)gridDim.x * blockDim.x) {
        T.nodes[i].il_stamp = UNSET32;
        T.nodes[i].ol_stamp = UNSET32;
        T.nodes[i].c0 = 0;
        T.nodes[i].c1 = 0;
        T.il_st64[i] = STAMP_UNSET;
        T.ol_st64[i] = STAMP_UNSET;
    }
}
__global__ void clear_side_kernel(SideSlot* s, uint64_t cap) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < cap; i += (uint64_t)gridDim.x * blockDim.x) {
        SideSlot z;
        z.key = KEY_EMPTY;
        z.count = 0;
        z.stamp = STAMP_UNSET;
        z.pad = 0;
        s[i] = z;
    }
}
// after fold: sums = [nc | il_adj | ol_adj | rc | rej, n_lines, 0, 0], stamps = [il | ol]
__global__ void export_nodes_kernel(Tables T, long long* sums, long long* stamps, uint64_t n_edges) {
    const uint64_t N = T.n_nodes;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < N; i += (uint64_t)gridDim.x * blockDim.x) {
        sums[i] = T.nc64[i];
        sums[N + i] = T.il_adj64[i];
        sums[2 * N + i] = T.ol_adj64[i];
        stamps[i] = (long long)T.il_st64[i];
        stamps[N + i] = (long long)T.ol_st64[i];
    }
    long long* rc = sums + 3 * N;
    for (uint64_t e = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; e < n_edges; e += (uint64_t)gridDim.x * blockDim.x)
        rc[e] = T.rc64[e];
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        long long* tail = sums + 3 * N + n_edges;
        tail[0] = (long long)T.sc[SC_REJ];
        tail[1] = (long long)T.sc[SC_LINES];
        tail[2] = 0;
        tail[3] = 0;
    }
}
// after export_nodes_kernel (same stream): links held by the ovf table
__global__ void export_ovf_kernel(Tables T, long long* rc) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i <= T.ovf_mask; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t e = T.ovf_edge[i];
        if (e != NO_EDGE) rc[e] = (long long)T.ovf[i].count;
    }
}
__global__ void compact_side_kernel(const SideSlot* s, uint64_t cap, unsigned long long* out, uint64_t rows,
                                    unsigned long long* cursor) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < cap; i += (uint64_t)gridDim.x * blockDim.x) {
        const SideSlot v = s[i];
        if (v.key == KEY_EMPTY) continue;
        const unsigned long long j = atomicAdd(cursor, 1ull);
        if (j < rows) {
            out[3 * j] = v.key;
            out[3 * j + 1] = v.count;
            out[3 * j + 2] = v.stamp;
        }
    }
}
