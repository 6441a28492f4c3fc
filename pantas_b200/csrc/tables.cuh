// tables.cuh -- device-resident graph / counter tables of the augment pass and the event sink the
// parsers write into (included by pantas_aug.cu inside its anonymous namespace).
//
// Replaces the reference's dictionaries (REF:n = /root/reference/scripts/alignments_augmentation_from_gaf.py:n)
//   nodes_info     REF:118,121-126   id -> (len, [IL dict, OL dict])
//   nodes_weights  REF:116,263-269   id -> NC
//   weights        REF:115,357-363   (from, to) -> RC
//
// Layout goal: one path step is ONE 16-byte load and ONE 32-bit RED on the same 16 bytes, and the whole hot
// table (16 B x nodes: 74 MB for the 4.6 M-node dm-full graph) stays resident in the 126 MB L2.
//
//   NodeHot[idx]  (16 B, idx = id - min_id; two nodes per 32-byte sector)
//     +0  meta   u32   bits 0..9   sequence length (0 = no such S line, 1023 = longer: len_full[idx])
//                      bits 10..19 d0, bits 20..29 d1: to_idx - idx of up to two out-links held inline
//                                  (two's complement, 0 = none)
//                      bit 30 / 31 IL / OL first-touch stamp is settled (no record still to come can lower it)
//     +4  t      u32   occurrences of the node that leave it by no inline link and no hashed link ("terminal")
//     +8  rc0    u32   RC of inline link 0
//     +12 rc1    u32   RC of inline link 1
//
//   A surviving path step adds 1 to exactly one counter: the RC of the link the read leaves the node by
//   (inline: rc0 / rc1; third out-link, |to - from| >= 512, self loop: the `ovf` 64-bit-key open-addressing
//   table; not in the GFA at all: the CAS-insert `novel` table), or `t` when the read ends there.  The
//   reference's other dense counters are sums of these and are formed once, at export:
//     NC[v]          = t[v] + sum of RC over links leaving v                                      (REF:263-269)
//     OL[v][len(v)]  =        sum of RC over links leaving v  + ol_ex[v]                          (REF:306-313,343-351)
//     IL[v][0]       =        sum of RC over links entering v + il_ex[v]                          (REF:298-305,335-342)
//   (a node occurrence increments IL[v][0] iff an in-link of this read ends there, OL[v][len] iff an out-link
//   starts there; `*_ex` = (counting ops of the node's compacted cs slice) - 1, non-zero only for nodes that
//   a mismatch / indel falls into).  Deletion-derived IL/OL keys go to the CAS-insert `sparse` table.
//
// First-touch stamps (Python dict insertion order of IL/OL keys) live in a cold
// array st32[idx] = {il, ol}: a step reads it only while the node's `settled` bit is clear, lowers it with
// RED.MIN, and sets the bit once the stamp is below the kernel's low-water mark (every tile before the mark
// is complete, tiles are handed out in file order, so nothing still to come can lower the stamp).
//
// The 32-bit counters and stamps are relative to an EPOCH (a run of chunks spanning < 4 GiB of GAF).
// fold_epoch_kernel adds them into the 64-bit totals and clears them; the host folds before an epoch could
// overflow and before every export, so results are exact 64-bit counts and global (file offset << 2 | e) stamps.
#pragma once

constexpr uint64_t KEY_EMPTY = 0xFFFFFFFFFFFFFFFFull;
constexpr uint64_t STAMP_UNSET = 0x7FFFFFFFFFFFFFFFull;   // INT64_MAX: exported as int64, reduced with MIN
constexpr uint32_t UNSET32 = 0xFFFFFFFFu;
constexpr uint32_t NO_EDGE = 0xFFFFFFFFu;
constexpr uint64_t EPOCH_SPAN = 0xFFFFFFF0ull;            // an epoch covers file offsets [base, base + EPOCH_SPAN)

constexpr uint32_t META_LEN_MASK = 0x3FFu, META_LEN_ESC = 0x3FFu;
constexpr int META_D0_SHIFT = 10, META_D1_SHIFT = 20;
constexpr uint32_t META_D_MASK = 0x3FFu;
constexpr uint32_t META_IL_SETTLED = 1u << 30, META_OL_SETTLED = 1u << 31;
constexpr int32_t INLINE_DELTA_MIN = -512, INLINE_DELTA_MAX = 511;

struct __align__(16) NodeHot {
    uint32_t meta;
    uint32_t t;
    uint32_t rc0;
    uint32_t rc1;
};
struct __align__(8) Stamp32 {
    uint32_t il, ol;
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
enum { SC_REJ = 0, SC_LINES, SC_ERR, SC_NOVEL_USED, SC_SPARSE_USED, SC_DEFERRED_TOTAL, SC_TILES, SC_LWM, SC_NDEFER,
       SC_WHY = 16 /* 16 hand-over reasons */, SC_PHASE = 32 /* 16 phase clocks (diagnostics) */, SC_COUNT = 48 };

struct Tables {
    NodeHot* nodes;
    Stamp32* st32;                 // cold: first-touch stamps of the open epoch
    uint32_t* len_full;            // cold: sequence lengths (read for nodes of >= 1023 bases only)
    int32_t* il_ex32;              // cold: (counting ops - 1) of multi-op node slices
    int32_t* ol_ex32;
    OvfSlot* ovf;
    uint32_t* ovf_edge;            // slot -> L-line edge index (export only)
    uint32_t* inl_edge;            // [2N] inline slot -> edge index (fold only)
    SideSlot* novel;
    SideSlot* sparse;
    uint32_t* novel_list;          // [novel cap] slots in use, in claim order (export walks these instead of the whole table)
    uint32_t* sparse_list;
    unsigned long long* sc;
    uint32_t* deferred;            // chunk-relative starts of records redone from global memory
    uint32_t* team_tile;           // [team_cap] tile every team of the running fast kernel works on (low-water mark)
    // 64-bit totals (cold: touched by fold / export only)
    long long* t64;
    long long* il_ex64;
    long long* ol_ex64;
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
    uint32_t team_cap;
};

__device__ __forceinline__ uint64_t mix64(uint64_t h) {
    h ^= h >> 33; h *= 0xff51afd7ed558ccdull; h ^= h >> 33; h *= 0xc4ceb9fe1a85ec53ull; h ^= h >> 33;
    return h;
}

// (out-of-line helpers take plain pointers: a reference to the kernel's parameter struct would force a local copy of it
// and turn every table access into a generic-address one)
__device__ __noinline__ void report_error_sc(unsigned long long* sc, int code, int64_t off) {
    atomicMin(&sc[SC_ERR], ((unsigned long long)off << 8) | (unsigned long long)code);
}
__device__ __forceinline__ void report_error(const Tables& T, int code, int64_t off) { report_error_sc(T.sc, code, off); }

// insert-or-increment in a 64-bit-key open-addressing table (linear probing, CAS claim); out of line: rare, and the hot
// loops stay small
__device__ __noinline__ void side_add(SideSlot* tab, uint64_t mask, unsigned long long* used, uint32_t* list, uint64_t key,
                                      uint64_t stamp, unsigned long long* sc, int full_code) {
    uint64_t h = mix64(key) & mask;
    for (uint64_t probes = 0; probes <= mask; probes++) {
        unsigned long long k = *(volatile unsigned long long*)&tab[h].key;
        if (k == KEY_EMPTY) {
            k = atomicCAS(&tab[h].key, KEY_EMPTY, (unsigned long long)key);
            if (k == KEY_EMPTY) {
                unsigned long long n = atomicAdd(used, 1ull);
                if (n <= mask) list[n] = (uint32_t)h;
                if (n * 4 >= (mask + 1) * 3) report_error_sc(sc, full_code, (int64_t)(stamp >> 2));
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
    report_error_sc(sc, full_code, (int64_t)(stamp >> 2));
}

__device__ __forceinline__ int32_t sext10(uint32_t v) { return (int32_t)(v << 22) >> 22; }

struct DevSink {
    const Tables& T;
    uint32_t rej;
    __device__ __forceinline__ explicit DevSink(const Tables& t) : T(t), rej(0) {}

    struct Stamps { uint32_t il, ol; };
    struct EdgePf { uint32_t from; uint32_t meta; };         // meta word of node `from`

    __device__ __forceinline__ bool id_to_idx(uint64_t id, uint32_t& idx) {
        const uint64_t d = id - T.min_id;                          // wraps to huge when id < min_id
        if (d >= T.n_nodes) return false;
        idx = (uint32_t)d;
        return true;
    }
    // the whole hot record from the L2 (counters are only ever RED-updated, never cached in L1)
    __device__ __forceinline__ uint4 load_hot(uint32_t idx) { return __ldcg(reinterpret_cast<const uint4*>(&T.nodes[idx])); }
    __device__ __forceinline__ uint32_t load_meta(uint32_t idx) { return __ldcg(&T.nodes[idx].meta); }
    __device__ __forceinline__ uint32_t load_len(uint32_t idx) {
        const uint32_t l = __ldcg(&T.nodes[idx].meta) & META_LEN_MASK;
        if (l == 0u) return pt::NODE_LEN_ABSENT;
        return l == META_LEN_ESC ? __ldg(&T.len_full[idx]) : l;
    }
    __device__ __forceinline__ Stamps load_stamps(uint32_t idx) {
        const uint2 s = __ldcg(reinterpret_cast<const uint2*>(&T.st32[idx]));
        Stamps r;
        r.il = s.x;
        r.ol = s.y;
        return r;
    }
    __device__ __forceinline__ void edge_pf_init(EdgePf& pf) { pf.from = 0xFFFFFFFFu; pf.meta = 0; }
    __device__ __forceinline__ void prefetch_edge(EdgePf& pf, uint32_t from) {
        pf.meta = load_meta(from);
        pf.from = from;
    }
    // which inline slot of `from` (meta word `meta`) holds the link to `to`: 0, 1 or -1
    static __device__ __forceinline__ int inline_slot(uint32_t meta, uint32_t from, uint32_t to) {
        const int32_t delta = (int32_t)(to - from);
        if (delta == 0 || delta < INLINE_DELTA_MIN || delta > INLINE_DELTA_MAX) return -1;
        const uint32_t d = (uint32_t)delta & META_D_MASK;
        if (d == ((meta >> META_D0_SHIFT) & META_D_MASK)) return 0;
        if (d == ((meta >> META_D1_SHIFT) & META_D_MASK)) return 1;
        return -1;
    }
    // the one counter update of a step: slot 0 / 1 = RC of that inline link, -1 = the read ends here (t)
    __device__ __forceinline__ void bump(uint32_t idx, int slot) {
        atomicAdd(&T.nodes[idx].t + (slot + 1), 1u);                           // t, rc0, rc1 are consecutive words
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
        side_add(T.novel, T.novel_mask, &T.sc[SC_NOVEL_USED], T.novel_list, key, stamp, T.sc, pt::PT_X_NOVEL_FULL);
    }
    // (counting ops - 1) of a node occurrence that has an in-link (il) / an out-link (ol) in its read
    __device__ __forceinline__ void extras(uint32_t idx, int32_t il_ex, int32_t ol_ex) {
        if (il_ex) atomicAdd(&T.il_ex32[idx], il_ex);
        if (ol_ex) atomicAdd(&T.ol_ex32[idx], ol_ex);
    }
    // first-touch stamps of IL[idx][0] / OL[idx][len] for a step at epoch-relative offset `rel`; lwm_rel: every record
    // before this epoch-relative offset has been counted already
    __device__ __forceinline__ void touch_stamps(uint32_t idx, bool il, bool ol, uint32_t rel, uint32_t lwm_rel) {
        const uint2 s = __ldcg(reinterpret_cast<const uint2*>(&T.st32[idx]));     // {il, ol}
        uint32_t settle = 0;
        if (il) {
            if (rel < s.x) atomicMin(&T.st32[idx].il, rel);
            if (min(rel, s.x) < lwm_rel) settle |= META_IL_SETTLED;
        }
        if (ol) {
            if (rel < s.y) atomicMin(&T.st32[idx].ol, rel);
            if (min(rel, s.y) < lwm_rel) settle |= META_OL_SETTLED;
        }
        if (settle) atomicOr(&T.nodes[idx].meta, settle);
    }

    // ---- the per-record interface of line_core.cuh (exact path)
    __device__ __forceinline__ void count_node(uint32_t idx) { atomicAdd(&T.nodes[idx].t, 1u); }
    // IL[idx][0] += il, OL[idx][len] += ol; has_in / has_out: an in-link / out-link of this read was (or will be) counted
    // at this occurrence (they carry 1 each, section "layout" above)
    __device__ __forceinline__ void dense(uint32_t idx, int64_t il, int64_t ol, uint64_t stamp, const Stamps& st, bool has_in,
                                          bool has_out) {
        extras(idx, (int32_t)(il - (has_in ? 1 : 0)), (int32_t)(ol - (has_out ? 1 : 0)));
        const uint32_t rel = (uint32_t)((int64_t)(stamp >> 2) - T.epoch_base);
        if (il > 0 && rel < st.il) atomicMin(&T.st32[idx].il, rel);
        if (ol > 0 && rel < st.ol) atomicMin(&T.st32[idx].ol, rel);
    }
    __device__ __forceinline__ void sparse(uint32_t idx, int dir, int64_t pos, uint64_t stamp) {
        const int64_t bias = 1ll << 30;
        if (pos < -bias || pos >= bias) { report_error(T, pt::PT_U_POSITION, (int64_t)(stamp >> 2)); return; }
        const uint64_t key = ((uint64_t)idx << 32) | ((uint64_t)dir << 31) | (uint64_t)(pos + bias);
        side_add(T.sparse, T.sparse_mask, &T.sc[SC_SPARSE_USED], T.sparse_list, key, stamp, T.sc, pt::PT_X_SPARSE_FULL);
    }
    // count_node(a) was called for this occurrence already: move its 1 from t to the link's counter
    __device__ __forceinline__ void edge(uint32_t a, uint32_t b, uint64_t stamp, const EdgePf& pf) {
        const uint32_t meta = pf.from == a ? pf.meta : load_meta(a);
        const int slot = inline_slot(meta, a, b);
        if (slot >= 0) atomicAdd(slot ? &T.nodes[a].rc1 : &T.nodes[a].rc0, 1u);
        else edge_far(a, b, stamp);
        atomicAdd(&T.nodes[a].t, 0xFFFFFFFFu);
    }
    __device__ __forceinline__ void reject() { rej++; }
    __device__ __forceinline__ void error(int code, int64_t off) { report_error(T, code, off); }
};

// ---------------------------------------------------------------- graph build / fold / reset / export

__global__ void init_nodes_kernel(Tables T, const uint32_t* len) {
    const uint64_t n = T.n_nodes;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t l = len[i];
        NodeHot r;
        r.meta = l == pt::NODE_LEN_ABSENT ? 0u : (l >= META_LEN_ESC ? META_LEN_ESC : l);
        r.t = 0;
        r.rc0 = 0;
        r.rc1 = 0;
        T.nodes[i] = r;
        T.len_full[i] = l;
        Stamp32 s;
        s.il = UNSET32;
        s.ol = UNSET32;
        T.st32[i] = s;
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
        if (delta != 0 && delta >= INLINE_DELTA_MIN && delta <= INLINE_DELTA_MAX) {
            const uint32_t d10 = (uint32_t)delta & META_D_MASK;
            uint32_t* w = &T.nodes[from].meta;
            uint32_t old = *w;
            for (;;) {
                const uint32_t o0 = (old >> META_D0_SHIFT) & META_D_MASK, o1 = (old >> META_D1_SHIFT) & META_D_MASK;
                uint32_t neu;
                int slot;
                if (o0 == d10 || o1 == d10) { atomicAdd(&stats[0], 1ull); placed = true; break; }   // duplicate key
                if (o0 == 0) { neu = old | (d10 << META_D0_SHIFT); slot = 0; }
                else if (o1 == 0) { neu = old | (d10 << META_D1_SHIFT); slot = 1; }
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
        if (DevSink::inline_slot(T.nodes[from].meta, (uint32_t)from, (uint32_t)to) >= 0) continue;    // duplicate of an inline key (counted in pass 1)
        uint64_t h = mix64(key) & T.ovf_mask;
        for (uint64_t probes = 0; probes <= T.ovf_mask; probes++) {
            const unsigned long long k = atomicCAS(&T.ovf[h].key, KEY_EMPTY, (unsigned long long)key);
            if (k == KEY_EMPTY) { T.ovf_edge[h] = (uint32_t)e; break; }
            if (k == key) { atomicAdd(&stats[0], 1ull); break; }          // duplicate key: caller must de-duplicate
            h = (h + 1) & T.ovf_mask;
        }
    }
}

// 32-bit epoch state -> 64-bit totals; clears the epoch state (and the settled bits: they refer to the epoch's stamps)
__global__ void fold_epoch_kernel(Tables T) {
    const uint64_t N = T.n_nodes;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < N; i += (uint64_t)gridDim.x * blockDim.x) {
        const NodeHot r = T.nodes[i];
        if ((r.t | r.rc0 | r.rc1) != 0u || (r.meta & (META_IL_SETTLED | META_OL_SETTLED)) != 0u) {
            T.t64[i] += (long long)(int32_t)r.t;           // the exact path moves a count out of t again (DevSink::edge)
            const uint32_t e0 = T.inl_edge[2 * i], e1 = T.inl_edge[2 * i + 1];
            if (r.rc0 != 0u && e0 != NO_EDGE) T.rc64[e0] += (long long)r.rc0;
            if (r.rc1 != 0u && e1 != NO_EDGE) T.rc64[e1] += (long long)r.rc1;
            NodeHot z;
            z.meta = r.meta & ~(META_IL_SETTLED | META_OL_SETTLED);
            z.t = 0;
            z.rc0 = 0;
            z.rc1 = 0;
            T.nodes[i] = z;
        }
        const int32_t ia = T.il_ex32[i], oa = T.ol_ex32[i];
        if (ia) { T.il_ex64[i] += ia; T.il_ex32[i] = 0; }
        if (oa) { T.ol_ex64[i] += oa; T.ol_ex32[i] = 0; }
        const Stamp32 s = T.st32[i];
        if (s.il != UNSET32) {
            const unsigned long long v = ((unsigned long long)(T.epoch_base + (int64_t)s.il) << 2) | 1ull;
            if (v < T.il_st64[i]) T.il_st64[i] = v;
        }
        if (s.ol != UNSET32) {
            const unsigned long long v = ((unsigned long long)(T.epoch_base + (int64_t)s.ol) << 2) | 1ull;
            if (v < T.ol_st64[i]) T.ol_st64[i] = v;
        }
        if ((s.il & s.ol) != UNSET32) {
            Stamp32 u;
            u.il = UNSET32;
            u.ol = UNSET32;
            T.st32[i] = u;
        }
    }
}
__global__ void reset_nodes_kernel(Tables T) {
    const uint64_t N = T.n_nodes;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < N; i += (uint64_t)gridDim.x * blockDim.x) {
        NodeHot r = T.nodes[i];
        r.meta &= ~(META_IL_SETTLED | META_OL_SETTLED);
        r.t = 0;
        r.rc0 = 0;
        r.rc1 = 0;
        T.nodes[i] = r;
        Stamp32 u;
        u.il = UNSET32;
        u.ol = UNSET32;
        T.st32[i] = u;
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
__global__ void reset_teams_kernel(Tables T) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < T.team_cap; i += gridDim.x * blockDim.x) T.team_tile[i] = 0;
    if (blockIdx.x == 0 && threadIdx.x == 0) T.sc[SC_LWM] = 0;
}

// ---- export.  Flat layout of include/pantas_aug.h: sums = [nc | il_adj | ol_adj | rc | rej, n_lines, 0, 0] with
// IL[v][0] = nc + il_adj, OL[v][len] = nc + ol_adj; stamps = [il | ol].  Step 1 writes the per-node terms, steps 2..4
// scatter every link's count to its two ends (header comment: NC, IL, OL as sums of RC).
// The export reads, it does not fold: a value is its 64-bit total (EXP_TOTALS: at least one epoch has been folded since the
// reset -- the arrays are all zero / unset otherwise and are not read) plus the open epoch's 32-bit state (EXP_LIVE).
enum : uint32_t { EXP_TOTALS = 1u, EXP_LIVE = 2u };
__global__ void export_nodes_kernel(Tables T, long long* sums, long long* stamps, uint64_t n_edges, uint32_t flags) {
    const uint64_t N = T.n_nodes;
    const bool totals = (flags & EXP_TOTALS) != 0u, live = (flags & EXP_LIVE) != 0u;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < N; i += (uint64_t)gridDim.x * blockDim.x) {
        long long t = 0, il_ex = 0, ol_ex = 0;
        unsigned long long il_st = STAMP_UNSET, ol_st = STAMP_UNSET;
        if (totals) {
            t = T.t64[i];
            il_ex = T.il_ex64[i];
            ol_ex = T.ol_ex64[i];
            il_st = T.il_st64[i];
            ol_st = T.ol_st64[i];
        }
        if (live) {
            t += (long long)(int32_t)T.nodes[i].t;         // the exact path moves a count out of t again (DevSink::edge)
            il_ex += T.il_ex32[i];
            ol_ex += T.ol_ex32[i];
            const Stamp32 s = T.st32[i];
            if (s.il != UNSET32) {
                const unsigned long long v = ((unsigned long long)(T.epoch_base + (int64_t)s.il) << 2) | 1ull;
                if (v < il_st) il_st = v;
            }
            if (s.ol != UNSET32) {
                const unsigned long long v = ((unsigned long long)(T.epoch_base + (int64_t)s.ol) << 2) | 1ull;
                if (v < ol_st) ol_st = v;
            }
        }
        sums[i] = t;                                       // + links leaving i
        sums[N + i] = il_ex - t;                           // + links entering i - links leaving i
        sums[2 * N + i] = ol_ex - t;
        stamps[i] = (long long)il_st;
        stamps[N + i] = (long long)ol_st;
    }
    long long* rc = sums + 3 * N;                          // (every link is written again by the kernel that owns its counter)
    for (uint64_t e = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; e < n_edges; e += (uint64_t)gridDim.x * blockDim.x) rc[e] = 0;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        long long* tail = sums + 3 * N + n_edges;
        tail[0] = (long long)T.sc[SC_REJ];
        tail[1] = (long long)T.sc[SC_LINES];
        tail[2] = 0;
        tail[3] = 0;
    }
}
__device__ __forceinline__ void export_link_ends(long long* sums, uint64_t N, uint64_t from, uint64_t to, unsigned long long c) {
    if (c == 0ull) return;
    unsigned long long* s = reinterpret_cast<unsigned long long*>(sums);
    atomicAdd(&s[from], c);                                // NC[from]
    atomicAdd(&s[N + from], 0ull - c);                     // il_adj[from] = IL - NC
    atomicAdd(&s[N + to], c);                              // IL[to][0]
}
// after export_nodes_kernel (same stream): inline links -- folded epochs in rc64, the open epoch in the node's rc0 / rc1
__global__ void export_inline_kernel(Tables T, long long* sums, uint32_t flags) {
    const uint64_t N = T.n_nodes;
    const bool totals = (flags & EXP_TOTALS) != 0u, live = (flags & EXP_LIVE) != 0u;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < N; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint2 ee = *reinterpret_cast<const uint2*>(&T.inl_edge[2 * i]);
        if ((ee.x & ee.y) == NO_EDGE) continue;
        const NodeHot r = T.nodes[i];
        if (ee.x != NO_EDGE) {
            const unsigned long long c = (totals ? (unsigned long long)T.rc64[ee.x] : 0ull) + (live ? (unsigned long long)r.rc0 : 0ull);
            sums[3 * N + ee.x] = (long long)c;
            export_link_ends(sums, N, i, (uint64_t)((int64_t)i + sext10((r.meta >> META_D0_SHIFT) & META_D_MASK)), c);
        }
        if (ee.y != NO_EDGE) {
            const unsigned long long c = (totals ? (unsigned long long)T.rc64[ee.y] : 0ull) + (live ? (unsigned long long)r.rc1 : 0ull);
            sums[3 * N + ee.y] = (long long)c;
            export_link_ends(sums, N, i, (uint64_t)((int64_t)i + sext10((r.meta >> META_D1_SHIFT) & META_D_MASK)), c);
        }
    }
}
// links held by the ovf table
__global__ void export_ovf_kernel(Tables T, long long* sums) {
    const uint64_t N = T.n_nodes;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i <= T.ovf_mask; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t e = T.ovf_edge[i];
        if (e == NO_EDGE) continue;
        const OvfSlot v = T.ovf[i];
        sums[3 * N + e] = (long long)v.count;
        export_link_ends(sums, N, v.key >> 32, v.key & 0xFFFFFFFFull, v.count);
    }
}
// links that are not in the GFA (the slots in use are listed: no scan of the whole table)
__global__ void export_novel_ends_kernel(Tables T, long long* sums) {
    const uint64_t N = T.n_nodes;
    const uint64_t used = min((uint64_t)T.sc[SC_NOVEL_USED], T.novel_mask + 1);
    for (uint64_t j = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; j < used; j += (uint64_t)gridDim.x * blockDim.x) {
        const SideSlot v = T.novel[T.novel_list[j]];
        export_link_ends(sums, N, v.key >> 32, v.key & 0xFFFFFFFFull, v.count);
    }
}
// rows {key, count, stamp} of the slots in use, in claim order
__global__ void compact_side_kernel(const SideSlot* s, const uint32_t* list, uint64_t used, unsigned long long* out, uint64_t rows) {
    const uint64_t n = used < rows ? used : rows;
    for (uint64_t j = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; j < n; j += (uint64_t)gridDim.x * blockDim.x) {
        const SideSlot v = s[list[j]];
        out[3 * j] = v.key;
        out[3 * j + 1] = v.count;
        out[3 * j + 2] = v.stamp;
    }
}
