// team_tiles.cuh -- the fast path of the augment kernel (included by aug_kernels.cuh after tables.cuh, the TMA
// helpers, ChunkArgs and defer_line(); compiled for sm_100a by pantas_aug.cu and for the CPU emulator by
// tests/hostsim/fastsim.cpp).
//
// Reference loop body: /root/reference/scripts/alignments_augmentation_from_gaf.py:142-363 (REF:n).
//
// The unit of execution is a TEAM: one CTA of two warps that owns 8 KiB tiles of the GAF chunk (+ 1 KiB of look-ahead:
// a record belongs to the tile it starts in), ten such CTAs per SM.  Teams share nothing but the global tables, so a
// phase that leaves lanes idle only idles its own two warps while the other 18 warps of the SM are in other phases; the
// barriers are 64-thread barriers.  One elected thread moves the tile with a 1-D TMA bulk copy (UBLKCP, L2 evict-first);
// the copy of the team's next tile is issued as soon as the bytes are dead (after `ids`).  Phases over a tile:
//
//   scan     one thread per 64 bytes (4 x LDS.128, lane-rotated: no bank conflicts), branch-free SWAR: a 64-bit
//            whitespace mask (bytes <= 0x20) and a 64-bit mask of path separators ('>' '<') OR non-tab whitespace
//            ('\n' rides along for free: sep & ws = record-end candidates).  Record starts are ranked with one ballot per
//            warp iteration (no atomics) into the warp's own list.
//   records  warp 0 = role B for every record, warp 1 = role A (a warp takes as long for one record as for 32; ~27
//            records per tile keep both warps' lanes busy):
//              B  the 12 column boundaries (single tabs, no empty column), MAPQ and '*' filters (REF:143-148), the three
//                 coordinates (REF:151-153), then one step-list entry per separator bit of the path column (+ a
//                 sentinel); list space comes from a warp prefix sum;
//              A  skips ten boundaries by popcount; first cs token, first dv:f: token (REF:154-160,172-180), dv filter,
//                 cs string parsed into the tile's op pool (REF:10-50, incl. cigar_clipping).
//            Anything unusual hands the record to the exact per-record path (line_core.cuh, augment_deferred_kernel).
//   ids      one thread per path step: SWAR decimal parse of the id out of shared memory -> node index -> ONE 16-byte
//            load of the node's hot record (four in flight per thread); keeps index, meta word and the node's share
//            of the query (REF:215-218: first / last node shortened) in shared memory.
//   walk     warp 0, one thread per record, a short loop over the record's steps: duplicate / unknown ids, the sum of
//            the node shares against the cs length (IndexError REF:227), and -- only for records whose cs string has
//            several ops -- the prefix sums the merge walk needs (REF:205-255).  Warp 1 meanwhile drains the previous
//            tile's list of links that are not inline (hash probes).
//   fold     one thread per step of a multi-op record: clear_align / compact_align (REF:63-107) folded over the op
//            pieces that overlap the node: dropped or not, counting ops, deletion-derived IL/OL keys.
//   count    one thread per surviving step, no loads from the tables: ONE 32-bit RED (tables.cuh), stamps only while
//            the node's settled bit is clear.  Links that are not inline are listed for the next tile's walk phase.
#pragma once

namespace teamp {

constexpr uint32_t NONE32 = 0xffffffffu;
constexpr uint32_t FULL = 0xffffffffu;
constexpr uint32_t THREADS = 64;
constexpr int MAX_STEPS = 250;                // longer paths take the exact path
constexpr int MAX_OPS = 48;                   // more cs ops: exact path
constexpr int32_t MAX_NTOT = 1 << 22;         // longer cs strings: exact path
constexpr uint32_t SL_BAD = 0xFFFFu;          // sL[] entry: unknown node / share that does not fit 16 bits

enum : uint8_t { ST_FAST = 0, ST_DONE = 1, ST_DEFER = 2 };      // per role; a record's status is the maximum
enum : uint32_t { OP_MATCH = 0, OP_SUB = 1, OP_DEL = 2, OP_INS = 3, OP_EQ = 4 };   // ':' '*' '-' '+' '='  (op = kind | len << 3)

struct __align__(4) Rec {
    int32_t start;        // int(tokens[7])                                    (role B)
    int32_t end_rel1;     // int(tokens[6]) - int(tokens[8]) - 1               (role B)
    uint32_t n_tot;       // sum of the cs op lengths                          (role A)
    int32_t start_add;    // cigar_clipping: start_pos += len of a leading '+' (role A, REF:46-47)
    uint16_t s0;          // first entry of the record in the step list        (role B)
    uint16_t nsteps;      //                                                   (role B)
    uint16_t ls;          // buffer position of the record's first byte        (role B)
    uint16_t op_off;      // first op of the record in the op pool             (role A)
    uint8_t nops;         //                                                   (role A)
    uint8_t stA, stB;     // ST_* per role; walk raises stB
    uint8_t whyA;         // WHY_* when role A says ST_DEFER
    uint16_t a_off;       // multi-op records: first entry in the prefix pool  (walk)
    uint8_t whyB;
    uint8_t single;       // 1: one ':' or '=' op -- every node with a positive share survives, one counting op  (role A)
};
static_assert(sizeof(Rec) == 32 && offsetof(Rec, nops) == 24, "rec_status reads nops / stA / stB / whyA as one word");

// step list entry
constexpr uint32_t SE_POS_MASK = 0xFFFFu;     // bits 0..15  buffer position of the separator (sentinel: end of the path column)
constexpr int SE_SLOT_SHIFT = 16;             // bits 16..21 record slot
constexpr uint32_t SE_SLOT_MASK = 0x3Fu;
constexpr uint32_t SE_FIRST = 1u << 22, SE_LAST = 1u << 23, SE_REV = 1u << 24, SE_SENT = 1u << 25, SE_DROPPED = 1u << 26;
constexpr int SE_NCNT_SHIFT = 27;             // bits 27..28 counting ops of the compacted slice (0..3)
constexpr uint32_t SE_NCNT_MASK = 3u << SE_NCNT_SHIFT;
constexpr uint32_t SE_INVALID = 0xFFFFFFFFu;

template <int TILE_, int OV_, int STEP_CAP_, int MIN_CTAS_>
struct Geo {
    static constexpr int TILE = TILE_;
    static constexpr int OV = OV_;
    static constexpr int BUF = 16 + TILE + OV + 16;               // [pre 16][tile][look-ahead][pad 16]
    static constexpr int NG = (16 + TILE + OV + 63) / 64;         // 64-byte groups = 64-bit mask words
    static constexpr int LINE_CAP = 64;                           // records per tile (slots: 6 bits)
    static constexpr int STEP_CAP = STEP_CAP_;                    // typical: 15 entries per 300-byte record
    static constexpr int OPS_CAP = 192;
    static constexpr int HEAVY_CAP = 256;                         // steps of multi-op records
    static constexpr int FAR_CAP = 48;                            // links that are not inline: typically 1 per record
    static constexpr int DEL_CAP = 24;                            // steps with deletion-derived keys
    static constexpr int MIN_CTAS = MIN_CTAS_;
    // masks are dead after `records`; sL / prefix pool / heavy list live from `ids` to `fold` in the same bytes
    static constexpr int MASK_BYTES = 16 * NG;
    static constexpr int WALK_BYTES = 2 * STEP_CAP + 4 * HEAVY_CAP + 2 * HEAVY_CAP;
    static constexpr int OFF_WM = (BUF + 127) & ~127;
    static constexpr int OFF_SM = OFF_WM + 8 * NG;
    static constexpr int OFF_SL = OFF_WM;
    static constexpr int OFF_SA = OFF_SL + 2 * STEP_CAP;
    static constexpr int OFF_HEAVY = OFF_SA + 4 * HEAVY_CAP;
    static constexpr int OFF_STEP = (OFF_WM + (MASK_BYTES > WALK_BYTES ? MASK_BYTES : WALK_BYTES) + 15) & ~15;
    static constexpr int OFF_SIDX = OFF_STEP + 4 * (STEP_CAP + 4);
    static constexpr int OFF_SMETA = OFF_SIDX + 4 * STEP_CAP;
    static constexpr int OFF_OPS = OFF_SMETA + 4 * STEP_CAP;
    static constexpr int OFF_FAR = OFF_OPS + 4 * OPS_CAP;
    static constexpr int OFF_DEL = OFF_FAR + 12 * FAR_CAP;
    static constexpr int OFF_REC = (OFF_DEL + 12 * DEL_CAP + 7) & ~7;
    static constexpr int OFF_LINES = OFF_REC + (int)sizeof(Rec) * LINE_CAP;
    static constexpr int SMEM_BYTES = (OFF_LINES + 2 * LINE_CAP + 127) & ~127;
    static_assert(BUF <= 65536, "step entries hold 16-bit positions");
    static_assert(LINE_CAP <= 64, "step entries hold 6-bit record slots");
    static_assert(STEP_CAP < 65536 && OPS_CAP < 65536 && HEAVY_CAP < 65536, "records hold 16-bit list offsets");
    static_assert((SMEM_BYTES + 1024 + 64) * MIN_CTAS <= 227 * 1024, "MIN_CTAS teams must fit one SM's shared memory");
};

// 0x80 flags at bits 7/15/23/31 -> 4-bit mask in the top nibble (no carries: the partial products
// of 2^21 + 2^14 + 2^7 + 1 land on distinct bits)
__device__ __forceinline__ uint32_t gather_top(uint32_t f) { return f * 0x00204081u; }
__device__ __forceinline__ uint32_t mask16(uint32_t f0, uint32_t f1, uint32_t f2, uint32_t f3) {
    uint32_t m = gather_top(f3) >> 28;
    m = __funnelshift_l(gather_top(f2), m, 4);
    m = __funnelshift_l(gather_top(f1), m, 4);
    m = __funnelshift_l(gather_top(f0), m, 4);
    return m;
}

// no "s:" / "v:" byte pair inside: neither regex of REF:154-156,172-174 can start in this token
__device__ __forceinline__ bool token_is_inert(const uint8_t* s, uint32_t a, uint32_t b) {
    if (b - a > 48u) return false;
    uint32_t prev = 0;
    for (uint32_t q = a; q < b; q++) {
        const uint32_t c = s[q];
        if (c == ':' && (prev == 's' || prev == 'v')) return false;
        prev = c;
    }
    return true;
}

// eight bytes at buffer position a (any alignment), first byte lowest; reads up to 11 bytes past a (the buffer is padded)
__device__ __forceinline__ unsigned long long ld8(const uint8_t* s, uint32_t a) {
    const uint32_t* w = reinterpret_cast<const uint32_t*>(s + (a & ~3u));
    const uint32_t sh = (a & 3u) * 8u;
    const uint32_t w0 = w[0], w1 = w[1], w2 = w[2];
    return (unsigned long long)__funnelshift_r(w0, w1, sh) | ((unsigned long long)__funnelshift_r(w1, w2, sh) << 32);
}
// is one of the lowest n (<= 8) bytes of x a ':' ?
__device__ __forceinline__ bool has_colon8(unsigned long long x, uint32_t n) {
    const unsigned long long keep = n >= 8u ? ~0ull : ~(~0ull << (8u * n));
    const unsigned long long y = (x ^ 0x3A3A3A3A3A3A3A3Aull) | ~keep;      // 0 exactly where a kept byte is ':'
    const unsigned long long t = (y & 0x7F7F7F7F7F7F7F7Full) + 0x7F7F7F7F7F7F7F7Full;
    return (~(t | y) & 0x8080808080808080ull) != 0ull;
}
// no ':' in [a, b)
__device__ __forceinline__ bool no_colon_w(const uint8_t* s, uint32_t a, uint32_t b) {
    if (b - a > 48u) return false;
    for (uint32_t q = a; q < b; q += 8u)
        if (has_colon8(ld8(s, q), b - q)) return false;
    return true;
}
constexpr unsigned long long TAG_CS3 = 0x3A7363ull;                          // "cs:"
constexpr unsigned long long TAG_DV5 = 0x3A663A7664ull;                      // "dv:f:"
constexpr unsigned long long TAG_AS5 = 0x3A693A5341ull;                      // "AS:i:"

// four ASCII digits, most significant in the lowest byte, already xor'ed with '0'
__device__ __forceinline__ uint32_t val4(uint32_t w) {
    w = ((w * 2561u) >> 8) & 0x00FF00FFu;
    return (w * 6553601u) >> 16;
}

// decimal digits [a, a + nd) out of shared memory, 1 <= nd <= 8.  false: not a canonical decimal
// (for a node id: the reference's dict lookup fails, KeyError REF:214)
__device__ __forceinline__ bool dec8(const uint8_t* s, uint32_t a, uint32_t nd, uint32_t& out) {
    const unsigned long long x = ld8(s, a) ^ 0x3030303030303030ull;
    if (nd > 1u && ((uint32_t)x & 0xFFu) == 0u) return false;                // leading zero
    // digits to the end of the 8-byte group, zeros (leading digits) in front; later bytes fall off
    const unsigned long long y = x << (8u * (8u - nd));
    const uint32_t ylo = (uint32_t)y, yhi = (uint32_t)(y >> 32);
    if ((((ylo + 0x76767676u) | ylo) | ((yhi + 0x76767676u) | yhi)) & 0x80808080u) return false;
    out = val4(ylo) * 10000u + val4(yhi);
    return true;
}
// node id of a path step: digits [a, a + nd), up to ten of them
__device__ __forceinline__ bool step_id(const uint8_t* s, uint32_t a, uint32_t nd, uint64_t& id) {
    if (nd - 1u > 9u) return false;                          // 1..10 digits
    uint32_t low;
    if (nd <= 8u) {
        if (!dec8(s, a, nd, low)) return false;
        id = low;
        return true;
    }
    // 9 or 10 digits: the leading one or two by hand, the last eight as above (zeros allowed in front of those)
    const uint32_t lead = nd - 8u;
    const uint32_t c0 = (uint32_t)s[a] - '0';
    if (c0 - 1u > 8u) return false;                          // '1'..'9'
    uint32_t high = c0;
    if (lead == 2u) {
        const uint32_t c1 = (uint32_t)s[a + 1u] - '0';
        if (c1 > 9u) return false;
        high = high * 10u + c1;
    }
    const unsigned long long x = ld8(s, a + lead) ^ 0x3030303030303030ull;
    const uint32_t xlo = (uint32_t)x, xhi = (uint32_t)(x >> 32);
    if ((((xlo + 0x76767676u) | xlo) | ((xhi + 0x76767676u) | xhi)) & 0x80808080u) return false;
    id = (uint64_t)high * 100000000ull + (uint64_t)(val4(xlo) * 10000u + val4(xhi));
    return true;
}
// plain digits [a, b), 1..8 of them, no leading zero (anything else: false, the exact path decides)
__device__ __forceinline__ bool small_uint(const uint8_t* s, uint32_t a, uint32_t b, int32_t& out) {
    const uint32_t n = b - a;
    uint32_t v;
    if (n - 1u > 7u || !dec8(s, a, n, v)) return false;
    out = (int32_t)v;
    return true;
}

// The walkers read the 64-bit mask words as 32-bit halves (one FLO / POPC per step instead of two).
// next whitespace bit at or after the walker's position (32 bytes of the tile per half word);
// false: ran off the end of the loaded bytes
__device__ __forceinline__ bool next_ws(const uint32_t* wm32, uint32_t nhalf, uint32_t& wi, uint32_t& m, uint32_t& pos) {
    while (m == 0u) {
        if (++wi >= nhalf) return false;
        m = wm32[wi];
    }
    pos = 32u * wi + (uint32_t)(__ffs((int)m) - 1);
    m &= m - 1u;
    return true;
}

// separator bits of half word w that lie in buffer positions [a, b)
__device__ __forceinline__ uint32_t sep_word(const uint32_t* sm32, uint32_t w, uint32_t a, uint32_t b) {
    uint32_t m = sm32[w];
    if (w == (a >> 5)) m &= ~0u << (a & 31u);
    if (w == (b >> 5)) m &= ~(~0u << (b & 31u));             // b & 31 == 0: nothing of this half word is below b
    return m;
}

__device__ __forceinline__ bool is_lower(uint32_t c) { return c - 'a' <= 25u; }

// status of a record = the worse of its two roles
__device__ __forceinline__ uint32_t rec_status(const Rec& R) {
    const uint32_t w = *reinterpret_cast<const uint32_t*>(&R.nops);        // nops | stA << 8 | stB << 16 | whyA << 24: one LDS
    return max((w >> 8) & 0xFFu, (w >> 16) & 0xFFu);
}

// exclusive prefix sum over the warp (all 32 lanes take part); total = sum over the warp
__device__ __forceinline__ uint32_t warp_excl_scan(uint32_t v, uint32_t lane, uint32_t& total) {
    uint32_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(FULL, incl, o);
        if (lane >= (uint32_t)o) incl += y;
    }
    total = __shfl_sync(FULL, incl, 31);
    return incl - v;
}

template <class G>
__global__ void __launch_bounds__(THREADS, G::MIN_CTAS) augment_team_kernel(ChunkArgs A, Tables T) {
    PT_DYNAMIC_SMEM(smem);
    __shared__ __align__(8) uint64_t mbar;
    __shared__ uint32_t s_cnt[2];                // record starts found by each warp
    __shared__ uint32_t s_nent, s_nheavy, s_nfar, s_ndel, s_far_take, s_lwm;

    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t lt_mask = (1u << lane) - 1u;
    uint8_t* const buf = smem;
    unsigned long long* const wm64 = reinterpret_cast<unsigned long long*>(smem + G::OFF_WM);
    unsigned long long* const sm64 = reinterpret_cast<unsigned long long*>(smem + G::OFF_SM);
    const uint32_t* const wm32 = reinterpret_cast<const uint32_t*>(smem + G::OFF_WM);   // the same masks, as half words
    const uint32_t* const sm32 = reinterpret_cast<const uint32_t*>(smem + G::OFF_SM);
    uint16_t* const sL = reinterpret_cast<uint16_t*>(smem + G::OFF_SL);        // share of the query per step (REF:215-218), SL_BAD
    uint32_t* const sA = reinterpret_cast<uint32_t*>(smem + G::OFF_SA);        // multi-op records: cs coordinate where the step's node starts
    uint16_t* const heavy = reinterpret_cast<uint16_t*>(smem + G::OFF_HEAVY);  // steps of multi-op records
    uint32_t* const steps = reinterpret_cast<uint32_t*>(smem + G::OFF_STEP);
    uint32_t* const sidx = reinterpret_cast<uint32_t*>(smem + G::OFF_SIDX);
    uint32_t* const smeta = reinterpret_cast<uint32_t*>(smem + G::OFF_SMETA);  // NodeHot.meta of the step's node
    uint32_t* const ops = reinterpret_cast<uint32_t*>(smem + G::OFF_OPS);
    uint32_t* const far = reinterpret_cast<uint32_t*>(smem + G::OFF_FAR);      // {from, to, separator position}: filled by `count`, drained during the next tile's `walk`
    uint32_t* const dels = reinterpret_cast<uint32_t*>(smem + G::OFF_DEL);     // {step, first del, last del}
    Rec* const recs = reinterpret_cast<Rec*>(smem + G::OFF_REC);
    uint16_t* const lines = reinterpret_cast<uint16_t*>(smem + G::OFF_LINES);  // record starts: warp 0 fills the list from the front, warp 1 from the back

    if (tid == 0) {
        mbar_init(&mbar, 1);
        s_nfar = 0;
        s_ndel = 0;
        s_far_take = 0;
        s_nheavy = 0;
        s_nent = 0;
    }
    __syncthreads();

    DevSink sink(T);
    const uint32_t ablate = A.ablate;            // diagnostics: 0 = everything, k = stop every tile after phase k (profiles/ ablation ladder)
    const uint64_t nbytes16 = (A.nbytes + 15ull) & ~15ull;
    uint32_t parity = 0;
    unsigned long long my_lines = 0, my_tiles = 0;
    int64_t far_base = 0;                        // file offset of buf[0] of the tile that filled the far-link list

    auto issue_load = [&](uint32_t tile) {
        const uint64_t t0 = (uint64_t)tile * G::TILE;
        const uint64_t lo = tile ? t0 - 16 : 0;
        const uint64_t hi = min(t0 + G::TILE + G::OV, nbytes16);
        const uint32_t bytes = (uint32_t)(hi - lo);
        fence_async_smem();
        mbar_expect_tx(&mbar, bytes);
        if (A.stream_hint & 1u) tma_load_1d_stream(buf + (tile ? 0u : 16u), A.gaf + lo, bytes, &mbar);
        else tma_load_1d(buf + (tile ? 0u : 16u), A.gaf + lo, bytes, &mbar);
    };
    // threads take entries of the far-link list from a shared counter
    auto drain_far = [&]() {
        const uint32_t n_far = min(s_nfar, (uint32_t)G::FAR_CAP);
        for (;;) {
            const uint32_t j = atomicAdd(&s_far_take, 1u);
            if (j >= n_far) break;
            sink.edge_far(far[3u * j], far[3u * j + 1u], (uint64_t)(far_base + (int64_t)far[3u * j + 2u] + 1) << 2);
        }
    };

    uint32_t tile = blockIdx.x;
    if (tile < A.n_tiles && tid == 0) issue_load(tile);

    for (; tile < A.n_tiles; tile += gridDim.x) {
        const uint64_t t0 = (uint64_t)tile * G::TILE;
        const uint32_t owned = (uint32_t)min((uint64_t)G::TILE, A.nbytes - t0);
        const uint64_t hi = min(t0 + G::TILE + G::OV, nbytes16);
        const uint32_t lim = 16u + (uint32_t)(min(hi, A.nbytes) - t0);     // data ends here in the buffer
        const int64_t base_off = A.file_off + (int64_t)t0 - 16;            // file offset of buf[0]
        const uint32_t own_end = 16u + owned;                               // records starting before this are ours
        const uint32_t nwords = (lim + 63u) >> 6;
        const uint32_t nxt_tile = tile + gridDim.x;
        if (tid == 0) {
            // low-water mark of the running kernel (tables.cuh): every tile below it is complete
            const unsigned long long lw = *(volatile unsigned long long*)&T.sc[SC_LWM];
            const int64_t rel = A.file_off - T.epoch_base + (int64_t)(lw * (unsigned long long)G::TILE);
            s_lwm = rel <= 0 ? 0u : (rel > 0xFFFFFFF0ll ? 0xFFFFFFF0u : (uint32_t)rel);
        }
        mbar_wait(&mbar, parity);
        parity ^= 1;
        if (ablate == 1u) {                                                 // TMA only
            __syncthreads();
            if (tid == 0 && nxt_tile < A.n_tiles) issue_load(nxt_tile);
            continue;
        }

        // ================= scan: whitespace / separator masks, record starts =================
        uint32_t my_cnt = 0;                                                // warp-uniform: record starts this warp has listed
        if (tile == 0 && warp == 0 && owned > 0u) {                         // the chunk starts at a record start
            if (lane == 0) lines[0] = 16;
            my_cnt = 1;
        }
        for (uint32_t g0 = 0; g0 < nwords; g0 += THREADS) {
            const uint32_t g = g0 + tid;
            unsigned long long cand = 0;
            if (g < nwords) {
                unsigned long long wm = 0, sm = 0;
                uint32_t hib = 0;
                // the four vectors of the group in a lane-dependent order: a quarter warp's eight LDS.128 then fall into
                // eight different 16-byte bank groups (lane stride 64 bytes alone would put them into two)
                const uint32_t rot = (lane >> 1) & 3u;
                uint4 q[4];
#pragma unroll
                for (int u = 0; u < 4; u++) q[u] = *reinterpret_cast<const uint4*>(buf + 64u * g + 16u * ((u + rot) & 3u));
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const uint32_t sh = 16u * ((u + rot) & 3u);
                    uint32_t wf[4], sf[4];
                    const uint32_t x4[4] = {q[u].x, q[u].y, q[u].z, q[u].w};
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        const uint32_t x = x4[k];
                        // exact for ASCII bytes; a byte >= 0x80 is a fatal PT_U_NON_ASCII error anyway
                        const uint32_t b = (x | 0x80808080u) - 0x21212121u;          // bit 7 clear: x <= 0x20
                        const uint32_t ts = ((x | 0x02020202u) ^ 0x3E3E3E3Eu) + 0x7F7F7F7Fu;   // bit 7 clear: '>' or '<'
                        const uint32_t t = x + 0x76767676u;                          // bit 7 set: x >= 0x0A
                        wf[k] = ~b & 0x80808080u;
                        sf[k] = (~ts & 0x80808080u) | (wf[k] & t);                   // separators, and whitespace that is not a tab
                        hib |= x;
                    }
                    wm |= (unsigned long long)mask16(wf[0], wf[1], wf[2], wf[3]) << sh;
                    sm |= (unsigned long long)mask16(sf[0], sf[1], sf[2], sf[3]) << sh;
                }
                const uint32_t room = lim > 64u * g ? lim - 64u * g : 0u;   // loaded bytes in this group
                unsigned long long keep = room < 64u ? ~(~0ull << room) : ~0ull;
                cand = wm & sm & keep;                                      // '\n' (record end), '\r', other odd whitespace
                if (g == 0) {
                    keep &= ~0xFFFFull;                                     // positions 0..15 are before the tile ...
                    cand &= tile != 0 ? ~0x7FFFull : ~0xFFFFull;            // ... but is the byte before the tile a newline?
                }
                wm64[g] = wm & keep;
                sm64[g] = sm & keep;
                if ((hib & 0x80808080u) != 0u) {                            // non-ASCII byte: not modelled
                    for (uint32_t p = max(64u * g, 16u); p < min(64u * g + 64u, min(lim, own_end)); p++)
                        if (buf[p] >= 0x80u) { report_error(T, pt::PT_U_NON_ASCII, base_off + (int64_t)p); break; }
                }
            }
            // record starts: one candidate per lane and round (a second round only when a 64-byte group holds two)
            while (__any_sync(FULL, cand != 0ull)) {
                bool is_start = false;
                uint32_t p = 0;
                if (cand != 0ull) {
                    p = 64u * g + (uint32_t)(__ffsll((long long)cand) - 1);
                    cand &= cand - 1ull;
                    const uint32_t c = buf[p];
                    if (c == '\n') {
                        is_start = p + 1u < own_end;
                    } else if (c == '\r' && p >= 16u && p < own_end) {      // lone '\r': a line break for the reference's text mode
                        const uint64_t abs_pos = t0 + p - 16u;
                        if (abs_pos + 1 < A.nbytes && buf[p + 1] != '\n') report_error(T, pt::PT_U_BARE_CR, base_off + (int64_t)p);
                    }
                }
                const uint32_t bal = __ballot_sync(FULL, is_start);
                if (is_start) {
                    const uint32_t j = my_cnt + (uint32_t)__popc(bal & lt_mask);
                    if (j < (uint32_t)G::LINE_CAP) lines[warp ? (uint32_t)G::LINE_CAP - 1u - j : j] = (uint16_t)(p + 1u);
                }
                my_cnt += (uint32_t)__popc(bal);
            }
        }
        if (lane == 0) s_cnt[warp] = my_cnt;
        __syncthreads();                                                    // ---- B1: masks + record lists complete; everyone is through the previous tile
        const uint32_t n0 = s_cnt[0], n1 = s_cnt[1];
        const uint32_t lwm_rel = s_lwm;
        if (tid == 0) {
            my_lines += n0 + n1;
            my_tiles++;
            *(volatile uint32_t*)&T.team_tile[blockIdx.x] = tile;           // my tiles before this one are complete
        }
        if (blockIdx.x == 0 && warp == 1) {
            // one team keeps the kernel's low-water mark: the smallest tile any team is still working on
            uint32_t m = 0xFFFFFFFFu;
            for (uint32_t i = lane; i < gridDim.x; i += 32u) m = min(m, *(volatile uint32_t*)&T.team_tile[i]);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) m = min(m, __shfl_xor_sync(FULL, m, o));
            if (lane == 0 && m != 0xFFFFFFFFu) atomicMax(&T.sc[SC_LWM], (unsigned long long)m);
        }
        if (ablate == 2u) {
            __syncthreads();
            if (tid == 0 && nxt_tile < A.n_tiles) issue_load(nxt_tile);
            continue;
        }

        if (n0 + n1 > (uint32_t)G::LINE_CAP) {
            // more records than the lists hold (pathological input): all of them take the exact path
            for (uint32_t p = 15u + tid; p + 1u < own_end; p += THREADS) {
                const bool nl = p == 15u ? (tile == 0 || buf[p] == '\n') : buf[p] == '\n';
                if (nl) defer_line(T, t0 + p + 1u - 16u, A.file_off, WHY_LINES_FULL);
            }
            if (warp == 1) drain_far();                                    // the previous tile's far links (normally done in `walk`)
            __syncthreads();
            if (tid == 0) {
                s_nfar = 0;
                s_far_take = 0;
                if (nxt_tile < A.n_tiles) issue_load(nxt_tile);
            }
            __syncthreads();
            continue;
        }
        const uint32_t n_lines = n0 + n1;

        // ================= records: warp 0 = role B, warp 1 = role A, one thread per record =================
        if (warp == 0) {
            uint32_t step_base = 0;                                         // warp-uniform: entries handed out so far
            for (uint32_t l0 = 0; l0 < n_lines; l0 += 32u) {
                const uint32_t l = l0 + lane;
                const bool have = l < n_lines;
                uint32_t st = ST_DONE, ns = 0, a5 = 0, b5 = 0, ls = 0;
                int why = WHY_LONG;
                int32_t plen = 0, start = 0, pend = 0;
                if (have) {
                    // ---------------- role B: columns, filters, coordinates
                    ls = lines[l < n0 ? l : (uint32_t)G::LINE_CAP - 1u - (l - n0)];
                    uint32_t wi = ls >> 5;
                    uint32_t wmk = wm32[wi] & (~0u << (ls & 31u));
                    uint32_t e[13];
                    e[0] = ls - 1u;
                    bool ran_off = false, gaps_ok = true;
                    uint32_t tabs = 0xFFFFFFFFu;              // AND of (byte == '\t') over the first 11 boundaries
#pragma unroll
                    for (int j = 1; j <= 12; j++) {
                        e[j] = 0;
                        if (!ran_off) {
                            if (!next_ws(wm32, 2u * nwords, wi, wmk, e[j])) ran_off = true;   // record runs past the look-ahead
                            else {
                                gaps_ok &= e[j] - e[j - 1] >= 2u;                        // no empty column
                                if (j < 12) tabs &= buf[e[j]] == '\t' ? 0xFFFFFFFFu : 0u;
                            }
                        }
                    }
                    bool slow = ran_off, done = false, no_tags = false;
                    int32_t mapq = 0;
                    if (!slow) {
                        // 11 single tabs, then a tab (tags follow) or the end of a 12-column record
                        const uint32_t c12 = buf[e[12]];
                        no_tags = c12 == '\n';
                        if (!gaps_ok || tabs == 0u || (c12 != '\t' && !no_tags)) { slow = true; why = WHY_COLUMNS; }
                    }
                    if (!slow) {
                        why = WHY_INTS;
                        slow = !small_uint(buf, e[11] + 1u, e[12], mapq);
                        if (!slow) {
                            if ((int64_t)mapq < A.thr) { sink.reject(); done = true; }                   // REF:143-146
                            else if (e[6] - e[5] == 2u && buf[e[5] + 1u] == '*') done = true;             // REF:147-148
                        }
                    }
                    if (!slow && !done)
                        slow = !small_uint(buf, e[6] + 1u, e[7], plen) || !small_uint(buf, e[7] + 1u, e[8], start) ||
                               !small_uint(buf, e[8] + 1u, e[9], pend);
                    if (!slow && !done && no_tags) { slow = true; why = WHY_TAGS; }   // no dv tag: ValueError (REF:179), the exact path reports
                    // ---- path column (REF:185-197): it must start with a separator; count the steps
                    if (!slow && !done) {
                        why = WHY_PATH;
                        a5 = e[5] + 1u;
                        b5 = e[6];
                        for (uint32_t w = a5 >> 5; w <= ((b5 - 1u) >> 5); w++) ns += (uint32_t)__popc(sep_word(sm32, w, a5, b5));
                        if (ns == 0u || ns > (uint32_t)MAX_STEPS || !((sm32[a5 >> 5] >> (a5 & 31u)) & 1u)) slow = true;
                    }
                    st = slow ? ST_DEFER : (done ? ST_DONE : ST_FAST);
                    if (st != ST_FAST) ns = 0;
                }
                // ---- list space for the steps of the warp's records (+ one sentinel each)
                uint32_t total;
                uint32_t off = step_base + warp_excl_scan(st == ST_FAST ? ns + 1u : 0u, lane, total);
                step_base += total;
                if (have) {
                    Rec& R = recs[l];
                    if (st == ST_FAST && off + ns + 1u > (uint32_t)G::STEP_CAP) {      // list full: exact path
                        st = ST_DEFER;
                        why = WHY_STEPS_FULL;
                        for (uint32_t i = off; i < (uint32_t)G::STEP_CAP; i++) steps[i] = SE_INVALID;
                    }
                    R.ls = (uint16_t)ls;
                    R.stB = (uint8_t)st;
                    R.whyB = (uint8_t)why;
                    R.nsteps = 0;
                    R.s0 = 0;
                    if (st == ST_FAST) {
                        R.start = start;
                        R.end_rel1 = plen - pend - 1;
                        R.s0 = (uint16_t)off;
                        R.nsteps = (uint16_t)ns;
                        // ---- one entry per path step, then the sentinel (end of the column)
                        const uint32_t common = (l << SE_SLOT_SHIFT) | (buf[a5] == '<' ? SE_REV : 0u) | (1u << SE_NCNT_SHIFT);
                        uint32_t i = off;
                        for (uint32_t w = a5 >> 5; w <= ((b5 - 1u) >> 5); w++) {
                            uint32_t m = sep_word(sm32, w, a5, b5);
                            const uint32_t wb = (32u * w) | common;
                            while (m) {
                                steps[i++] = wb + (uint32_t)(__ffs((int)m) - 1);
                                m &= m - 1u;
                            }
                        }
                        steps[off] |= SE_FIRST;
                        steps[off + ns - 1u] |= SE_LAST;
                        steps[off + ns] = b5 | (l << SE_SLOT_SHIFT) | SE_SENT;
                    }
                }
            }
            if (lane == 0) s_nent = step_base;
        } else {
            uint32_t ops_base = 0;                                          // warp-uniform: op-pool words handed out so far
            for (uint32_t l0 = 0; l0 < n_lines; l0 += 32u) {
                const uint32_t l = l0 + lane;
                const bool have = l < n_lines;
                // ---------------- role A, part 1: tags -> dv filter, where the cs string is
                uint32_t st = ST_DONE;
                int why = WHY_LONG;
                bool idle = true, slow = false, done = false, perfect = false;
                uint32_t cs_a = 0, cs_b = 0, room = 0, q = 0, n_tot = 0;
                if (have) {
                    const uint32_t ls = lines[l < n0 ? l : (uint32_t)G::LINE_CAP - 1u - (l - n0)];
                    uint32_t wi = ls >> 5;
                    uint32_t wmk = wm32[wi] & (~0u << (ls & 31u));
                    uint32_t e11 = 0, e12 = 0;
                    bool ran_off = false;
                    // the first ten column boundaries are role B's business: skip them a half word at a time
                    uint32_t skip = 10;
                    for (;;) {
                        const uint32_t c = (uint32_t)__popc(wmk);
                        if (c > skip) break;
                        skip -= c;
                        if (++wi >= 2u * nwords) { ran_off = true; break; }
                        wmk = wm32[wi];
                    }
                    for (; skip != 0u && !ran_off; skip--) wmk &= wmk - 1u;
                    if (!ran_off && !next_ws(wm32, 2u * nwords, wi, wmk, e11)) ran_off = true;
                    if (!ran_off && !next_ws(wm32, 2u * nwords, wi, wmk, e12)) ran_off = true;
                    // role B decides about everything up to column 12; here: is there anything left to do?
                    int32_t mapq = 0;
                    idle = ran_off || buf[e12] != '\t' || !small_uint(buf, e11 + 1u, e12, mapq) || (int64_t)mapq < A.thr;
                    uint32_t dv_a = 0, dv_b = 0, dv3 = 0;
                    unsigned long long cs8 = 0;
                    if (!idle) {
                        // ---- tags: [inert]* cs [inert]* dv in any order, within the first few tags
                        why = WHY_TAGS;
                        uint32_t a = e12 + 1u, b = 0;
                        if (!next_ws(wm32, 2u * nwords, wi, wmk, b)) slow = true;
                        for (int j = 13; !slow; j++) {
                            const unsigned long long t8 = ld8(buf, a);          // the token's first eight bytes
                            if (!cs_b && b - a >= 3u && (t8 & 0xFFFFFFull) == TAG_CS3) {
                                cs_a = a;
                                cs_b = b;
                                cs8 = t8;
                            } else if (!dv_b && b - a >= 6u && (t8 & 0xFFFFFFFFFFull) == TAG_DV5 && pt::is_digit((uint32_t)(t8 >> 40) & 0xFFu) &&
                                       no_colon_w(buf, a + 5u, b)) {
                                dv_a = a + 5u;
                                dv_b = b;
                                dv3 = (uint32_t)(t8 >> 40);                     // the first three bytes of the number
                            } else if (b - a >= 6u && (t8 & 0xFFFFFFFFFFull) == TAG_AS5) {
                                // "AS:i:<int>", the tag the aligner writes first: inert when nothing after the prefix is a ':'
                                const bool ok = b - a <= 8u ? !has_colon8(t8 >> 40, b - a - 5u) : no_colon_w(buf, a + 5u, b);
                                if (!ok) { slow = true; break; }
                            } else if (!token_is_inert(buf, a, b)) {
                                slow = true;
                                break;
                            }
                            if (cs_b && dv_b) break;
                            if (buf[b] == '\n' || j >= 18) { slow = true; break; }      // end of the record: a tag is missing
                            a = b + 1u;
                            if (!next_ws(wm32, 2u * nwords, wi, wmk, b)) { slow = true; break; }
                        }
                        // ---- dv filter (REF:172-180).  The reference parses cs first, but that has no side effects and
                        //      cannot raise, so a record that dv filters out needs no cs class
                        if (!slow) {
                            const uint32_t f = dv3 & 0xFFu, g = dv_a + 1u < dv_b ? (dv3 >> 8) & 0xFFu : 0u, h = dv_a + 2u < dv_b ? (dv3 >> 16) & 0xFFu : 0u;
                            if (f == '0' && g == '.' && h == '0') {
                                // 0.0xxx: never greater
                            } else if (pt::dv_token_greater(buf, (int)dv_a, (int)dv_b)) {
                                done = true;
                            }
                        }
                        // ---- cs string (REF:10-37): "cs:Z:" then ops spelled the way an aligner spells them
                        if (!slow && !done) {
                            why = WHY_CS;
                            if (cs_b - cs_a < 7u || (cs8 & 0xFFFF000000ull) != 0x3A5A000000ull) slow = true;      // "cs:Z:"
                            q = cs_a + 5u;
                            uint32_t one;
                            if (!slow && ((cs8 >> 40) & 0xFFu) == ':' && cs_b - q - 1u <= 7u && dec8(buf, q + 1u, cs_b - q - 1u, one) && one != 0u &&
                                one <= (uint32_t)MAX_NTOT) {
                                perfect = true;                                 // cs:Z::<n> -- a perfect match
                                n_tot = one;
                                room = 1;
                            } else if (!slow) {
                                // every op takes at least two bytes: room for (bytes / 2) ops is enough
                                room = min((cs_b - q) >> 1, (uint32_t)MAX_OPS);
                            }
                        }
                    }
                }
                const bool want = have && !idle && !slow && !done;
                // ---- op-pool space for the warp's records
                uint32_t total;
                const uint32_t op_off = ops_base + warp_excl_scan(want ? room : 0u, lane, total);
                ops_base += total;
                if (have && !idle) {
                    Rec& R = recs[l];
                    uint32_t nops = 0;
                    int32_t start_add = 0;
                    if (want) {
                        if (op_off + room > (uint32_t)G::OPS_CAP) {
                            slow = true;
                        } else if (perfect) {
                            ops[op_off] = OP_MATCH | (n_tot << 3);
                            nops = 1;
                        } else {
                            // ---------------- role A, part 2: the general cs string
                            //   ':'<digits>  '*'<2 letters>  '-'<letters>  '+'<letters>  '='<LETTERS>, every length >= 1
                            while (!slow && q < cs_b) {
                                const uint32_t c = buf[q++];
                                uint32_t kind, len = 0;
                                if (c == ':') {
                                    kind = OP_MATCH;
                                    uint32_t nd = 0;
                                    while (q < cs_b && pt::is_digit(buf[q])) { len = len * 10u + (buf[q] - '0'); q++; nd++; }
                                    if (nd == 0u || nd > 7u) slow = true;
                                } else if (c == '*') {
                                    kind = OP_SUB;
                                    if (q + 2u > cs_b || !is_lower(buf[q]) || !is_lower(buf[q + 1])) slow = true;
                                    q += 2u;
                                    len = 1;
                                } else if (c == '-' || c == '+') {
                                    kind = c == '-' ? OP_DEL : OP_INS;
                                    while (q < cs_b && is_lower(buf[q])) { q++; len++; }
                                } else if (c == '=') {
                                    kind = OP_EQ;
                                    while (q < cs_b && (uint32_t)buf[q] - 'A' <= 24u) { q++; len++; }      // 'A'..'Y': "cs:Z:" cannot hide in here
                                } else {
                                    slow = true;
                                    kind = 0;
                                }
                                // the text must end where the next op starts
                                if (q < cs_b) {
                                    const uint32_t d = buf[q];
                                    if (d != ':' && d != '*' && d != '-' && d != '+' && d != '=') slow = true;
                                }
                                if (len == 0u || len > (uint32_t)MAX_NTOT || nops >= room) slow = true;
                                if (!slow) {
                                    ops[op_off + nops] = kind | (len << 3);
                                    nops++;
                                    n_tot += len;
                                    if (n_tot > (uint32_t)MAX_NTOT) slow = true;
                                }
                            }
                            if (nops == 0u) slow = true;
                            // cigar_clipping (REF:40-50): only when there are exactly two ops
                            if (!slow && nops == 2u) {
                                const uint32_t o0 = ops[op_off], o1 = ops[op_off + 1u];
                                if ((o0 & 7u) == OP_INS && (o1 & 7u) == OP_MATCH) {
                                    start_add = (int32_t)(o0 >> 3);
                                    ops[op_off] = o1;
                                    nops = 1;
                                    n_tot = o1 >> 3;
                                } else if ((o0 & 7u) == OP_MATCH && (o1 & 7u) == OP_INS) {
                                    nops = 1;
                                    n_tot = o0 >> 3;
                                }
                            }
                        }
                        if (!slow) {
                            const uint32_t k0 = ops[op_off] & 7u;
                            R.n_tot = n_tot;
                            R.op_off = (uint16_t)op_off;
                            R.start_add = start_add;
                            R.single = (uint8_t)((nops == 1u && (k0 == OP_MATCH || k0 == OP_EQ)) ? 1 : 0);
                        }
                    }
                    st = slow ? ST_DEFER : (done ? ST_DONE : ST_FAST);
                    // nops | stA | (stB: role B's byte) | whyA -- three byte stores, stB is not touched
                    R.nops = (uint8_t)nops;
                    R.stA = (uint8_t)st;
                    R.whyA = (uint8_t)why;
                } else if (have) {
                    Rec& R = recs[l];
                    R.nops = 0;
                    R.stA = ST_DONE;
                    R.whyA = (uint8_t)why;
                }
            }
        }
        __syncthreads();                                                    // ---- B2: records, ops, step list complete
        const uint32_t n_ent = min(s_nent, (uint32_t)G::STEP_CAP);         // step entries incl. sentinels
        if (ablate == 3u) {
            __syncthreads();
            if (tid == 0 && nxt_tile < A.n_tiles) issue_load(nxt_tile);
            continue;
        }

        // ================= ids: one thread per path step: id -> node index -> the node's hot record =================
        // UI steps per thread and iteration: their 16-byte loads are all in flight before the first is used.
        {
            constexpr int UI = 4;
            for (uint32_t s00 = 0; s00 < n_ent; s00 += THREADS * UI) {
                uint32_t idx_[UI], se_[UI];
                uint4 hot_[UI];
#pragma unroll
                for (int u = 0; u < UI; u++) {
                    const uint32_t s = s00 + THREADS * u + tid;
                    uint32_t idx = NONE32, se = SE_INVALID;
                    if (s < n_ent) {
                        se = steps[s];
                        if (se != SE_INVALID && !(se & SE_SENT)) {
                            const uint32_t p = se & SE_POS_MASK;
                            const uint32_t end = steps[s + 1u] & SE_POS_MASK;       // next separator, or the sentinel
                            uint64_t id;
                            uint32_t ix;
                            // the separator the path began with (REF:186-194: a mixed path is a KeyError)
                            if (buf[p] == ((se & SE_REV) ? '<' : '>') && step_id(buf, p + 1u, end - p - 1u, id) && sink.id_to_idx(id, ix)) idx = ix;
                        }
                    }
                    idx_[u] = idx;                                          // NONE32: KeyError in the reference, `walk` hands the record over
                    se_[u] = se;
                    hot_[u] = make_uint4(0u, 0u, 0u, 0u);
                    if (idx != NONE32) hot_[u] = sink.load_hot(idx);        // issued at once: in flight while the next id is parsed
                }
#pragma unroll
                for (int u = 0; u < UI; u++) {
                    const uint32_t s = s00 + THREADS * u + tid;
                    if (s < n_ent) {
                        const uint32_t se = se_[u], meta = hot_[u].x, len = meta & META_LEN_MASK;
                        uint32_t share = SL_BAD;                            // absent node, >= 1023 bases, unknown id: exact path
                        if (idx_[u] != NONE32 && len - 1u < META_LEN_ESC - 1u) {
                            int32_t L = (int32_t)len;
                            if (se & (SE_FIRST | SE_LAST)) {
                                const Rec& R = recs[(se >> SE_SLOT_SHIFT) & SE_SLOT_MASK];
                                int64_t L64 = L;
                                if (se & SE_FIRST) L64 -= (int64_t)R.start + R.start_add;           // REF:215-216
                                if (se & SE_LAST) L64 -= R.end_rel1;                                // REF:217-218
                                L = L64 <= 0 ? 0 : (L64 >= (int64_t)SL_BAD ? (int32_t)SL_BAD : (int32_t)L64);
                            }
                            share = (uint32_t)L;
                        }
                        sidx[s] = idx_[u];
                        smeta[s] = meta;
                        sL[s] = (uint16_t)share;
                    }
                }
            }
        }
        __syncthreads();                                                    // ---- B3: node indices complete; the bytes and the masks are dead
        if (tid == 0 && nxt_tile < A.n_tiles) issue_load(nxt_tile);        // overlaps walk + fold + count
        if (ablate == 4u) continue;

        // ================= walk: warp 0, one thread per record; warp 1 drains the previous tile's far links =================
        if (warp == 0) {
            uint32_t hbase = 0;                                             // warp-uniform: prefix-pool entries handed out
            for (uint32_t l0 = 0; l0 < n_lines; l0 += 32u) {
                const uint32_t l = l0 + lane;
                bool fast = false, multi = false;
                uint32_t ns = 0, s0 = 0;
                if (l < n_lines) {
                    const Rec& R = recs[l];
                    fast = rec_status(R) == ST_FAST;
                    if (fast) {
                        ns = R.nsteps;
                        s0 = R.s0;
                        multi = R.single == 0;
                    }
                }
                uint32_t total;
                const uint32_t a_off = hbase + warp_excl_scan(multi ? ns : 0u, lane, total);
                hbase += total;
                if (fast) {
                    Rec& R = recs[l];
                    bool bad = false;
                    if (multi && a_off + ns > (uint32_t)G::HEAVY_CAP) {         // prefix pool full: exact path
                        for (uint32_t h = a_off; h < (uint32_t)G::HEAVY_CAP; h++) heavy[h] = 0xFFFFu;
                        bad = true;
                        multi = false;
                        ns = 0;
                    }
                    uint32_t run = 0, a_last = 0, prev = NONE32;
                    bool any = false;
                    for (uint32_t k = 0; k < ns; k++) {
                        const uint32_t v = sL[s0 + k], i = sidx[s0 + k];
                        // unknown id: KeyError (REF:214); collapsible duplicate (REF:188): the exact path redoes the record
                        bad |= v == SL_BAD || i == prev;
                        prev = i;
                        if (multi) {
                            sA[a_off + k] = run;
                            heavy[a_off + k] = (uint16_t)(s0 + k);
                        }
                        if (v != 0u) { a_last = run; any = true; }
                        run += v;
                    }
                    // a node with bases left but no cs left: IndexError (REF:227); the last such node starts furthest right
                    if (any && a_last >= R.n_tot) bad = true;
                    if (bad) {
                        R.stB = ST_DEFER;
                        R.whyB = WHY_WALK;
                    } else if (!multi && ns) {
                        // one ':' / '=' op: a node is not in `align` iff no bases are left for it -- only the two ends can be
                        if (sL[s0] == 0u) steps[s0] |= SE_DROPPED;
                        if (sL[s0 + ns - 1u] == 0u) steps[s0 + ns - 1u] |= SE_DROPPED;
                    }
                    R.a_off = (uint16_t)a_off;
                }
            }
            if (lane == 0) s_nheavy = min(hbase, (uint32_t)G::HEAVY_CAP);
        } else {
            drain_far();
            __syncwarp();
            if (lane == 0) {
                s_nfar = 0;
                s_far_take = 0;
                s_ndel = 0;
            }
        }
        far_base = base_off;                                                // the list `count` fills below belongs to this tile
        __syncthreads();                                                    // ---- B4: hand-over decisions of `walk`, prefix pool, heavy list
        if (ablate == 5u) continue;

        // ================= fold: every step of a multi-op record folds the cs ops that overlap its node =================
        const uint32_t n_heavy = s_nheavy;
        if (n_heavy != 0u) {
            for (uint32_t h = tid; h < n_heavy; h += THREADS) {
                const uint32_t s = heavy[h];
                if (s == 0xFFFFu) continue;                                 // (pool overflow: that record went to the exact path)
                const uint32_t se = steps[s];
                Rec& R = recs[(se >> SE_SLOT_SHIFT) & SE_SLOT_MASK];
                if (rec_status(R) != ST_FAST) continue;
                const uint32_t Lk = sL[s];
                if (Lk == 0u) {                                             // no bases left for this node: it is not in `align`
                    steps[s] = se | SE_DROPPED;
                    continue;
                }
                const uint32_t Ak = sA[h], n_tot = R.n_tot;                 // Ak < n_tot (walk)
                const uint32_t* op = ops + R.op_off;
                const uint32_t nops = R.nops;
                const uint32_t Bk = min(Ak + Lk, n_tot);
                // pieces of the node = ops overlapping [Ak, Bk), clipped; compact_align as a running fold (REF:63-94)
                uint32_t j = 0, o_start = 0, o_end = op[0] >> 3;
                while (o_end <= Ak) {                                       // ends before the node starts (j < nops: Ak < n_tot)
                    j++;
                    o_start = o_end;
                    o_end += op[j] >> 3;
                }
                uint32_t nP = 0, nQ = 0, p0 = 0, qlast_op = 0, qlast_len = 0, first_op = 0, first_len = 0, n_count = 0;
                for (;;) {
                    const uint32_t kind = op[j] & 7u;
                    const uint32_t take = min(o_end, Bk) - max(o_start, Ak);
                    bool push = false;
                    uint32_t push_len = take;
                    if (nP == 0u) { p0 = kind; push = kind != OP_SUB; }
                    else if (nQ == 0u) { push = true; push_len = take + 1u; }
                    else if (kind == qlast_op || kind == OP_SUB) qlast_len += take;
                    else push = true;
                    if (push) {
                        if (nQ == 1u) { first_op = qlast_op; first_len = qlast_len; }
                        qlast_op = kind;
                        qlast_len = push_len;
                        nQ++;
                        if (kind != OP_DEL && kind != OP_SUB) n_count++;
                    }
                    nP++;
                    if (o_end >= Bk || j + 1u >= nops) break;
                    j++;
                    o_start = o_end;
                    o_end += op[j] >> 3;
                }
                if (nQ == 1u) { first_op = qlast_op; first_len = qlast_len; }
                if (nP == 1u && (p0 == OP_DEL || p0 == OP_INS)) {           // clear_align drops the node (REF:101-102)
                    steps[s] = se | SE_DROPPED;
                    continue;
                }
                if (n_count > 3u) {                                         // more counting ops than the entry holds: exact path
                    R.stB = ST_DEFER;
                    R.whyB = WHY_WALK;
                    continue;
                }
                steps[s] = (se & ~SE_NCNT_MASK) | (n_count << SE_NCNT_SHIFT);
                const bool first_del = nQ > 0u && first_op == OP_DEL, last_del = nQ > 0u && qlast_op == OP_DEL;
                if (first_del || last_del) {                                // deletion-derived IL/OL keys (REF:281-297,317-333)
                    const uint32_t k = atomicAdd(&s_ndel, 1u);
                    if (k < (uint32_t)G::DEL_CAP) {
                        dels[3u * k] = s;
                        dels[3u * k + 1u] = first_len | (first_del ? 0x80000000u : 0u);
                        dels[3u * k + 2u] = qlast_len | (last_del ? 0x80000000u : 0u);
                    } else {
                        R.stB = ST_DEFER;                                   // list full: exact path (nothing counted yet)
                        R.whyB = WHY_WALK;
                    }
                }
            }
            __syncthreads();                                                // ---- B5: every hand-over decision is made; nothing counted so far
        }
        for (uint32_t l = tid; l < n_lines; l += THREADS) {
            const Rec& R = recs[l];
            if (rec_status(R) == ST_DEFER) defer_line(T, t0 + R.ls - 16u, A.file_off, R.stB == ST_DEFER ? R.whyB : R.whyA);
        }

        // surviving neighbours of step s inside its record (dropped nodes are skipped)
        auto prev_survivor = [&](uint32_t s) -> uint32_t {                  // NONE32: s is the first survivor
            uint32_t t = s;
            while (!(steps[t] & SE_FIRST)) {
                t--;
                if (!(steps[t] & SE_DROPPED)) return t;
            }
            return NONE32;
        };
        auto next_survivor = [&](uint32_t s) -> uint32_t {                  // NONE32: s is the last survivor
            uint32_t t = s;
            while (!(steps[t] & SE_LAST)) {
                t++;
                if (!(steps[t] & SE_DROPPED)) return t;
            }
            return NONE32;
        };

        // ================= count: one thread per surviving step (REF:263-363) =================
        const int64_t rel_base = base_off + 1 - T.epoch_base;               // epoch-relative offset of buffer position -1 + ...
        for (uint32_t s = tid; s < n_ent; s += THREADS) {
            const uint32_t se = steps[s];
            if (se == SE_INVALID || (se & (SE_SENT | SE_DROPPED)) || rec_status(recs[(se >> SE_SLOT_SHIFT) & SE_SLOT_MASK]) != ST_FAST) continue;
            const uint32_t idx = sidx[s], meta = smeta[s];
            const bool rev = (se & SE_REV) != 0u;
            const uint32_t ps = prev_survivor(s), nx = next_survivor(s);
            const bool first = ps == NONE32, last = nx == NONE32;           // among the surviving nodes (REF:276-353: i == 0, i == last)
            const uint32_t n_count = (se >> SE_NCNT_SHIFT) & 3u;
            // this step owns the link that LEAVES its node (REF:357-359): to the next survivor forward, to the previous one reverse;
            // the link that enters it is the neighbour's
            const bool has_out = rev ? !first : !last, has_in = rev ? !last : !first;
            int slot = -1;
            if (has_out) {
                const uint32_t other = sidx[rev ? ps : nx];
                slot = DevSink::inline_slot(meta, idx, other);
                if (slot < 0) {
                    // not inline: hash-table work, listed and done during the next tile's walk phase.  Stamped like
                    // the reference's insertion: when the later of the two steps is reached
                    const uint32_t ep = rev ? (se & SE_POS_MASK) : (steps[nx] & SE_POS_MASK);
                    const uint32_t j = atomicAdd(&s_nfar, 1u);
                    if (j < (uint32_t)G::FAR_CAP) {
                        far[3u * j] = idx;
                        far[3u * j + 1u] = other;
                        far[3u * j + 2u] = ep;
                    } else {
                        sink.edge_far(idx, other, (uint64_t)(base_off + (int64_t)ep + 1) << 2);
                    }
                }
            }
            if (!has_out || slot >= 0) sink.bump(idx, slot);                                          // REF:263-269, 357-363
            if (n_count != 1u) sink.extras(idx, has_in ? (int32_t)n_count - 1 : 0, has_out ? (int32_t)n_count - 1 : 0);   // REF:298-351
            if (n_count != 0u) {
                const bool need_il = has_in && !(meta & META_IL_SETTLED), need_ol = has_out && !(meta & META_OL_SETTLED);
                if (need_il || need_ol)
                    sink.touch_stamps(idx, need_il, need_ol, (uint32_t)(rel_base + (int64_t)(se & SE_POS_MASK)), lwm_rel);
            }
        }
        // no barrier: the list of deletion keys is complete since `fold`, the far-link list is read one tile later
        {
            // deletion-derived keys: position of the deletion inside the node, IL or OL by orientation
            const uint32_t n_del = min(s_ndel, (uint32_t)G::DEL_CAP);
            for (uint32_t j = tid; j < n_del; j += THREADS) {
                const uint32_t s = dels[3u * j], f = dels[3u * j + 1u], g = dels[3u * j + 2u];
                const uint32_t se = steps[s];
                if (rec_status(recs[(se >> SE_SLOT_SHIFT) & SE_SLOT_MASK]) != ST_FAST) continue;
                const uint32_t idx = sidx[s];
                const int64_t len = (int64_t)(smeta[s] & META_LEN_MASK);
                const bool rev = (se & SE_REV) != 0u;
                const bool not_first = prev_survivor(s) != NONE32, not_last = next_survivor(s) != NONE32;
                const bool first_del = (f >> 31) != 0u, last_del = (g >> 31) != 0u;
                const int64_t first_len = f & 0x7FFFFFFFu, last_len = g & 0x7FFFFFFFu;     // >= 1: no op is empty here
                const uint64_t stamp = (uint64_t)(base_off + (int64_t)(se & SE_POS_MASK) + 1) << 2;
                if (!rev) {
                    if (first_del && not_first) sink.sparse(idx, 0, first_len, stamp | 0u);               // REF:282-289
                    if (last_del && not_last) sink.sparse(idx, 1, len - last_len - 1, stamp | 2u);         // REF:290-297
                } else {
                    if (first_del && not_first) sink.sparse(idx, 1, len - 1 - first_len, stamp | 0u);      // REF:318-325
                    if (last_del && not_last) sink.sparse(idx, 0, last_len, stamp | 2u);                   // REF:326-333
                }
            }
        }
        // No barrier at the end of the tile: a thread that is done goes on to scan the next tile (its bytes arrived long ago).
        // The scan writes the masks (= sL / prefix pool / heavy list, dead since `fold`) and the record-start lists; nothing a
        // straggler of this tile still reads, and every thread passes the scan's barrier only after it is through here.
    }
    __syncthreads();                                                        // the last tile's far-link list is complete

    drain_far();
    if (tid == 0) *(volatile uint32_t*)&T.team_tile[blockIdx.x] = 0xFFFFFFFFu;   // nothing of mine is pending any more

    // rejected-record count: warp reduce, one RED per warp
    uint32_t r = sink.rej;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(FULL, r, o);
    if (lane == 0u && r) atomicAdd(&T.sc[SC_REJ], (unsigned long long)r);
    if (tid == 0) {
        if (my_lines) atomicAdd(&T.sc[SC_LINES], my_lines);
        if (my_tiles) atomicAdd(&T.sc[SC_TILES], my_tiles);
    }
}

}  // namespace teamp
