// fast_tiles.cuh -- the fast path of the augment kernel (included by aug_kernels.cuh after tables.cuh, the TMA
// helpers, ChunkArgs and defer_line(); compiled for sm_100a by pantas_aug.cu and for the CPU emulator by
// tests/hostsim/fastsim.cpp).
//
// Reference loop body: /root/reference/scripts/alignments_augmentation_from_gaf.py:142-363 (REF:n).
//
// A persistent CTA takes tiles of the GAF chunk (TILE bytes + OV bytes of look-ahead, one 1-D TMA
// bulk copy, UBLKCP, L2 evict-first) and runs phases over the tile in shared memory:
//
//   scan     one thread per 64 bytes (4 x LDS.128 in a lane-rotated order: no bank conflicts), branch-free
//            SWAR: a 64-bit whitespace mask and a 64-bit path-separator ('>' '<') mask per group; record
//            starts ('\n') go to a list; lone '\r' and non-ASCII bytes are (fatal) errors.
//   records  two roles per record, whole warps per role, the records packed into as few warps as they need
//            (a warp takes as long for one record as for 32); both walk the whitespace mask as 32-bit halves:
//              B  the 12 column boundaries (single tabs, no empty column), MAPQ and '*' filters
//                 (REF:143-148), the three coordinates (REF:151-153), then the separator mask of
//                 the path column -> one entry per path step (+ a sentinel) in the tile's step list;
//              A  skips ten boundaries by popcount; the tags, classified on whole words: first cs token,
//                 first dv:f: token (REF:154-160,172-180), the dv filter, and the cs string parsed into the
//                 tile's op pool (REF:10-50, incl. cigar_clipping).
//            Anything unusual -- other whitespace in the columns, integers that are not plain
//            digits, tags that could confuse the reference's regexes, '~' or zero-length or oddly
//            spelled cs ops -- hands the record to the exact per-record path (line_core.cuh via
//            augment_deferred_kernel).  The warps this phase leaves idle drain the PREVIOUS tile's list of
//            links that are not inline (hash-table probes: their latency costs nothing here).
//   ids      one thread per path step: SWAR decimal parse of the id out of shared memory, node index, one
//            16-byte load of the node record's read half, issued as soon as the index is known; length,
//            inline link deltas and two stamp flags stay in shared memory.  After this phase the bytes are
//            dead and the next tile's TMA copy is issued: it overlaps the remaining phases.
//   walk     one thread per path step.  The reference's merge walk (REF:205-255) gives node k the
//            ops that overlap [A_k, A_k + L_k) in cs coordinates, where A is the prefix sum of the
//            node lengths L (first / last node shortened, REF:215-218): a block-wide prefix sum (walk 1),
//            then every step finds its op pieces independently and folds clear_align /
//            compact_align (REF:63-107) over them: dropped or not, number of counting ops,
//            deletion-derived IL/OL keys (walk 2: steps of single-op records at once, the others listed
//            and folded in a dense second pass).  Collapsible duplicate ids (REF:188), unknown ids and a
//            cs string shorter than the path hand the record over.  Nothing has been counted yet,
//            so the hand-over is clean.
//   count    one thread per surviving step, no global loads: NC / IL / OL / RC events (REF:263-363): one
//            RED.ADD.64 on the node's sector (tables.cuh), RED.MIN for first-touch stamps only when this
//            tile could lower them.  Links that are not inline go to the list the next records phase
//            drains; deletion-derived keys follow at once.  No barrier closes the tile.
#pragma once

namespace fastp {

constexpr uint32_t NONE32 = 0xffffffffu;
// sidx[] entry: node index in the low 30 bits (larger indices: slow path), two flags on top: could this tile still lower the
// node's IL / OL first-touch stamp?  (NONE32: no such node)
constexpr uint32_t SX_IDX_MASK = 0x3FFFFFFFu, SX_NEED_IL = 1u << 30, SX_NEED_OL = 1u << 31;
constexpr int MAX_STEPS = 250;                // longer paths take the slow path
constexpr int MAX_OPS = 48;                   // more cs ops: slow path
constexpr uint32_t L_CLAMP = 1u << 23;        // step lengths are clamped here (> any cs length the fast path takes)
constexpr int32_t MAX_NTOT = 1 << 22;

enum : uint8_t { ST_FAST = 0, ST_DONE = 1, ST_DEFER = 2 };      // per role; a record's status is the maximum
enum : uint32_t { OP_MATCH = 0, OP_SUB = 1, OP_DEL = 2, OP_INS = 3, OP_EQ = 4 };   // ':' '*' '-' '+' '='  (op = kind | len << 3)

struct __align__(4) LineRecF {
    int32_t start;        // int(tokens[7])                                    (role B)
    int32_t end_rel1;     // int(tokens[6]) - int(tokens[8]) - 1               (role B)
    int32_t start_add;    // cigar_clipping: start_pos += len of a leading '+' (role A, REF:46-47)
    uint32_t n_tot;       // sum of the cs op lengths                          (role A)
    uint32_t base;        // step-length prefix at the record's first step     (walk)
    uint16_t s0;          // first entry of the record in the step list        (role B)
    uint16_t nsteps;      //                                                   (role B)
    uint16_t ls;          // buffer position of the record's first byte        (role B)
    uint16_t op_off;      // first op of the record in the op pool             (role A)
    uint8_t nops;         //                                                   (role A)
    uint8_t stA, stB;     // ST_* per role; walk raises stB
    uint8_t whyA, whyB;   // WHY_* when the role says ST_DEFER
    uint8_t pad[3];
};
static_assert(sizeof(LineRecF) == 36 && offsetof(LineRecF, nops) == 28, "rec_status reads nops / stA / stB as one word");

// step list entry
constexpr uint32_t SE_POS_MASK = 0xFFFFu;     // bits 0..15  buffer position of the separator (sentinel: end of the path column)
constexpr int SE_SLOT_SHIFT = 16;             // bits 16..24 record slot
constexpr uint32_t SE_SLOT_MASK = 0x1FFu;
constexpr uint32_t SE_FIRST = 1u << 25, SE_LAST = 1u << 26, SE_REV = 1u << 27, SE_SENT = 1u << 28, SE_DROPPED = 1u << 29;
constexpr int SE_NCNT_SHIFT = 30;             // bits 30..31 counting ops of the compacted slice (0..3)
constexpr uint32_t SE_INVALID = 0xFFFFFFFFu;

template <int TILE_, int OV_, int THREADS_>
struct Geo {
    static constexpr int TILE = TILE_;
    static constexpr int OV = OV_;
    static constexpr int THREADS = THREADS_;
    static constexpr int BUF = 16 + TILE + OV + 16;               // [pre 16][tile][look-ahead][pad 16]
    static constexpr int NV = ((16 + TILE + OV) / 16 + 3) & ~3;   // 16-byte vectors, padded to whole 64-bit mask words
    static constexpr int LINE_CAP = ((TILE + 111) / 112 + 7) & ~7;   // typical: one record per 300 bytes
    static constexpr int STEP_CAP = (((TILE + OV) / 12 + LINE_CAP) + 63) & ~63;   // typical: 14 steps per 300 bytes, + sentinels
    static constexpr int OPS_CAP = (LINE_CAP * 4 + 63) & ~63;
    static constexpr int FAR_CAP = (STEP_CAP / 8 + 31) & ~31;     // links that are not inline: typically 1-2 per record
    static constexpr int DEL_CAP = (LINE_CAP / 2 + 31) & ~31;     // steps with deletion-derived keys
    static constexpr int MASK_BYTES = 4 * NV;                     // whitespace + separator masks; dead after `records`:
    static constexpr int LIST_BYTES = 2 * STEP_CAP;                    // ... walk 2's list of multi-op steps reuses the space
    static constexpr int OFF_WM = (BUF + 127) & ~127;
    static constexpr int OFF_HEAVY = OFF_WM;
    static constexpr int OFF_SM = OFF_WM + 2 * NV;
    static constexpr int OFF_STEP = (OFF_WM + (MASK_BYTES > LIST_BYTES ? MASK_BYTES : LIST_BYTES) + 15) & ~15;
    static constexpr int OFF_SIDX = OFF_STEP + 4 * STEP_CAP;
    static constexpr int OFF_SINFO = OFF_SIDX + 4 * STEP_CAP;
    static constexpr int OFF_SD01 = OFF_SINFO + 4 * (STEP_CAP + 4);   // inline link deltas of every step's node
    static constexpr int OFF_OPS = OFF_SD01 + 4 * STEP_CAP;
    static constexpr int OFF_LINES = OFF_OPS + 4 * OPS_CAP;
    static constexpr int OFF_REC = (OFF_LINES + 2 * LINE_CAP + 7) & ~7;
    static constexpr int OFF_DEL = (OFF_REC + (int)sizeof(LineRecF) * LINE_CAP + 7) & ~7;   // deletion keys: read while the next tile is scanned
    static constexpr int SMEM_BYTES = (OFF_DEL + 12 * DEL_CAP + 127) & ~127;
    static constexpr int FIT = (227 * 1024) / (SMEM_BYTES + 1024);                    // CTAs per SM by shared memory
    static constexpr int REG = 1024 / THREADS < 1 ? 1 : 1024 / THREADS;               // ... leaving >= 64 registers per thread
    static constexpr int MIN_CTAS = FIT < 1 ? 1 : (FIT < REG ? FIT : REG);
    static_assert(BUF <= 65536, "step entries hold 16-bit positions");
    static_assert(12 * FAR_CAP <= 4 * (STEP_CAP + 4), "the far-link list lives in the step-length prefix array");
    static_assert(LINE_CAP <= 512, "step entries hold 9-bit record slots");
    static_assert(STEP_CAP < 65536 && OPS_CAP < 65536, "records hold 16-bit list offsets");
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
// SWAR byte classes.  Exact when every byte of x is ASCII; a vector with a byte >= 0x80 is a fatal
// PT_U_NON_ASCII error anyway, so what these return for it does not matter.
__device__ __forceinline__ uint32_t flag_ws(uint32_t x) { return ~((x | 0x80808080u) - 0x21212121u) & 0x80808080u; }   // <= 0x20
__device__ __forceinline__ uint32_t flag_eq7(uint32_t y) { return ~(y + 0x7F7F7F7Fu) & 0x80808080u; }                 // y == 0
__device__ __forceinline__ uint32_t flag_tab(uint32_t x) { return flag_eq7(x ^ 0x09090909u); }
__device__ __forceinline__ uint32_t flag_sep(uint32_t x) { return flag_eq7((x | 0x02020202u) ^ 0x3E3E3E3Eu); }        // '>' or '<'

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
__device__ __forceinline__ bool no_colon(const uint8_t* s, uint32_t a, uint32_t b) {
    if (b - a > 48u) return false;
    bool ok = true;
    for (uint32_t q = a; q < b; q++) ok &= s[q] != ':';
    return ok;
}
// a tag the aligner always writes first: "AS:i:<int>" -- inert when nothing after the prefix is a ':'
__device__ __forceinline__ bool tag_is_inert(const uint8_t* s, uint32_t a, uint32_t b) {
    if (b - a >= 6u && s[a] == 'A' && s[a + 1] == 'S' && s[a + 2] == ':' && s[a + 3] == 'i' && s[a + 4] == ':')
        return no_colon(s, a + 5u, b);
    return token_is_inert(s, a, b);
}

// ---- the same tests on whole words (role A: a third of the shared-memory loads of the byte versions)
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
// no ':' in [a, b); first8 = ld8(s, a) if the caller has it already
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

// node id of a path step: digits [a, a + nd) out of shared memory.  false: not a canonical decimal
// (the reference's dict lookup fails: KeyError REF:214) or not a u32.
__device__ __forceinline__ bool step_id(const uint8_t* s, uint32_t a, uint32_t nd, uint64_t& id) {
    if (nd - 1u > 9u) return false;                          // 1..10 digits
    if (nd <= 8u) {
        const uint32_t* w = reinterpret_cast<const uint32_t*>(s + (a & ~3u));
        const uint32_t sh = (a & 3u) * 8u;
        const uint32_t w0 = w[0], w1 = w[1], w2 = w[2];
        const uint32_t lo = __funnelshift_r(w0, w1, sh) ^ 0x30303030u;     // bytes a .. a+3
        const uint32_t hi = __funnelshift_r(w1, w2, sh) ^ 0x30303030u;     // bytes a+4 .. a+7
        if (nd > 1u && (lo & 0xFFu) == 0u) return false;     // leading zero: not the S line's spelling
        // digits to the end of the 8-byte group, zeros (leading digits) in front; later bytes fall off
        const uint64_t y = (((uint64_t)hi << 32) | lo) << (8u * (8u - nd));
        const uint32_t ylo = (uint32_t)y, yhi = (uint32_t)(y >> 32);
        if ((((ylo + 0x76767676u) | ylo) | ((yhi + 0x76767676u) | yhi)) & 0x80808080u) return false;
        id = (uint64_t)(val4(ylo) * 10000u + val4(yhi));
        return true;
    }
    if (s[a] == '0') return false;
    uint64_t v = 0;
    for (uint32_t q = a; q < a + nd; q++) {
        const uint32_t d = (uint32_t)s[q] - '0';
        if (d > 9u) return false;
        v = v * 10u + d;
    }
    id = v;
    return true;
}



// plain digits [a, b), 1..8 of them, no leading zero (anything else: false, the slow path decides)
__device__ __forceinline__ bool small_uint(const uint8_t* s, uint32_t a, uint32_t b, int32_t& out) {
    const uint32_t n = b - a;
    uint64_t v;
    if (n - 1u > 7u || !step_id(s, a, n, v)) return false;
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
__device__ __forceinline__ uint32_t rec_status(const LineRecF& R) {
    const uint32_t w = *reinterpret_cast<const uint32_t*>(&R.nops);        // nops | stA << 8 | stB << 16 | whyA << 24: one LDS
    return max((w >> 8) & 0xFFu, (w >> 16) & 0xFFu);
}

template <class G>
__global__ void __launch_bounds__(G::THREADS, G::MIN_CTAS) augment_fast_kernel(ChunkArgs A, Tables T) {
    constexpr uint32_t THREADS = G::THREADS;
    constexpr uint32_t NWARPS = THREADS / 32;
    PT_DYNAMIC_SMEM(smem);
    __shared__ __align__(8) uint64_t mbar;
    __shared__ uint32_t s_nlines, s_nsteps, s_nops, s_nfar, s_ndel, s_far_next, s_nheavy;
    __shared__ uint32_t s_wsum[NWARPS];

    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    uint8_t* const buf = smem;
    unsigned long long* const wm64 = reinterpret_cast<unsigned long long*>(smem + G::OFF_WM);
    unsigned long long* const sm64 = reinterpret_cast<unsigned long long*>(smem + G::OFF_SM);
    const uint32_t* const wm32 = reinterpret_cast<const uint32_t*>(smem + G::OFF_WM);   // the same masks, as half words
    const uint32_t* const sm32 = reinterpret_cast<const uint32_t*>(smem + G::OFF_SM);
    uint32_t* const far = reinterpret_cast<uint32_t*>(smem + G::OFF_SINFO);    // {from, to, separator position}: filled by `count`, when the step-length
                                                                                 // prefix (same bytes) is dead; drained during the next tile's `records`
    uint32_t* const dels = reinterpret_cast<uint32_t*>(smem + G::OFF_DEL);     // {step, first del, last del}
    uint16_t* const heavy = reinterpret_cast<uint16_t*>(smem + G::OFF_HEAVY);  // steps of multi-op records, for walk 2's dense pass (reuses the masks)
    uint32_t* const steps = reinterpret_cast<uint32_t*>(smem + G::OFF_STEP);
    uint32_t* const sidx = reinterpret_cast<uint32_t*>(smem + G::OFF_SIDX);
    uint32_t* const sinfo = reinterpret_cast<uint32_t*>(smem + G::OFF_SINFO);  // step-length prefix
    uint32_t* const sd01 = reinterpret_cast<uint32_t*>(smem + G::OFF_SD01);    // NodeRec.d01 of the step's node (from `ids`: `count` loads nothing)
    uint32_t* const ops = reinterpret_cast<uint32_t*>(smem + G::OFF_OPS);
    uint16_t* const lines = reinterpret_cast<uint16_t*>(smem + G::OFF_LINES);
    LineRecF* const recs = reinterpret_cast<LineRecF*>(smem + G::OFF_REC);

    if (tid == 0) {
        mbar_init(&mbar, 1);
        s_nlines = 0;
        s_nsteps = 0;
        s_nops = 0;
        s_nfar = 0;
        s_ndel = 0;
        s_far_next = 0;
        s_nheavy = 0;
    }
    __syncthreads();

    DevSink sink(T);
    // diagnostics (PANTAS_PHASE_CLOCKS=1): cycles thread 0 spends between the barriers = wall cycles of the CTA per phase
    const bool phase_clk = (A.stream_hint & 2u) != 0u && tid == 0;
    long long t_prev = phase_clk ? clock64() : 0;
    auto phase_done = [&](int k) {
        if (phase_clk) {
            const long long now = clock64();
            atomicAdd(&T.sc[SC_PHASE + k], (unsigned long long)(now - t_prev));
            t_prev = now;
        }
    };
    const uint64_t nbytes16 = (A.nbytes + 15ull) & ~15ull;
    uint32_t parity = 0;
    unsigned long long my_lines = 0, my_tiles = 0;

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

    // Links that are not inline (hash-table probes, one dependent L2 / DRAM round trip each) are listed by `count` and
    // done one tile later, during `records`, by the warps that phase leaves idle: the probes cost no time of their own.
    // Threads take entries from a shared counter; far_base = file offset of buf[0] of the tile that listed them.
    auto drain_far = [&](int64_t far_base) {
        const uint32_t n_far = min(s_nfar, (uint32_t)G::FAR_CAP);
        for (;;) {
            const uint32_t j = atomicAdd(&s_far_next, 1u);
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
        const uint32_t nvec = (lim + 15u) >> 4, nwords = (nvec + 3u) >> 2;
        mbar_wait(&mbar, parity);
        parity ^= 1;
        phase_done(0);

        // ================= scan: whitespace / separator masks, record starts =================
        for (uint32_t g = tid; g < nwords; g += THREADS) {
            unsigned long long wm = 0, sm = 0;
            uint32_t oth = 0, hib = 0;
            // the four vectors of the group in a lane-dependent order: a quarter warp's eight LDS.128 then fall into
            // eight different 16-byte bank groups (lane stride 64 bytes alone would put them into two)
            const uint32_t rot = (lane >> 1) & 3u;
            uint4 q[4];
#pragma unroll
            for (int u = 0; u < 4; u++) q[u] = *reinterpret_cast<const uint4*>(buf + 64u * g + 16u * ((u + rot) & 3u));
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const uint32_t sh = 16u * ((u + rot) & 3u);
                const uint32_t w0 = flag_ws(q[u].x), w1 = flag_ws(q[u].y), w2 = flag_ws(q[u].z), w3 = flag_ws(q[u].w);
                wm |= (unsigned long long)mask16(w0, w1, w2, w3) << sh;
                sm |= (unsigned long long)mask16(flag_sep(q[u].x), flag_sep(q[u].y), flag_sep(q[u].z), flag_sep(q[u].w)) << sh;
                // whitespace that is not a tab: '\n' (record start), '\r' (lone: error); the rest only matters to the walkers
                oth |= (w0 & ~flag_tab(q[u].x)) | (w1 & ~flag_tab(q[u].y)) | (w2 & ~flag_tab(q[u].z)) | (w3 & ~flag_tab(q[u].w));
                hib |= q[u].x | q[u].y | q[u].z | q[u].w;
            }
            const uint32_t room = lim > 64u * g ? lim - 64u * g : 0u;       // loaded bytes in this group
            unsigned long long keep = room < 64u ? ~(~0ull << room) : ~0ull;
            if (g == 0) keep &= ~0xFFFFull;                                 // positions 0..15 are before the tile
            wm &= keep;
            sm &= keep;
            wm64[g] = wm;
            sm64[g] = sm;
            if (g == 0 && tile == 0 && owned > 0u) {                        // the chunk starts at a record start
                const uint32_t j = atomicAdd(&s_nlines, 1u);
                if (j < (uint32_t)G::LINE_CAP) lines[j] = 16;
            }
            if (oth != 0u) {
                unsigned long long om = 0;
#pragma unroll
                for (int u = 0; u < 4; u++)
                    om |= (unsigned long long)mask16(flag_ws(q[u].x) & ~flag_tab(q[u].x), flag_ws(q[u].y) & ~flag_tab(q[u].y),
                                                     flag_ws(q[u].z) & ~flag_tab(q[u].z), flag_ws(q[u].w) & ~flag_tab(q[u].w))
                          << (16u * ((u + rot) & 3u));
                if (g == 0 && tile != 0) keep |= 0x8000ull;                 // is the byte before the tile a newline?
                om &= keep;
                while (om) {
                    const uint32_t p = 64u * g + (uint32_t)(__ffsll((long long)om) - 1);
                    om &= om - 1ull;
                    const uint32_t c = buf[p];
                    if (c == '\n') {
                        if (p + 1u < own_end) {
                            const uint32_t j = atomicAdd(&s_nlines, 1u);
                            if (j < (uint32_t)G::LINE_CAP) lines[j] = (uint16_t)(p + 1u);
                        }
                    } else if (c == '\r' && p >= 16u && p < own_end) {
                        const uint64_t abs_pos = t0 + p - 16u;
                        if (abs_pos + 1 < A.nbytes && buf[p + 1] != '\n')
                            report_error(T, pt::PT_U_BARE_CR, base_off + (int64_t)p);
                    }
                }
            }
            if ((hib & 0x80808080u) != 0u) {                                // non-ASCII byte: not modelled
                for (uint32_t p = max(64u * g, 16u); p < min(64u * g + 64u, min(lim, own_end)); p++)
                    if (buf[p] >= 0x80u) { report_error(T, pt::PT_U_NON_ASCII, base_off + (int64_t)p); break; }
            }
        }
        __syncthreads();                                                    // ---- masks + record list complete
        phase_done(1);
        const uint32_t n_lines_all = s_nlines;
        if (tid == 0) { my_lines += n_lines_all; my_tiles++; }

        if (n_lines_all > (uint32_t)G::LINE_CAP) {
            // more records than the list holds (pathological input): all of them take the slow path
            for (uint32_t p = 15u + tid; p + 1u < own_end; p += THREADS) {
                const bool nl = p == 15u ? (tile == 0 || buf[p] == '\n') : buf[p] == '\n';
                if (nl) defer_line(T, t0 + p + 1u - 16u, A.file_off, WHY_LINES_FULL);
            }
            drain_far(base_off - (int64_t)gridDim.x * G::TILE);            // the previous tile's far links (normally done in `records`)
            __syncthreads();
            if (tid == 0) {
                s_nlines = 0;
                s_nfar = 0;
                s_far_next = 0;
                const uint32_t nxt = tile + gridDim.x;
                if (nxt < A.n_tiles) issue_load(nxt);
            }
            __syncthreads();
            continue;
        }
        const uint32_t n_lines = n_lines_all;

        // ================= records: two threads per record =================
        // Whole warps per role (a warp that mixes the roles runs them one after the other).  A warp takes as long for
        // one record as for 32, so the records are packed into as few warps as they need: the phase is one pass of
        // each role, the other warps wait at the barrier and cost no issue slots.
        static_assert(NWARPS >= 2u && NWARPS % 2u == 0u, "half of the warps per role");
        constexpr uint32_t RW = NWARPS / 2u;
        const bool roleA = warp >= RW;
#ifndef PT_ROLE_LANES
#define PT_ROLE_LANES 32u                                                   // records per role warp (tuning: fewer = shorter divergent streams, more warps)
#endif
        constexpr uint32_t RL = PT_ROLE_LANES;
        const uint32_t nrw = min((n_lines + RL - 1u) / RL, RW);            // warps per role in use
        for (uint32_t l = RL * (roleA ? warp - RW : warp) + lane; l < n_lines && lane < RL && (roleA ? warp - RW : warp) < nrw; l += RL * nrw) {
            LineRecF& R = recs[l];
            const uint32_t ls = lines[l];
            uint32_t wi = ls >> 5;
            uint32_t wmk = wm32[wi] & (~0u << (ls & 31u));
            uint32_t st = ST_FAST;
            int why = WHY_LONG;
            if (!roleA) {
                // ---------------- role B: columns, filters, coordinates, path steps
                uint32_t e[13];
                e[0] = ls - 1u;
                bool ran_off = false, gaps_ok = true;
                uint32_t tabs = 0xFFFFFFFFu;                  // AND of (byte == '\t') over the first 11 boundaries
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
                int32_t mapq = 0, plen = 0, start = 0, pend = 0;
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
                if (!slow && !done && no_tags) { slow = true; why = WHY_TAGS; }   // no dv tag: ValueError (REF:179), slow path reports
                // ---- path column (REF:185-197): it must start with a separator; count the steps
                uint32_t ns = 0, off = 0, a5 = 0, b5 = 0;
                if (!slow && !done) {
                    why = WHY_PATH;
                    a5 = e[5] + 1u;
                    b5 = e[6];
                    for (uint32_t w = a5 >> 5; w <= ((b5 - 1u) >> 5); w++) ns += (uint32_t)__popc(sep_word(sm32, w, a5, b5));
                    if (ns == 0u || ns > (uint32_t)MAX_STEPS || !((sm32[a5 >> 5] >> (a5 & 31u)) & 1u)) {
                        slow = true;
                    } else {
                        off = atomicAdd(&s_nsteps, ns + 1u);                      // any order: a record only needs a contiguous range
                        if (off + ns + 1u > (uint32_t)G::STEP_CAP) {              // list full: slow path
                            slow = true;
                            why = WHY_STEPS_FULL;
                            for (uint32_t i = off; i < (uint32_t)G::STEP_CAP; i++) steps[i] = SE_INVALID;
                        }
                    }
                }
                st = slow ? ST_DEFER : (done ? ST_DONE : ST_FAST);
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
                    const uint32_t common = (l << SE_SLOT_SHIFT) | (buf[a5] == '<' ? SE_REV : 0u);
                    uint32_t i = off;
                    for (uint32_t w = a5 >> 5; w <= ((b5 - 1u) >> 5); w++) {
                        uint32_t m = sep_word(sm32, w, a5, b5);
                        while (m) {
                            const uint32_t q = 32u * w + (uint32_t)(__ffs((int)m) - 1);
                            m &= m - 1u;
                            steps[i] = q | common | (i == off ? SE_FIRST : 0u) | (i + 1u == off + ns ? SE_LAST : 0u);
                            i++;
                        }
                    }
                    steps[off + ns] = b5 | (l << SE_SLOT_SHIFT) | SE_SENT;
                }
            } else {
                // ---------------- role A: tags -> dv filter, cs ops
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
                bool idle = ran_off || buf[e12] != '\t' || !small_uint(buf, e11 + 1u, e12, mapq) || (int64_t)mapq < A.thr;
                bool slow = false, done = false;
                uint32_t cs_a = 0, cs_b = 0, dv_a = 0, dv_b = 0, dv3 = 0;
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
                    // ---- cs string (REF:10-37): "cs:Z:" then ops spelled the way an aligner spells them:
                    //      ':'<digits>  '*'<2 letters>  '-'<letters>  '+'<letters>  '='<LETTERS>, every length >= 1
                    if (!slow && !done) {
                        why = WHY_CS;
                        uint32_t n_tot = 0, nops = 0, op_off = 0;
                        int32_t start_add = 0;
                        if (cs_b - cs_a < 7u || (cs8 & 0xFFFF000000ull) != 0x3A5A000000ull) slow = true;      // "cs:Z:"
                        uint32_t q = cs_a + 5u;
                        uint64_t one;
                        if (!slow && ((cs8 >> 40) & 0xFFu) == ':' && cs_b - q - 1u <= 7u && step_id(buf, q + 1u, cs_b - q - 1u, one) && one != 0u) {
                            // cs:Z::<n> -- a perfect match
                            op_off = atomicAdd(&s_nops, 1u);
                            if (op_off < (uint32_t)G::OPS_CAP) ops[op_off] = OP_MATCH | ((uint32_t)one << 3);
                            else slow = true;
                            nops = 1;
                            n_tot = (uint32_t)one;
                        } else if (!slow) {
                            // every op takes at least two bytes: room for (bytes / 2) ops is enough
                            const uint32_t room = min((cs_b - q) >> 1, (uint32_t)MAX_OPS);
                            op_off = atomicAdd(&s_nops, room);
                            if (op_off + room > (uint32_t)G::OPS_CAP) slow = true;
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
                                    while (q < cs_b && buf[q] - 'A' <= 24u) { q++; len++; }      // 'A'..'Y': "cs:Z:" cannot hide in here
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
                        R.n_tot = n_tot;
                        R.op_off = (uint16_t)op_off;
                        R.nops = (uint8_t)nops;
                        R.start_add = start_add;
                    }
                }
                st = slow ? ST_DEFER : (done ? ST_DONE : ST_FAST);
                R.stA = (uint8_t)st;
                R.whyA = (uint8_t)why;
            }
        }
        drain_far(base_off - (int64_t)gridDim.x * G::TILE);                // idle warps at once, the others when their records are done
        __syncthreads();                                                    // ---- records, ops, step list complete; the far-link list is empty
        phase_done(2);
        const uint32_t n_ent = min(s_nsteps, (uint32_t)G::STEP_CAP);       // step entries incl. sentinels
        if (tid == 0) s_nlines = 0;                                         // everyone has read it

        // ================= ids: one thread per path step: id -> node index, the node record's read half =================
        // UI steps per thread and iteration: their node-record loads (one 16-byte LDG each) are all in flight before the
        // first is stored.  Everything the later phases need of the record is kept in shared memory -- length (sinfo, until
        // walk 1), inline link deltas (sd01), and whether this tile could still lower a first-touch stamp (two flag bits in
        // sidx; stamps only decrease, so an older value only errs towards one RED.MIN too many): `count` loads nothing.
        {
#ifndef PT_IDS_UNROLL
#define PT_IDS_UNROLL 4
#endif
            constexpr int UI = PT_IDS_UNROLL;
            const int64_t rel0 = base_off + 16 - T.epoch_base;              // below every stamp this tile can produce
            const uint32_t rel_tile = rel0 < 0 ? 0u : (uint32_t)rel0;
            for (uint32_t s00 = 0; s00 < n_ent; s00 += THREADS * UI) {
                uint32_t idx_[UI];
                DevSink::Hot hot_[UI];
#pragma unroll
                for (int u = 0; u < UI; u++) {
                    const uint32_t s = s00 + THREADS * u + tid;
                    uint32_t idx = NONE32;
                    if (s < n_ent) {
                        const uint32_t se = steps[s];
                        if (se != SE_INVALID && !(se & SE_SENT)) {
                            const uint32_t p = se & SE_POS_MASK;
                            const uint32_t end = steps[s + 1u] & SE_POS_MASK;       // next separator, or the sentinel
                            uint64_t id;
                            uint32_t ix;
                            if (buf[p] == ((se & SE_REV) ? '<' : '>') && step_id(buf, p + 1u, end - p - 1u, id) && sink.id_to_idx(id, ix) &&
                                ix < SX_IDX_MASK)
                                idx = ix;
                        }
                    }
                    idx_[u] = idx;                                          // NONE32: KeyError in the reference, `walk` hands the record over
                    // issued at once: the load is in flight while the next id is parsed
                    hot_[u].len = 0; hot_[u].il = 0; hot_[u].ol = 0; hot_[u].d01 = 0;
                    if (idx != NONE32) hot_[u] = sink.load_hot(idx);
                }
#pragma unroll
                for (int u = 0; u < UI; u++) {
                    const uint32_t s = s00 + THREADS * u + tid;
                    if (s < n_ent) {
                        sidx[s] = idx_[u] == NONE32 ? NONE32
                                                    : (idx_[u] | (hot_[u].il > rel_tile ? SX_NEED_IL : 0u) | (hot_[u].ol > rel_tile ? SX_NEED_OL : 0u));
                        sinfo[s] = hot_[u].len;
                        sd01[s] = hot_[u].d01;
                    }
                }
            }
        }
        __syncthreads();                                                    // ---- node indices complete; the bytes and the masks are dead
        phase_done(3);
        if (tid == 0) {
            s_nsteps = 0;
            s_nops = 0;
            s_nfar = 0;                                                     // (drained during `records`)
            s_far_next = 0;
            s_ndel = 0;
            s_nheavy = 0;
            const uint32_t nxt = tile + gridDim.x;
            if (nxt < A.n_tiles) issue_load(nxt);                           // overlaps walk + count
        }

        // ================= walk 1: step lengths and their block-wide prefix sum =================
        const uint32_t per = (n_ent + THREADS - 1u) / THREADS;              // consecutive entries per thread
        const uint32_t sa = min(tid * per, n_ent), sb = min(sa + per, n_ent);
        {
            uint32_t local = 0;
            for (uint32_t s = sa; s < sb; s++) {
                const uint32_t se = steps[s];
                uint32_t Lc = 0;
                if (se != SE_INVALID && !(se & SE_SENT)) {
                    LineRecF& R = recs[(se >> SE_SLOT_SHIFT) & SE_SLOT_MASK];
                    if (rec_status(R) == ST_FAST) {
                        const uint32_t idx = sidx[s];
                        const uint32_t len = sinfo[s];                      // left there by `ids`
                        // unknown id: KeyError (REF:214); collapsible duplicate (REF:188): the slow path redoes the record
                        if (idx == NONE32 || len == pt::NODE_LEN_ABSENT ||
                            (!(se & SE_FIRST) && sidx[s - 1u] != NONE32 && (idx & SX_IDX_MASK) == (sidx[s - 1u] & SX_IDX_MASK))) {
                            R.stB = ST_DEFER;
                            R.whyB = WHY_WALK;
                        } else {
                            int64_t L = (int64_t)len;
                            if (se & SE_FIRST) L -= (int64_t)R.start + R.start_add;           // REF:215-216
                            if (se & SE_LAST) L -= R.end_rel1;                                // REF:217-218
                            Lc = L <= 0 ? 0u : (L > (int64_t)L_CLAMP ? L_CLAMP : (uint32_t)L);
                        }
                    }
                }
                sinfo[s] = local;
                local += Lc;
            }
            uint32_t incl = local;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= (uint32_t)o) incl += y;
            }
            if (lane == 31u) s_wsum[warp] = incl;
            __syncthreads();
            uint32_t basev = incl - local;
            for (uint32_t w = 0; w < warp; w++) basev += s_wsum[w];
            for (uint32_t s = sa; s < sb; s++) {
                const uint32_t g = sinfo[s] + basev;
                sinfo[s] = g;
                const uint32_t se = steps[s];
                if (se != SE_INVALID && (se & SE_FIRST)) recs[(se >> SE_SLOT_SHIFT) & SE_SLOT_MASK].base = g;
            }
        }
        __syncthreads();                                                    // ---- prefix sums, record bases, duplicate / unknown ids known
        phase_done(4);

        // ================= walk 2: every step folds the cs ops that overlap its node =================
        // Pass 1 settles the steps of single-op records (a perfect match, mostly) and lists the others; pass 2 folds the listed
        // steps, dense: a warp pays for its slowest lane, and one read in eight has a cs string with several ops.
        for (uint32_t s = tid; s < n_ent; s += THREADS) {
            const uint32_t se = steps[s];
            if (se == SE_INVALID || (se & SE_SENT)) continue;
            LineRecF& R = recs[(se >> SE_SLOT_SHIFT) & SE_SLOT_MASK];
            if (rec_status(R) != ST_FAST) continue;
            const uint32_t Ak = sinfo[s] - R.base;                          // cs coordinate where this node starts
            const uint32_t Lk = sinfo[s + 1u] - sinfo[s];                   // the sentinel closes the last step
            if (Lk == 0u) {                                                 // no bases left for this node: it is not in `align`
                steps[s] = se | SE_DROPPED;
                continue;
            }
            const uint32_t n_tot = R.n_tot;
            if (Ak >= n_tot) {                                              // cs used up before the path ends: IndexError (REF:227)
                R.stB = ST_DEFER;
                R.whyB = WHY_WALK;
                continue;
            }
            const uint32_t* op = ops + R.op_off;
            const uint32_t nops = R.nops;
            if (nops == 1u) {                                               // one op (a perfect match, mostly): its piece is the whole slice
                const uint32_t kind = op[0] & 7u;
                steps[s] = (kind == OP_DEL || kind == OP_INS) ? (se | SE_DROPPED)                       // REF:101-102
                                                              : (se | ((kind != OP_SUB ? 1u : 0u) << SE_NCNT_SHIFT));
                continue;
            }
            {                                                               // several ops: pass 2
                const uint32_t h = atomicAdd(&s_nheavy, 1u);
                heavy[h] = (uint16_t)s;
            }
        }
        __syncthreads();                                                    // ---- list of multi-op steps complete
        for (uint32_t h = tid; h < s_nheavy; h += THREADS) {
            const uint32_t s = heavy[h];
            const uint32_t se = steps[s];
            LineRecF& R = recs[(se >> SE_SLOT_SHIFT) & SE_SLOT_MASK];
            const uint32_t Ak = sinfo[s] - R.base, Lk = sinfo[s + 1u] - sinfo[s], n_tot = R.n_tot;
            const uint32_t* op = ops + R.op_off;
            const uint32_t nops = R.nops;
            const uint32_t Bk = min(Ak + Lk, n_tot);
            // pieces of the node = ops overlapping [Ak, Bk), clipped; compact_align as a running fold (REF:63-94)
            uint32_t j = 0, o_start = 0, o_end = op[0] >> 3;
            while (o_end <= Ak) {                                           // ends before the node starts (j < nops: Ak < n_tot)
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
            if (nP == 1u && (p0 == OP_DEL || p0 == OP_INS)) {               // clear_align drops the node (REF:101-102)
                steps[s] = se | SE_DROPPED;
                continue;
            }
            if (n_count > 3u) {                                             // more counting ops than the entry holds: slow path
                R.stB = ST_DEFER;
                R.whyB = WHY_WALK;
                continue;
            }
            steps[s] = se | (n_count << SE_NCNT_SHIFT);
            const bool first_del = nQ > 0u && first_op == OP_DEL, last_del = nQ > 0u && qlast_op == OP_DEL;
            if (first_del || last_del) {                                    // deletion-derived IL/OL keys (REF:281-297,317-333)
                const uint32_t k = atomicAdd(&s_ndel, 1u);
                if (k < (uint32_t)G::DEL_CAP) {
                    dels[3u * k] = s;
                    dels[3u * k + 1u] = first_len | (first_del ? 0x80000000u : 0u);
                    dels[3u * k + 2u] = qlast_len | (last_del ? 0x80000000u : 0u);
                } else {
                    R.stB = ST_DEFER;                                       // list full: slow path (nothing counted yet)
                    R.whyB = WHY_WALK;
                }
            }
        }
        __syncthreads();                                                    // ---- every hand-over decision is made; nothing counted so far
        phase_done(5);
        for (uint32_t l = tid; l < n_lines; l += THREADS) {
            const LineRecF& R = recs[l];
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

        // ================= count: one thread per surviving step =================
        // (no global loads here: `ids` left what is needed of the node records in shared memory)
        {
            for (uint32_t s = tid; s < n_ent; s += THREADS) {
                {
                    const uint32_t se = steps[s];
                    if (se == SE_INVALID || (se & (SE_SENT | SE_DROPPED)) || rec_status(recs[(se >> SE_SLOT_SHIFT) & SE_SLOT_MASK]) != ST_FAST)
                        continue;
                    const uint32_t sx = sidx[s], idx = sx & SX_IDX_MASK;
                    const bool rev = (se & SE_REV) != 0u;
                    const uint32_t ps = prev_survivor(s), nx = next_survivor(s);
                    const bool first = ps == NONE32, last = nx == NONE32;   // among the surviving nodes (REF:276-353: i == 0, i == last)
                    const int64_t n_count = (int64_t)(se >> SE_NCNT_SHIFT);
                    const bool il_cond = rev ? !last : !first, ol_cond = rev ? !first : !last;
                    const uint64_t stamp = (uint64_t)(base_off + (int64_t)(se & SE_POS_MASK) + 1) << 2;
                    // this thread owns the link that LEAVES its node: to the next survivor forward, to the previous one reverse
                    const bool have_edge = rev ? !first : !last;
                    int eslot = -1;
                    uint32_t other = 0;
                    if (have_edge) {
                        other = sidx[rev ? ps : nx] & SX_IDX_MASK;
                        eslot = DevSink::inline_slot(sd01[s], idx, other);
                    }
                    sink.bump(idx, eslot);                                              // REF:263-269, 357-363
                    if (have_edge && eslot < 0) {
                        // not inline: hash-table work, collected and done below with every lane busy.  Stamped like
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
                    sink.dense_flagged(idx, il_cond ? n_count : 0, ol_cond ? n_count : 0, stamp | 1u, (sx & SX_NEED_IL) != 0u,
                                       (sx & SX_NEED_OL) != 0u);           // REF:298-351
                }
            }
        }
        phase_done(6);
        // no barrier: the list of deletion keys is complete since walk 2, the far-link list is read one tile later
        {
            // deletion-derived keys: position of the deletion inside the node, IL or OL by orientation
            const uint32_t n_del = min(s_ndel, (uint32_t)G::DEL_CAP);
            for (uint32_t j = tid; j < n_del; j += THREADS) {
                const uint32_t s = dels[3u * j], f = dels[3u * j + 1u], g = dels[3u * j + 2u];
                const uint32_t se = steps[s];
                if (rec_status(recs[(se >> SE_SLOT_SHIFT) & SE_SLOT_MASK]) != ST_FAST) continue;
                const uint32_t idx = sidx[s] & SX_IDX_MASK;
                const int64_t len = (int64_t)sink.load_len(idx);
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
        // The scan writes the masks, the record-start list and its counter; nothing a straggler of this tile still reads
        // (step list, node indices, records, the two lists), and every thread passes the scan's barrier only after it is
        // through here.
        phase_done(7);
    }
    __syncthreads();                                                        // the last tile's far-link list is complete

    drain_far(A.file_off + (int64_t)(tile - gridDim.x) * G::TILE - 16);    // the last tile's far links

    // rejected-record count: warp reduce, one RED per warp
    uint32_t r = sink.rej;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
    if (lane == 0u && r) atomicAdd(&T.sc[SC_REJ], (unsigned long long)r);
    if (tid == 0) {
        if (my_lines) atomicAdd(&T.sc[SC_LINES], my_lines);
        if (my_tiles) atomicAdd(&T.sc[SC_TILES], my_tiles);
    }
}

}  // namespace fastp
