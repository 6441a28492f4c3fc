// fast_tiles.cuh -- the fast path of the augment kernel (included by pantas_aug.cu inside its
// anonymous namespace, after tables.cuh, the TMA helpers, ChunkArgs and defer_line()).
//
// Reference loop body: /root/reference/scripts/alignments_augmentation_from_gaf.py:142-363 (REF:n).
//
// A persistent CTA takes tiles of the GAF chunk (TILE bytes + OV bytes of look-ahead, one 1-D TMA
// bulk copy, UBLKCP) and runs six data-parallel phases over the tile in shared memory:
//
//   scan     one thread per 16 bytes (LDS.128), branch-free SWAR: a 16-bit whitespace mask and a
//            16-bit path-separator ('>' '<') mask per vector; record starts ('\n') go to a list;
//            lone '\r' and non-ASCII bytes are (fatal) errors.
//   records  one thread per record: walks the whitespace mask 64 bytes at a time to get the 12
//            column boundaries and the tag boundaries, then MAPQ / '*' / dv filters
//            (REF:143-148,172-180), the three coordinates (REF:151-153) and the cs string
//            (REF:154-160), classified as
//              SIMPLE  cs:Z::<n>                        (a perfect match)
//              STAR    only ':' and '*' ops, <= 4 '*'   (substitutions only)
//            Everything else -- any other cs op, any whitespace other than single tabs in the
//            first 12 columns, tags that could confuse the reference's regexes, integers that
//            are not plain digits, ... -- is handed to the exact thread-per-record path
//            (line_core.cuh via augment_deferred_kernel).  The thread then walks the separator
//            mask of its path column and appends one entry per path step to the tile's step list.
//   ids      one thread per path step: SWAR decimal parse of the id out of shared memory, node
//            index, L2 prefetch of the node record.  After this phase the bytes are dead and the
//            next tile's TMA copy is issued: it overlaps the table traffic of the last two phases.
//   walk     one thread per record: node lengths -> position of every node in the cs string
//            (REF:205-255 reduces to a prefix sum for the two classes), the one-counting-op test of
//            compact_align (REF:63-94), and every condition under which the slow path must redo
//            the record (duplicate / unknown ids, first or last node without bases REF:215-218,
//            cs shorter than the path REF:227).  Nothing has been counted yet, so the hand-over
//            is clean.
//   count    one thread per path step: NC / IL / OL / RC events (REF:263-363): one RED.ADD.64 on
//            the node's sector (tables.cuh), RED.MIN for first-touch stamps only when earlier.
//
// For the two classes every node with L > 0 survives clear_align (REF:97-107) and its compacted
// slice has exactly one counting op iff the slice holds a ':' piece.
#pragma once

namespace fastp {

constexpr uint32_t NONE32 = 0xffffffffu;
constexpr int MAX_STARS = 4;
constexpr int MAX_STEPS = 250;                // longer paths take the slow path
constexpr uint32_t L_CLAMP = 1u << 23;        // step lengths are clamped here (> any cs length the fast path takes)
constexpr int32_t MAX_NTOT = 1 << 22;

enum : uint8_t { ST_FAST = 0, ST_DEFER = 1, ST_DONE = 2 };

struct __align__(4) LineRecF {
    int32_t start;        // int(tokens[7])
    int32_t end_rel1;     // int(tokens[6]) - int(tokens[8]) - 1
    int32_t n_tot;        // sum of the cs op lengths
    uint16_t s0;          // first entry of the record in the step list
    uint16_t nsteps;
    uint16_t ls;          // buffer position of the record's first byte
    uint16_t b5;          // buffer position of the end of the path column
    uint16_t star[MAX_STARS];   // cs coordinate of every '*' op
    uint8_t nstar;
    uint8_t status;       // ST_*
};

// step list entry: bits 0..15 buffer position of the separator, then flags
constexpr uint32_t SE_POS_MASK = 0xFFFFu;
constexpr uint32_t SE_FIRST = 1u << 16, SE_LAST = 1u << 17, SE_REV = 1u << 18, SE_COUNTS = 1u << 19;
constexpr int SE_SLOT_SHIFT = 20;             // bits 20..31 record slot
constexpr uint32_t SE_INVALID = 0xFFFFFFFFu;

template <int TILE_, int OV_, int THREADS_>
struct Geo {
    static constexpr int TILE = TILE_;
    static constexpr int OV = OV_;
    static constexpr int THREADS = THREADS_;
    static constexpr int BUF = 16 + TILE + OV + 16;               // [pre 16][tile][look-ahead][pad 16]
    static constexpr int NV = ((16 + TILE + OV) / 16 + 3) & ~3;   // 16-byte vectors, padded to whole 64-bit mask words
    static constexpr int STEP_CAP = ((TILE + OV) / 12 + 63) & ~63;   // typical: 14 steps per 300 bytes
    static constexpr int LINE_CAP = ((TILE + 111) / 112 + 7) & ~7;   // typical: one record per 300 bytes
    static constexpr int OFF_WM = (BUF + 127) & ~127;
    static constexpr int OFF_SM = OFF_WM + 2 * NV;
    static constexpr int OFF_STEP = OFF_SM + 2 * NV;
    static constexpr int OFF_SIDX = OFF_STEP + 4 * STEP_CAP;
    static constexpr int OFF_LINES = OFF_SIDX + 4 * STEP_CAP;
    static constexpr int OFF_REC = (OFF_LINES + 2 * LINE_CAP + 7) & ~7;
    static constexpr int FAR_CAP = (STEP_CAP / 6 + 31) & ~31;           // links that are not inline: typically 1-2 per record
    static constexpr int OFF_FAR = (OFF_REC + (int)sizeof(LineRecF) * LINE_CAP + 15) & ~15;
    static constexpr int SMEM_BYTES = (OFF_FAR + 12 * FAR_CAP + 127) & ~127;
    static constexpr int FIT = (227 * 1024) / (SMEM_BYTES + 1024);                    // CTAs per SM by shared memory
    static constexpr int REG = 1024 / THREADS < 1 ? 1 : 1024 / THREADS;               // ... leaving >= 64 registers per thread
    static constexpr int MIN_CTAS = FIT < 1 ? 1 : (FIT < REG ? FIT : REG);
    static_assert(BUF <= 65536, "step entries hold 16-bit positions");
    static_assert(LINE_CAP < 4095, "step entries hold 12-bit record slots");
    static_assert(STEP_CAP < 65536, "records hold 16-bit step list offsets");
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

// a tag the aligner always writes first: "AS:i:<int>" -- inert when nothing after the prefix is a ':'
__device__ __forceinline__ bool no_colon(const uint8_t* s, uint32_t a, uint32_t b) {
    if (b - a > 48u) return false;
    bool ok = true;
    for (uint32_t q = a; q < b; q++) ok &= s[q] != ':';
    return ok;
}
__device__ __forceinline__ bool tag_is_inert(const uint8_t* s, uint32_t a, uint32_t b) {
    if (b - a >= 6u && s[a] == 'A' && s[a + 1] == 'S' && s[a + 2] == ':' && s[a + 3] == 'i' && s[a + 4] == ':')
        return no_colon(s, a + 5u, b);
    return token_is_inert(s, a, b);
}

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

// next whitespace bit at or after the walker's position (64 bytes of the tile per mask word);
// false: ran off the end of the loaded bytes
__device__ __forceinline__ bool next_ws(const unsigned long long* wm64, uint32_t nwords, uint32_t& wi, unsigned long long& m,
                                        uint32_t& pos) {
    while (m == 0ull) {
        if (++wi >= nwords) return false;
        m = wm64[wi];
    }
    pos = 64u * wi + (uint32_t)(__ffsll((long long)m) - 1);
    m &= m - 1ull;
    return true;
}

// separator bits of mask word w that lie in buffer positions [a, b)
__device__ __forceinline__ unsigned long long sep_word(const unsigned long long* sm64, uint32_t w, uint32_t a, uint32_t b) {
    unsigned long long m = sm64[w];
    if (w == (a >> 6)) m &= ~0ull << (a & 63u);
    if (w == (b >> 6)) m &= ~(~0ull << (b & 63u));           // b & 63 == 0: nothing of this word is below b
    return m;
}

template <class G>
__global__ void __launch_bounds__(G::THREADS, G::MIN_CTAS) augment_fast_kernel(ChunkArgs A, Tables T) {
    constexpr uint32_t THREADS = G::THREADS;
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t mbar;
    __shared__ uint32_t s_nlines, s_nsteps, s_nfar;

    const uint32_t tid = threadIdx.x;
    uint8_t* const buf = smem;
    uint16_t* const wm16 = reinterpret_cast<uint16_t*>(smem + G::OFF_WM);
    uint16_t* const sm16 = reinterpret_cast<uint16_t*>(smem + G::OFF_SM);
    const unsigned long long* const wm64 = reinterpret_cast<const unsigned long long*>(smem + G::OFF_WM);
    const unsigned long long* const sm64 = reinterpret_cast<const unsigned long long*>(smem + G::OFF_SM);
    uint32_t* const steps = reinterpret_cast<uint32_t*>(smem + G::OFF_STEP);
    uint32_t* const sidx = reinterpret_cast<uint32_t*>(smem + G::OFF_SIDX);
    uint16_t* const lines = reinterpret_cast<uint16_t*>(smem + G::OFF_LINES);
    LineRecF* const recs = reinterpret_cast<LineRecF*>(smem + G::OFF_REC);
    uint32_t* const far = reinterpret_cast<uint32_t*>(smem + G::OFF_FAR);     // {from, to, separator position} x FAR_CAP

    if (tid == 0) {
        mbar_init(&mbar, 1);
        s_nlines = 0;
        s_nsteps = 0;
        s_nfar = 0;
    }
    __syncthreads();

    DevSink sink(T);
    const uint64_t nbytes16 = (A.nbytes + 15ull) & ~15ull;
    uint32_t parity = 0;
    unsigned long long my_lines = 0, my_tiles = 0;

    auto issue_load = [&](uint32_t tile) {
        const uint64_t t0 = (uint64_t)tile * G::TILE;
        const uint64_t lo = tile ? t0 - 16 : 0;
        const uint64_t hi = min(t0 + G::TILE + G::OV, nbytes16);
        const uint32_t bytes = (uint32_t)(hi - lo);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_expect_tx(&mbar, bytes);
        tma_load_1d(buf + (tile ? 0u : 16u), A.gaf + lo, bytes, &mbar);
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
        const uint32_t nvec = (lim + 15u) >> 4, nvec4 = (nvec + 3u) & ~3u, nwords = nvec4 >> 2;
        mbar_wait(&mbar, parity);
        parity ^= 1;

        // ================= scan: whitespace / separator masks, record starts =================
        // one thread per 64 bytes (four LDS.128): one 64-bit word of each mask per thread and iteration
        unsigned long long* const wm64w = reinterpret_cast<unsigned long long*>(smem + G::OFF_WM);
        unsigned long long* const sm64w = reinterpret_cast<unsigned long long*>(smem + G::OFF_SM);
        for (uint32_t g = tid; g < nwords; g += THREADS) {
            unsigned long long wm = 0, sm = 0;
            uint32_t oth = 0, hib = 0;
            uint4 q[4];
#pragma unroll
            for (int u = 0; u < 4; u++) q[u] = *reinterpret_cast<const uint4*>(buf + 64u * g + 16u * u);
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const uint32_t w0 = flag_ws(q[u].x), w1 = flag_ws(q[u].y), w2 = flag_ws(q[u].z), w3 = flag_ws(q[u].w);
                wm |= (unsigned long long)mask16(w0, w1, w2, w3) << (16 * u);
                sm |= (unsigned long long)mask16(flag_sep(q[u].x), flag_sep(q[u].y), flag_sep(q[u].z), flag_sep(q[u].w)) << (16 * u);
                // whitespace that is not a tab: '\n' (record start), '\r' (lone: error); the rest only matters to the walkers
                oth |= (w0 & ~flag_tab(q[u].x)) | (w1 & ~flag_tab(q[u].y)) | (w2 & ~flag_tab(q[u].z)) | (w3 & ~flag_tab(q[u].w));
                hib |= q[u].x | q[u].y | q[u].z | q[u].w;
            }
            const uint32_t room = lim > 64u * g ? lim - 64u * g : 0u;       // loaded bytes in this group
            unsigned long long keep = room < 64u ? ~(~0ull << room) : ~0ull;
            if (g == 0) keep &= ~0xFFFFull;                                 // positions 0..15 are before the tile
            wm &= keep;
            sm &= keep;
            wm64w[g] = wm;
            sm64w[g] = sm;
            if (g == 0 && tile == 0 && owned > 0u) {                        // the chunk starts at a record start
                const uint32_t j = atomicAdd(&s_nlines, 1u);
                if (j < (uint32_t)G::LINE_CAP) lines[j] = 16;
            }
            if (oth != 0u) {
                unsigned long long om = 0;
#pragma unroll
                for (int u = 0; u < 4; u++)
                    om |= (unsigned long long)mask16(flag_ws(q[u].x) & ~flag_tab(q[u].x), flag_ws(q[u].y) & ~flag_tab(q[u].y),
                                                     flag_ws(q[u].z) & ~flag_tab(q[u].z), flag_ws(q[u].w) & ~flag_tab(q[u].w)) << (16 * u);
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
        const uint32_t n_lines_all = s_nlines;
        if (tid == 0) { my_lines += n_lines_all; my_tiles++; }

        if (n_lines_all > (uint32_t)G::LINE_CAP) {
            // more records than the list holds (pathological input): all of them take the slow path
            for (uint32_t p = 15u + tid; p + 1u < own_end; p += THREADS) {
                const bool nl = p == 15u ? (tile == 0 || buf[p] == '\n') : buf[p] == '\n';
                if (nl) defer_line(T, t0 + p + 1u - 16u, A.file_off, WHY_LINES_FULL);
            }
            __syncthreads();
            if (tid == 0) {
                s_nlines = 0;
                const uint32_t nxt = tile + gridDim.x;
                if (nxt < A.n_tiles) issue_load(nxt);
            }
            __syncthreads();
            continue;
        }
        const uint32_t n_lines = n_lines_all;

        // ================= records: one thread per record =================
        for (uint32_t l = tid; l < n_lines; l += THREADS) {
            bool slow = false, done = false;
            int why = WHY_LONG;
            const uint32_t ls = lines[l];
            uint32_t ns = 0, a5 = 0, b5 = 0;
            int32_t mapq = 0, plen = 0, start = 0, pend = 0, n_tot = 0;
            uint32_t nstar = 0;
            uint16_t star[MAX_STARS] = {0, 0, 0, 0};
            uint32_t wi = ls >> 6;
            unsigned long long wmk = wm64[wi] & (~0ull << (ls & 63u));
            uint32_t e[15];
            e[0] = ls - 1u;
            bool gaps_ok = true;
            uint32_t tabs = 0xFFFFFFFFu;                      // AND of (byte == '\t') over the first 12 boundaries
#pragma unroll
            for (int j = 1; j <= 14; j++) {
                e[j] = 0;
                if (!slow) {
                    if (!next_ws(wm64, nwords, wi, wmk, e[j])) slow = true;          // record runs past the look-ahead
                    else if (j <= 12) {
                        gaps_ok &= e[j] - e[j - 1] >= 2u;                            // no empty column
                        if (j < 12) tabs &= buf[e[j]] == '\t' ? 0xFFFFFFFFu : 0u;
                    }
                }
            }
            // 11 single tabs, then a tab (tags follow) or the end of a 12-column record
            bool no_tags = false;
            if (!slow) {
                const uint32_t c12 = buf[e[12]];
                no_tags = c12 == '\n';
                if (!gaps_ok || tabs == 0u || (c12 != '\t' && !no_tags)) { slow = true; why = WHY_COLUMNS; }
            }
            if (!slow) {
                slow = !small_uint(buf, e[11] + 1u, e[12], mapq);
                why = WHY_INTS;
                if (!slow) {
                    if ((int64_t)mapq < A.thr) { sink.reject(); done = true; }                       // REF:143-146
                    else if (e[6] - e[5] == 2u && buf[e[5] + 1u] == '*') done = true;                 // REF:147-148
                }
            }
            if (!slow && !done)
                slow = !small_uint(buf, e[6] + 1u, e[7], plen) || !small_uint(buf, e[7] + 1u, e[8], start) ||
                       !small_uint(buf, e[8] + 1u, e[9], pend);
            // ---- tags: [inert]* cs [inert]* dv in any order, within the first few tags
            uint32_t cs_a = 0, cs_b = 0, dv_a = 0, dv_b = 0;
            if (!slow && !done && no_tags) { slow = true; why = WHY_TAGS; }       // no dv tag: ValueError (REF:179), slow path reports
            if (!slow && !done) {
                why = WHY_TAGS;
                uint32_t a = e[12] + 1u, b = e[13];
                for (int j = 13;; j++) {
                    if (!cs_b && b - a >= 3u && buf[a] == 'c' && buf[a + 1] == 's' && buf[a + 2] == ':') {
                        cs_a = a;
                        cs_b = b;
                    } else if (!dv_b && b - a >= 6u && buf[a] == 'd' && buf[a + 1] == 'v' && buf[a + 2] == ':' &&
                               buf[a + 3] == 'f' && buf[a + 4] == ':' && pt::is_digit(buf[a + 5]) &&
                               no_colon(buf, a + 5u, b)) {
                        dv_a = a + 5u;
                        dv_b = b;
                    } else if (!tag_is_inert(buf, a, b)) {
                        slow = true;
                        break;
                    }
                    if (cs_b && dv_b) break;
                    if (buf[b] == '\n' || j >= 18) { slow = true; break; }          // end of the record: a tag is missing
                    a = b + 1u;
                    if (j == 13) b = e[14];
                    else if (!next_ws(wm64, nwords, wi, wmk, b)) { slow = true; break; }
                }
            }
            // ---- dv filter (REF:172-180).  The reference parses cs first, but that has no side effects and cannot
            //      raise, so a record that dv filters out needs no cs class
            if (!slow && !done) {
                const uint32_t f = buf[dv_a], g = dv_a + 1u < dv_b ? buf[dv_a + 1u] : 0u, h = dv_a + 2u < dv_b ? buf[dv_a + 2u] : 0u;
                if (f == '0' && g == '.' && h == '0') {
                    // 0.0xxx: never greater
                } else if (pt::dv_token_greater(buf, (int)dv_a, (int)dv_b)) {
                    done = true;
                }
            }
            // ---- cs string: "cs:Z:" then ':'<digits> and '*'<2 letters> ops only (REF:10-37)
            if (!slow && !done) {
                why = WHY_CS;
                if (cs_b - cs_a < 7u || buf[cs_a + 3] != 'Z' || buf[cs_a + 4] != ':') slow = true;
                uint32_t q = cs_a + 5u;
                uint64_t one;
                if (!slow && buf[q] == ':' && cs_b - q - 1u <= 7u && step_id(buf, q + 1u, cs_b - q - 1u, one) && one != 0u) {
                    n_tot = (int32_t)one;                                    // cs:Z::<n> -- a perfect match
                    q = cs_b;
                }
                while (!slow && q < cs_b) {
                    const uint32_t c = buf[q];
                    if (c == ':') {
                        uint32_t v = 0, nd = 0;
                        q++;
                        while (q < cs_b && pt::is_digit(buf[q])) { v = v * 10u + (buf[q] - '0'); q++; nd++; }
                        if (nd == 0u || nd > 7u || v == 0u) slow = true;
                        n_tot += (int32_t)v;
                    } else if (c == '*') {
                        if (q + 3u > cs_b || nstar >= (uint32_t)MAX_STARS || n_tot > 0xFFFF) { slow = true; break; }
                        const uint32_t x = buf[q + 1], y = buf[q + 2];
                        if ((x | 0x20u) - 'a' > 25u || (y | 0x20u) - 'a' > 25u) { slow = true; break; }
#pragma unroll
                        for (int k = 0; k < MAX_STARS; k++)
                            if ((uint32_t)k == nstar) star[k] = (uint16_t)n_tot;
                        nstar++;
                        n_tot += 1;
                        q += 3u;
                    } else {
                        slow = true;
                    }
                }
                if (n_tot <= 0 || n_tot > MAX_NTOT) slow = true;
            }
            // ---- path column (REF:185-197): it must start with a separator; count the steps
            uint32_t off = 0;
            if (!slow && !done) {
                why = WHY_PATH;
                a5 = e[5] + 1u;
                b5 = e[6];
                for (uint32_t w = a5 >> 6; w <= ((b5 - 1u) >> 6); w++) ns += (uint32_t)__popcll(sep_word(sm64, w, a5, b5));
                if (ns == 0u || ns > (uint32_t)MAX_STEPS || !((sm64[a5 >> 6] >> (a5 & 63u)) & 1ull)) {
                    slow = true;
                } else {
                    off = atomicAdd(&s_nsteps, ns);                               // any order: a record only needs a contiguous range
                    if (off + ns > (uint32_t)G::STEP_CAP) {                       // list full: slow path
                        slow = true;
                        why = WHY_STEPS_FULL;
                        for (uint32_t i = off; i < (uint32_t)G::STEP_CAP; i++) steps[i] = SE_INVALID;
                    }
                }
            }
            LineRecF& R = recs[l];
            R.ls = (uint16_t)ls;
            R.status = slow ? ST_DEFER : (done ? ST_DONE : ST_FAST);
            R.nstar = (uint8_t)why;                                        // reason, read by `walk` when status is ST_DEFER
            R.nsteps = 0;
            if (!slow && !done) {
                R.start = start;
                R.end_rel1 = plen - pend - 1;
                R.n_tot = n_tot;
                R.s0 = (uint16_t)off;
                R.nsteps = (uint16_t)ns;
                R.b5 = (uint16_t)b5;
#pragma unroll
                for (int k = 0; k < MAX_STARS; k++) R.star[k] = star[k];
                R.nstar = (uint8_t)nstar;
                // ---- one entry per path step
                const uint32_t common = (l << SE_SLOT_SHIFT) | (buf[a5] == '<' ? SE_REV : 0u);
                uint32_t i = off;
                for (uint32_t w = a5 >> 6; w <= ((b5 - 1u) >> 6); w++) {
                    unsigned long long m = sep_word(sm64, w, a5, b5);
                    while (m) {
                        const uint32_t q = 64u * w + (uint32_t)(__ffsll((long long)m) - 1);
                        m &= m - 1ull;
                        steps[i] = q | common | (i == off ? SE_FIRST : 0u) | (i + 1u == off + ns ? SE_LAST : 0u);
                        i++;
                    }
                }
            }
        }
        __syncthreads();                                                    // ---- records + step list complete
        const uint32_t n_steps = min(s_nsteps, (uint32_t)G::STEP_CAP);
        if (tid == 0) s_nlines = 0;                                         // everyone has read it

        // ================= ids: one thread per path step: id -> node index =================
        for (uint32_t s = tid; s < n_steps; s += THREADS) {
            const uint32_t se = steps[s];
            uint32_t idx = NONE32;
            if (se != SE_INVALID) {
                const uint32_t p = se & SE_POS_MASK;
                const uint32_t end = (se & SE_LAST) ? (uint32_t)recs[se >> SE_SLOT_SHIFT].b5 : (steps[s + 1u] & SE_POS_MASK);
                uint64_t id;
                uint32_t ix;
                if (buf[p] == ((se & SE_REV) ? '<' : '>') && step_id(buf, p + 1u, end - p - 1u, id) && sink.id_to_idx(id, ix)) {
                    idx = ix;
                    sink.prefetch_node(ix);
                }
            }
            sidx[s] = idx;                                                  // NONE32: KeyError in the reference, `walk` hands the record over
        }
        __syncthreads();                                                    // ---- node indices complete; the bytes are dead
        if (tid == 0) {
            s_nsteps = 0;
            s_nfar = 0;
            const uint32_t nxt = tile + gridDim.x;
            if (nxt < A.n_tiles) issue_load(nxt);                           // overlaps walk + count
        }

        // ================= walk: one thread per record: cs coordinates, slow-path conditions =================
        for (uint32_t l = tid; l < n_lines; l += THREADS) {
            LineRecF& R = recs[l];
            uint32_t st = R.status;
            if (st == ST_FAST) {
                const uint32_t s0 = R.s0, ns = R.nsteps, n_tot = (uint32_t)R.n_tot, nstar = R.nstar;
                const int32_t start = R.start, end_rel1 = R.end_rel1;
                uint32_t x[MAX_STARS];
#pragma unroll
                for (int j = 0; j < MAX_STARS; j++) x[j] = R.star[j];
                uint32_t pos = 0, prev = NONE32;
                bool bad = false;
#pragma unroll 4
                for (uint32_t k = 0; k < ns; k++) {
                    const uint32_t idx = sidx[s0 + k];
                    const uint32_t len = sink.load_len(idx == NONE32 ? 0u : idx);
                    // unknown id (KeyError REF:214), collapsible duplicate (REF:188), cs used up (IndexError REF:227)
                    bad |= idx == NONE32 || idx == prev || len == pt::NODE_LEN_ABSENT || pos >= n_tot;
                    int64_t L = (int64_t)len;
                    if (k == 0u) L -= start;                                // REF:215-216
                    if (k + 1u == ns) L -= end_rel1;                        // REF:217-218
                    bad |= L <= 0;                                          // a node without bases drops out of the walk: slow path
                    const uint32_t Lc = L <= 0 ? 0u : (L > (int64_t)L_CLAMP ? L_CLAMP : (uint32_t)L);
                    const uint32_t endp = min(pos + Lc, n_tot);
                    uint32_t stars_in = 0;
                    if (nstar != 0u) {
#pragma unroll
                        for (int j = 0; j < MAX_STARS; j++) stars_in += ((uint32_t)j < nstar && x[j] >= pos && x[j] < endp) ? 1u : 0u;
                    }
                    if (stars_in < endp - pos) steps[s0 + k] |= SE_COUNTS;  // the slice holds a ':' piece (REF:63-94)
                    pos += Lc;
                    prev = idx;
                }
                if (bad) {
                    st = ST_DEFER;
                    for (uint32_t k = 0; k < ns; k++) steps[s0 + k] = SE_INVALID;
                }
            }
            if (st == ST_DEFER) defer_line(T, t0 + R.ls - 16u, A.file_off, R.status == ST_DEFER ? (int)R.nstar : (int)WHY_WALK);
        }
        __syncthreads();                                                    // ---- nothing counted so far; hand-overs done

        // ================= count: one thread per path step =================
        // UB steps per thread and iteration: their node-record loads (one 16-byte LDG each, L2 hits
        // thanks to the prefetch) are all in flight before the first one is used.
        {
            constexpr int UB = 2;
            for (uint32_t s00 = 0; s00 < n_steps; s00 += THREADS * UB) {
                uint32_t se_[UB], idx_[UB];
                DevSink::Hot hot_[UB];
#pragma unroll
                for (int u = 0; u < UB; u++) {
                    const uint32_t s = s00 + THREADS * u + tid;
                    se_[u] = s < n_steps ? steps[s] : SE_INVALID;
                    idx_[u] = 0;
                    hot_[u].len = 0; hot_[u].il = 0; hot_[u].ol = 0; hot_[u].d01 = 0;
                    if (se_[u] != SE_INVALID) {
                        idx_[u] = sidx[s];
                        hot_[u] = sink.load_hot(idx_[u]);
                    }
                }
#pragma unroll
                for (int u = 0; u < UB; u++) {
                    const uint32_t s = s00 + THREADS * u + tid;
                    const uint32_t se = se_[u], idx = idx_[u];
                    if (se == SE_INVALID) continue;
                    const bool first = (se & SE_FIRST) != 0u, last = (se & SE_LAST) != 0u, rev = (se & SE_REV) != 0u;
                    const int64_t n_count = (se & SE_COUNTS) ? 1 : 0;
                    const bool il_cond = rev ? !last : !first, ol_cond = rev ? !first : !last;
                    const uint64_t stamp = (uint64_t)(base_off + (int64_t)(se & SE_POS_MASK) + 1) << 2;
                    // this thread owns the link that LEAVES its node: (k -> k+1) forward, (k -> k-1) reverse
                    const bool have_edge = rev ? !first : !last;
                    int eslot = -1;
                    uint32_t other = 0;
                    if (have_edge) {
                        other = rev ? sidx[s - 1u] : sidx[s + 1u];
                        eslot = DevSink::inline_slot(hot_[u].d01, idx, other);
                    }
                    sink.bump(idx, eslot);                                              // REF:263-269, 357-363
                    if (have_edge && eslot < 0) {
                        // not inline: hash-table work, collected and done below with every lane busy.  Stamped like
                        // the reference's insertion: when the later of the two steps is reached
                        const uint32_t ep = rev ? (se & SE_POS_MASK) : (steps[s + 1u] & SE_POS_MASK);
                        const uint32_t j = atomicAdd(&s_nfar, 1u);
                        if (j < (uint32_t)G::FAR_CAP) {
                            far[3u * j] = idx;
                            far[3u * j + 1u] = other;
                            far[3u * j + 2u] = ep;
                        } else {
                            sink.edge_far(idx, other, (uint64_t)(base_off + (int64_t)ep + 1) << 2);
                        }
                    }
                    DevSink::Stamps st;
                    st.il = hot_[u].il;
                    st.ol = hot_[u].ol;
                    sink.dense(idx, il_cond ? n_count : 0, ol_cond ? n_count : 0, stamp | 1u, st);   // REF:298-351
                }
            }
        }
        __syncthreads();                                                    // ---- far-link list complete
        {
            const uint32_t n_far = min(s_nfar, (uint32_t)G::FAR_CAP);
            for (uint32_t j = tid; j < n_far; j += THREADS)
                sink.edge_far(far[3u * j], far[3u * j + 1u], (uint64_t)(base_off + (int64_t)far[3u * j + 2u] + 1) << 2);
        }
        // no barrier here: the next tile's scan writes only the masks and the record list, which
        // nobody reads any more, and its first barrier orders everything else
    }

    // rejected-record count: warp reduce, one RED per warp
    uint32_t r = sink.rej;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
    if ((tid & 31u) == 0u && r) atomicAdd(&T.sc[SC_REJ], (unsigned long long)r);
    if (tid == 0) {
        if (my_lines) atomicAdd(&T.sc[SC_LINES], my_lines);
        if (my_tiles) atomicAdd(&T.sc[SC_TILES], my_tiles);
    }
}

}  // namespace fastp
