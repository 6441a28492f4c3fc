// fast_tiles.cuh -- the warp-autonomous fast path of the augment kernel (included by pantas_aug.cu
// inside its anonymous namespace, after tables.cuh, the TMA helpers, ChunkArgs and defer_line()).
//
// Reference loop body: /root/reference/scripts/alignments_augmentation_from_gaf.py:142-363 (REF:n).
//
// One WARP owns one mini-tile of the GAF chunk (T bytes + OV bytes of look-ahead) in its own slice
// of shared memory and needs no block-level barrier:
//
//   load     lane 0 issues a 1-D TMA bulk copy (UBLKCP) of [t0 - 16, t0 + T + OV) onto the warp's
//            mbarrier; the copy of the NEXT mini-tile is issued as soon as the bytes of this one
//            are no longer needed (after sweep A), so it overlaps the table traffic of sweep B.
//   scan     byte-parallel and branch-free: every lane takes 16 bytes per iteration (LDS.128) and
//            turns them, with SWAR compares and a multiply-gather, into two 16-bit masks -- ASCII
//            whitespace and path separators ('>' '<') -- stored as one word per 16-byte vector.
//            Record starts ('\n') are appended to a list; lone '\r' and non-ASCII bytes are errors.
//   lines    one lane per record (records are alike, so this is balanced): walks the whitespace
//            mask from the record start to get the 12 column boundaries and the tag boundaries
//            without touching the bytes, then MAPQ / '*' / dv filters (REF:143-148,172-180), the
//            three coordinates (REF:151-153) and the cs string (REF:154-160), classified as
//              SIMPLE  cs:Z::<n>                        (a perfect match)
//              STAR    only ':' and '*' ops, <= 4 '*'   (substitutions only)
//            Everything else -- any other cs op, any whitespace other than single tabs in the
//            first 12 columns, tags that could confuse the reference's regexes, integers that
//            are not plain digits, ... -- is handed to the exact thread-per-record path
//            (line_core.cuh via augment_deferred_kernel) BEFORE any counter is touched.
//            The lane then walks the separator mask of its path column and appends one entry per
//            path step {position, digits, record, first/last} to the warp's step list.
//   sweep A  one lane per path step: SWAR decimal parse of the id out of shared memory, node
//            index, L2 prefetch of the node record; collapsible duplicates (REF:188) and
//            malformed steps mark the record for the slow path.
//   check    one lane per record: first / last node must keep a positive length after the
//            start / end offsets (REF:215-218), otherwise slow path.
//   sweep B  one lane per path step: node length -> warp prefix sum -> position of the node in the
//            cs string -> NC / IL / OL / RC events (REF:263-363): one RED.ADD.64 on the node's
//            sector (tables.cuh), RED.MIN for first-touch stamps only when earlier.
//
// For the two classes above every node with L > 0 survives clear_align (REF:97-107) and its
// compacted slice has exactly one counting op iff the slice holds a ':' piece (REF:63-94), so
// nothing sequential is left of the reference's walk except a prefix sum.
#pragma once

namespace fastp {

constexpr uint32_t FULL = 0xffffffffu;
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
    uint32_t base;        // running step-length prefix at the record's first step (sweep B)
    uint16_t s0;          // first entry of the record in the step list
    uint16_t nsteps;
    uint16_t ls;          // buffer position of the record's first byte
    uint16_t star[MAX_STARS];   // cs coordinate of every '*' op
    uint8_t nstar;
    uint8_t status;       // ST_*
};

// step list entry
constexpr uint32_t SE_POS_MASK = 0x3FFFu;     // bits 0..13  buffer position of the separator
constexpr int SE_ND_SHIFT = 14;               // bits 14..17 digits of the id (0 / 15 = malformed)
constexpr int SE_SLOT_SHIFT = 18;             // bits 18..24 record slot
constexpr uint32_t SE_FIRST = 1u << 25, SE_LAST = 1u << 26, SE_REV = 1u << 27;

template <int T_, int OV_, int WARPS_>
struct Geo {
    static constexpr int T = T_;
    static constexpr int OV = OV_;
    static constexpr int WARPS = WARPS_;                          // warps per CTA
    static constexpr int BUF = 16 + T + OV + 16;                  // [pre 16][tile][look-ahead][pad 16]
    static constexpr int NVEC_CAP = (16 + T + OV) / 16 + 1;
    static constexpr int STEP_CAP = ((T + OV) / 12 + 63) & ~63;   // typical: 14 steps per 300 bytes
    static constexpr int LINE_CAP = (T / 112 + 7) & ~7;           // <= 127 (slot field)
    static constexpr int OFF_MASK = (BUF + 127) & ~127;
    static constexpr int OFF_STEP = OFF_MASK + 4 * NVEC_CAP;
    static constexpr int OFF_SIDX = OFF_STEP + 4 * STEP_CAP;
    static constexpr int OFF_LINES = OFF_SIDX + 4 * STEP_CAP;
    static constexpr int OFF_REC = (OFF_LINES + 2 * LINE_CAP + 7) & ~7;
    static constexpr int WARP_BYTES = (OFF_REC + (int)sizeof(LineRecF) * LINE_CAP + 127) & ~127;
    static_assert(BUF <= 16384, "step entries hold 14-bit positions");
    static_assert(LINE_CAP <= 127, "step entries hold 7-bit record slots");
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

// next whitespace bit at or after the walker's position; false: ran off the end of the loaded bytes
__device__ __forceinline__ bool next_ws(const uint32_t* masks, uint32_t nvec, uint32_t& v, uint32_t& m, uint32_t& pos) {
    while (m == 0u) {
        if (++v >= nvec) return false;
        m = masks[v] & 0xFFFFu;
    }
    pos = 16u * v + (uint32_t)(__ffs((int)m) - 1);
    m &= m - 1u;
    return true;
}

template <class G>
__global__ void __launch_bounds__(G::WARPS * 32) augment_fast_kernel(ChunkArgs A, Tables T) {
    constexpr int WARPS = G::WARPS;
    extern __shared__ __align__(128) uint8_t smem_all[];
    __shared__ __align__(8) uint64_t mbar_all[WARPS];
    __shared__ uint32_t s_nl[WARPS];

    const uint32_t lane = threadIdx.x & 31u, wib = threadIdx.x >> 5;
    uint8_t* const wsm = smem_all + (size_t)wib * G::WARP_BYTES;
    uint8_t* const buf = wsm;
    uint32_t* const masks = reinterpret_cast<uint32_t*>(wsm + G::OFF_MASK);
    uint32_t* const steps = reinterpret_cast<uint32_t*>(wsm + G::OFF_STEP);
    uint32_t* const sidx = reinterpret_cast<uint32_t*>(wsm + G::OFF_SIDX);
    uint16_t* const lines = reinterpret_cast<uint16_t*>(wsm + G::OFF_LINES);
    LineRecF* const recs = reinterpret_cast<LineRecF*>(wsm + G::OFF_REC);
    uint64_t* const mbar = &mbar_all[wib];
    uint32_t* const nl_cnt = &s_nl[wib];

    if (lane == 0) mbar_init(mbar, 1);
    __syncwarp();

    DevSink sink(T);
    const uint64_t nbytes16 = (A.nbytes + 15ull) & ~15ull;
    const uint32_t n_warps = gridDim.x * WARPS;
    uint32_t parity = 0;
    unsigned long long my_lines = 0, my_tiles = 0;

    auto issue_load = [&](uint32_t tile) {
        const uint64_t t0 = (uint64_t)tile * G::T;
        const uint64_t lo = tile ? t0 - 16 : 0;
        const uint64_t hi = min(t0 + G::T + G::OV, nbytes16);
        const uint32_t bytes = (uint32_t)(hi - lo);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_expect_tx(mbar, bytes);
        tma_load_1d(buf + (tile ? 0u : 16u), A.gaf + lo, bytes, mbar);
    };

    uint32_t tile = blockIdx.x * WARPS + wib;
    if (tile < A.n_tiles && lane == 0) issue_load(tile);

    for (; tile < A.n_tiles; tile += n_warps) {
        const uint64_t t0 = (uint64_t)tile * G::T;
        const uint32_t owned = (uint32_t)min((uint64_t)G::T, A.nbytes - t0);
        const uint64_t hi = min(t0 + G::T + G::OV, nbytes16);
        const uint32_t lim = 16u + (uint32_t)(min(hi, A.nbytes) - t0);     // data ends here in the buffer
        const int64_t base_off = A.file_off + (int64_t)t0 - 16;            // file offset of buf[0]
        const uint32_t own_end = 16u + owned;                               // records starting before this are ours
        if (lane == 0) {
            *nl_cnt = 0;
            if (tile == 0) { lines[0] = 16; *nl_cnt = 1; }                  // the chunk starts at a record start
        }
        mbar_wait(mbar, parity);
        parity ^= 1;
        __syncwarp();

        // ================= scan: whitespace / separator masks, record starts =================
        const uint32_t nvec = (lim + 15u) >> 4;
        for (uint32_t v0 = 0; v0 < nvec; v0 += 32) {
            const uint32_t v = v0 + lane;
            if (v >= nvec) break;
            const uint4 q4 = *reinterpret_cast<const uint4*>(buf + 16u * v);
            const uint32_t w0 = flag_ws(q4.x), w1 = flag_ws(q4.y), w2 = flag_ws(q4.z), w3 = flag_ws(q4.w);
            uint32_t wm = mask16(w0, w1, w2, w3);
            uint32_t sm = mask16(flag_sep(q4.x), flag_sep(q4.y), flag_sep(q4.z), flag_sep(q4.w));
            // whitespace that is not a tab: '\n' (record start), '\r' (lone: error), the rest only matters to the walkers
            const uint32_t o0 = w0 & ~flag_tab(q4.x), o1 = w1 & ~flag_tab(q4.y), o2 = w2 & ~flag_tab(q4.z), o3 = w3 & ~flag_tab(q4.w);
            const uint32_t room = lim - 16u * v;                            // > 0
            const uint32_t keep = room < 16u ? (1u << room) - 1u : 0xFFFFu;
            wm &= keep;
            sm &= keep;
            if (v == 0) { wm = 0; sm = 0; }                                 // bytes before the tile
            masks[v] = wm | (sm << 16);
            if ((o0 | o1 | o2 | o3) != 0u) {
                uint32_t om = mask16(o0, o1, o2, o3) & keep;
                if (v == 0) om &= 0x8000u;                                  // only: is the byte before the tile a newline?
                while (om) {
                    const uint32_t p = 16u * v + (uint32_t)(__ffs((int)om) - 1);
                    om &= om - 1u;
                    const uint32_t c = buf[p];
                    if (c == '\n') {
                        if (p + 1u < own_end && !(tile == 0 && p == 15u)) {
                            const uint32_t j = atomicAdd(nl_cnt, 1u);
                            if (j < (uint32_t)G::LINE_CAP) lines[j] = (uint16_t)(p + 1u);
                        }
                    } else if (c == '\r' && p >= 16u && p < own_end) {
                        const uint64_t abs_pos = t0 + p - 16u;
                        if (abs_pos + 1 < A.nbytes && buf[p + 1] != '\n')
                            report_error(T, pt::PT_U_BARE_CR, base_off + (int64_t)p);
                    }
                }
            }
            const uint32_t hib = (q4.x | q4.y | q4.z | q4.w) & 0x80808080u;
            if (hib != 0u) {                                                // non-ASCII byte: not modelled
                uint32_t hm = mask16(q4.x & 0x80808080u, q4.y & 0x80808080u, q4.z & 0x80808080u, q4.w & 0x80808080u) & keep;
                if (v == 0) hm = 0;
                if (hm) {
                    const uint32_t p = 16u * v + (uint32_t)(__ffs((int)hm) - 1);
                    if (p < own_end) report_error(T, pt::PT_U_NON_ASCII, base_off + (int64_t)p);
                }
            }
        }
        __syncwarp();
        const uint32_t n_lines_all = *(volatile uint32_t*)nl_cnt;
        if (lane == 0) { my_lines += n_lines_all; my_tiles++; }

        if (n_lines_all > (uint32_t)G::LINE_CAP) {
            // more records than the list holds (pathological input): all of them take the slow path
            for (uint32_t p = 15u + lane; p + 1u < own_end; p += 32u) {
                const bool nl = (p == 15u && tile == 0) || (p >= (tile ? 15u : 16u) && buf[p] == '\n');
                if (nl) defer_line(T, t0 + p + 1u - 16u, A.file_off);
            }
            __syncwarp();
            const uint32_t nxt = tile + n_warps;
            if (nxt < A.n_tiles && lane == 0) issue_load(nxt);
            continue;
        }
        const uint32_t n_lines = n_lines_all;

        // ================= lines: one lane per record =================
        uint32_t n_steps = 0;                                 // entries in the step list (warp-uniform)
        for (uint32_t l0 = 0; l0 < n_lines; l0 += 32) {
            const uint32_t l = l0 + lane;
            const bool valid = l < n_lines;
            bool slow = false, done = false;
            uint32_t ls = 16, ns = 0, a5 = 0, b5 = 0;
            int32_t mapq = 0, plen = 0, start = 0, pend = 0, n_tot = 0;
            uint32_t nstar = 0;
            uint16_t star[MAX_STARS] = {0, 0, 0, 0};
            if (valid) {
                ls = lines[l];
                uint32_t wv = ls >> 4, wmk = (masks[wv] & 0xFFFFu) & ~((1u << (ls & 15u)) - 1u);
                uint32_t e[15];
                e[0] = ls - 1u;
                bool gaps_ok = true;
                uint32_t tabs = 0xFFFFFFFFu;                  // AND of (byte == '\t') over the first 12 boundaries
#pragma unroll
                for (int j = 1; j <= 14; j++) {
                    e[j] = 0;
                    if (!slow) {
                        if (!next_ws(masks, nvec, wv, wmk, e[j])) slow = true;      // record runs past the look-ahead
                        else if (j <= 12) {
                            gaps_ok &= e[j] - e[j - 1] >= 2u;                        // no empty column
                            tabs &= buf[e[j]] == '\t' ? 0xFFFFFFFFu : 0u;
                        }
                    }
                }
                slow = slow || !gaps_ok || tabs == 0u;
                if (!slow) {
                    slow = !small_uint(buf, e[11] + 1u, e[12], mapq);
                    if (!slow) {
                        if ((int64_t)mapq < A.thr) { sink.reject(); done = true; }                   // REF:143-146
                        else if (e[6] - e[5] == 2u && buf[e[5] + 1u] == '*') done = true;             // REF:147-148
                    }
                }
                if (!slow && !done)
                    slow = !small_uint(buf, e[6] + 1u, e[7], plen) || !small_uint(buf, e[7] + 1u, e[8], start) ||
                           !small_uint(buf, e[8] + 1u, e[9], pend);
                // ---- tags: [inert]* cs [inert]* dv in any order, within the first few tags
                uint32_t cs_a = 0, cs_b = 0, dv_a = 0, dv_b = 0;
                if (!slow && !done) {
                    uint32_t a = e[12] + 1u, b = e[13];
                    for (int j = 13;; j++) {
                        if (!cs_b && b - a >= 3u && buf[a] == 'c' && buf[a + 1] == 's' && buf[a + 2] == ':') {
                            cs_a = a;
                            cs_b = b;
                        } else if (!dv_b && b - a >= 6u && buf[a] == 'd' && buf[a + 1] == 'v' && buf[a + 2] == ':' &&
                                   buf[a + 3] == 'f' && buf[a + 4] == ':' && pt::is_digit(buf[a + 5]) &&
                                   token_is_inert(buf, a + 5u, b)) {
                            dv_a = a + 5u;
                            dv_b = b;
                        } else if (!token_is_inert(buf, a, b)) {
                            slow = true;
                            break;
                        }
                        if (cs_b && dv_b) break;
                        if (buf[b] == '\n' || j >= 18) { slow = true; break; }     // end of the record: a tag is missing
                        a = b + 1u;
                        if (j == 13) b = e[14];
                        else if (!next_ws(masks, nvec, wv, wmk, b)) { slow = true; break; }
                    }
                }
                // ---- cs string: "cs:Z:" then ':'<digits> and '*'<2 letters> ops only (REF:10-37)
                if (!slow && !done) {
                    if (cs_b - cs_a < 7u || buf[cs_a + 3] != 'Z' || buf[cs_a + 4] != ':') slow = true;
                    uint32_t q = cs_a + 5u;
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
                // ---- dv filter (REF:172-180), after cs like the reference (cs has no side effects)
                if (!slow && !done) {
                    const uint32_t f = buf[dv_a], g = dv_a + 1u < dv_b ? buf[dv_a + 1u] : 0u, h = dv_a + 2u < dv_b ? buf[dv_a + 2u] : 0u;
                    if (f == '0' && g == '.' && h == '0') {
                        // 0.0xxx: never greater
                    } else if (pt::dv_token_greater(buf, (int)dv_a, (int)dv_b)) {
                        done = true;
                    }
                }
                // ---- path column (REF:185-197): count the separators; it must start with one
                if (!slow && !done) {
                    a5 = e[5] + 1u;
                    b5 = e[6];
                    const uint32_t va = a5 >> 4, vb = (b5 - 1u) >> 4;
                    for (uint32_t v = va; v <= vb; v++) {
                        uint32_t m = masks[v] >> 16;
                        if (v == va) m &= ~((1u << (a5 & 15u)) - 1u);
                        if (v == vb && (b5 & 15u)) m &= (1u << (b5 & 15u)) - 1u;
                        ns += __popc(m);
                    }
                    if (ns == 0u || ns > (uint32_t)MAX_STEPS || !((masks[va] >> 16) & (1u << (a5 & 15u)))) { slow = true; ns = 0; }
                }
                if (slow || done) ns = 0;
            }
            // ---- room in the step list: warp prefix sum of the step counts
            uint32_t incl = ns;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t y = __shfl_up_sync(FULL, incl, o);
                if (lane >= (uint32_t)o) incl += y;
            }
            const uint32_t off = n_steps + incl - ns;
            n_steps = min(n_steps + __shfl_sync(FULL, incl, 31), (uint32_t)G::STEP_CAP);
            if (valid) {
                if (!slow && !done && off + ns > (uint32_t)G::STEP_CAP) {                  // list full: slow path
                    slow = true;
                    for (uint32_t i = off; i < (uint32_t)G::STEP_CAP; i++) steps[i] = 0xFFFFFFFFu;   // names no record
                }
                LineRecF& R = recs[l];
                R.ls = (uint16_t)ls;
                R.status = slow ? ST_DEFER : (done ? ST_DONE : ST_FAST);
                R.nsteps = 0;
                if (!slow && !done) {
                    R.start = start;
                    R.end_rel1 = plen - pend - 1;
                    R.n_tot = n_tot;
                    R.base = 0;
                    R.s0 = (uint16_t)off;
                    R.nsteps = (uint16_t)ns;
#pragma unroll
                    for (int k = 0; k < MAX_STARS; k++) R.star[k] = star[k];
                    R.nstar = (uint8_t)nstar;
                    // ---- one entry per path step
                    const uint32_t common = (l << SE_SLOT_SHIFT) | (buf[a5] == '<' ? SE_REV : 0u);
                    const uint32_t va = a5 >> 4, vb = (b5 - 1u) >> 4;
                    uint32_t prev = a5, i = off;
                    bool malformed = false;
                    for (uint32_t v = va; v <= vb; v++) {
                        uint32_t m = masks[v] >> 16;
                        if (v == va) m &= ~((2u << (a5 & 15u)) - 1u);        // the first separator is `prev`
                        if (v == vb && (b5 & 15u)) m &= (1u << (b5 & 15u)) - 1u;
                        while (m) {
                            const uint32_t q = 16u * v + (uint32_t)(__ffs((int)m) - 1);
                            m &= m - 1u;
                            const uint32_t nd = q - prev - 1u;
                            malformed |= nd - 1u > 9u;
                            steps[i] = prev | (min(nd, 15u) << SE_ND_SHIFT) | common | (i == off ? SE_FIRST : 0u);
                            i++;
                            prev = q;
                        }
                    }
                    const uint32_t nd = b5 - prev - 1u;
                    malformed |= nd - 1u > 9u;
                    steps[i] = prev | (min(nd, 15u) << SE_ND_SHIFT) | common | (i == off ? SE_FIRST : 0u) | SE_LAST;
                    if (malformed) R.status = ST_DEFER;       // empty or > 10 digit id: KeyError, the slow path reports it
                }
            }
        }
        __syncwarp();

        // ================= sweep A: one lane per path step: id -> node index =================
        {
            uint32_t prev_last = NONE32;      // node index of lane 31 of the previous pass
            for (uint32_t s00 = 0; s00 < n_steps; s00 += 32) {
                const uint32_t s = s00 + lane;
                uint32_t idx = NONE32, se = 0;
                bool live = false;
                if (s < n_steps) {
                    se = steps[s];
                    const uint32_t slot = (se >> SE_SLOT_SHIFT) & 0x7Fu;
                    LineRecF& R = recs[slot < n_lines ? slot : 0u];
                    if (slot < n_lines && R.status == ST_FAST) {
                        live = true;
                        const uint32_t p = se & SE_POS_MASK;
                        uint64_t id;
                        uint32_t ix;
                        if (buf[p] == ((se & SE_REV) ? '<' : '>') && step_id(buf, p + 1u, (se >> SE_ND_SHIFT) & 15u, id) &&
                            sink.id_to_idx(id, ix)) {
                            idx = ix;
                            sink.prefetch_node(ix);
                        } else {
                            R.status = ST_DEFER;              // KeyError in the reference: the slow path reports it
                        }
                    }
                    sidx[s] = idx;
                }
                // consecutive duplicate ids collapse (REF:188): rare, slow path
                uint32_t before = __shfl_up_sync(FULL, idx, 1);
                if (lane == 0) before = prev_last;
                prev_last = __shfl_sync(FULL, idx, 31);
                if (live && !(se & SE_FIRST) && idx != NONE32 && before == idx) recs[(se >> SE_SLOT_SHIFT) & 0x7Fu].status = ST_DEFER;
            }
        }
        __syncwarp();
        // the bytes of this mini-tile are not needed any more: fetch the next one under sweep B
        {
            const uint32_t nxt = tile + n_warps;
            if (nxt < A.n_tiles && lane == 0) issue_load(nxt);
        }

        // ================= check: first / last node keep a positive length (REF:215-218); hand-over =================
        for (uint32_t l0 = 0; l0 < n_lines; l0 += 32) {
            const uint32_t l = l0 + lane;
            if (l >= n_lines) continue;
            LineRecF& R = recs[l];
            if (R.status == ST_FAST) {
                const uint32_t ns = R.nsteps;
                const uint32_t i0 = sidx[R.s0], i1 = sidx[R.s0 + ns - 1u];
                const uint32_t len0 = sink.load_len(i0), len1 = sink.load_len(i1);
                bool bad = len0 == pt::NODE_LEN_ABSENT || len1 == pt::NODE_LEN_ABSENT;   // KeyError: slow path reports
                if (!bad) {
                    int64_t L0 = (int64_t)len0 - R.start;
                    int64_t L1 = (int64_t)len1 - R.end_rel1;
                    if (ns == 1u) { L0 -= R.end_rel1; L1 = L0; }
                    bad = L0 <= 0 || L1 <= 0;
                }
                if (bad) R.status = ST_DEFER;
            }
            if (R.status == ST_DEFER) defer_line(T, t0 + R.ls - 16u, A.file_off);
        }
        __syncwarp();

        // ================= sweep B: one lane per path step: count =================
        // UB passes of 32 steps per iteration: their node-record loads (one 16-byte LDG each, L2 hits
        // thanks to sweep A's prefetch) are all in flight before the first one is used.
        {
            constexpr int UB = 4;
            uint32_t run = 0;                 // running sum of step lengths over the mini-tile (mod 2^32)
            for (uint32_t s00 = 0; s00 < n_steps; s00 += 32 * UB) {
                uint32_t se_[UB], idx_[UB];
                DevSink::Hot hot_[UB];
                bool act_[UB];
#pragma unroll
                for (int u = 0; u < UB; u++) {
                    const uint32_t s = s00 + 32u * u + lane;
                    se_[u] = 0; idx_[u] = NONE32; act_[u] = false;
                    hot_[u].len = 0; hot_[u].il = 0; hot_[u].ol = 0; hot_[u].d01 = 0;
                    if (s < n_steps) {
                        se_[u] = steps[s];
                        const uint32_t slot = (se_[u] >> SE_SLOT_SHIFT) & 0x7Fu;
                        if (slot < n_lines && recs[slot].status == ST_FAST) {
                            act_[u] = true;
                            idx_[u] = sidx[s];
                            hot_[u] = sink.load_hot(idx_[u]);
                        }
                    }
                }
#pragma unroll
                for (int u = 0; u < UB; u++) {
                    if (s00 + 32u * u >= n_steps) break;                  // warp-uniform
                    const uint32_t s = s00 + 32u * u + lane;
                    const bool active = act_[u];
                    const uint32_t se = se_[u], idx = idx_[u];
                    const bool first = (se & SE_FIRST) != 0u, last = (se & SE_LAST) != 0u, rev = (se & SE_REV) != 0u;
                    LineRecF& R = recs[active ? (se >> SE_SLOT_SHIFT) & 0x7Fu : 0u];
                    bool fatal = false;
                    uint32_t Lk = 0;
                    if (active) {
                        uint32_t len = hot_[u].len;
                        if (len == pt::NODE_LEN_ABSENT) { fatal = true; len = 1; }   // KeyError REF:214
                        int64_t L = (int64_t)len;
                        if (first) L -= R.start;
                        if (last) L -= R.end_rel1;
                        // > 0: interior nodes have len >= 1, the ends were checked; clamped so that
                        // 250 steps cannot wrap the 32-bit prefix (n_tot <= MAX_NTOT < L_CLAMP)
                        Lk = L > (int64_t)L_CLAMP ? L_CLAMP : (uint32_t)L;
                    }
                    uint32_t incl = Lk;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const uint32_t y = __shfl_up_sync(FULL, incl, o);
                        if (lane >= (uint32_t)o) incl += y;
                    }
                    const uint32_t excl = incl - Lk + run;
                    run += __shfl_sync(FULL, incl, 31);
                    if (active && first) R.base = excl;
                    __syncwarp();
                    if (active) {
                        const uint32_t Ak = excl - R.base;        // cs coordinate where this node starts
                        const uint32_t n_tot = (uint32_t)R.n_tot;
                        if (Ak >= n_tot) fatal = true;            // cs used up before the path ends (IndexError REF:227)
                        if (fatal) {
                            defer_line(T, t0 + R.ls - 16u, A.file_off);     // the slow path reports the exact error
                        } else {
                            const uint32_t Bk = min(Ak + Lk, n_tot);
                            uint32_t stars_in = 0;
                            const uint32_t nstar = R.nstar;
#pragma unroll
                            for (int j = 0; j < MAX_STARS; j++) {
                                const uint32_t x = R.star[j];
                                stars_in += ((uint32_t)j < nstar && x >= Ak && x < Bk) ? 1u : 0u;
                            }
                            const int64_t n_count = stars_in < Bk - Ak ? 1 : 0;   // slice holds a ':' piece (REF:63-94)
                            const bool il_cond = rev ? !last : !first, ol_cond = rev ? !first : !last;
                            const uint64_t stamp = (uint64_t)(base_off + (int64_t)(se & SE_POS_MASK) + 1) << 2;
                            // this lane owns the link that LEAVES its node: (k -> k+1) forward, (k -> k-1) reverse
                            const bool have_edge = rev ? !first : !last;
                            int eslot = -1;
                            uint32_t other = 0;
                            if (have_edge) {
                                other = rev ? sidx[s - 1u] : sidx[s + 1u];
                                eslot = DevSink::inline_slot(hot_[u].d01, idx, other);
                            }
                            sink.bump(idx, eslot);                                          // REF:263-269, 357-363
                            if (have_edge && eslot < 0) {
                                // stamped like the reference's insertion: when the later of the two steps is reached
                                const uint64_t es = rev ? stamp : (uint64_t)(base_off + (int64_t)(steps[s + 1u] & SE_POS_MASK) + 1) << 2;
                                sink.edge_far(idx, other, es);
                            }
                            DevSink::Stamps st;
                            st.il = hot_[u].il;
                            st.ol = hot_[u].ol;
                            sink.dense(idx, il_cond ? n_count : 0, ol_cond ? n_count : 0, stamp | 1u, st);   // REF:298-351
                        }
                    }
                    __syncwarp();
                }
            }
        }
        __syncwarp();
    }

    uint32_t r = sink.rej;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(FULL, r, o);
    if (lane == 0) {
        if (r) atomicAdd(&T.sc[SC_REJ], (unsigned long long)r);
        if (my_lines) atomicAdd(&T.sc[SC_LINES], my_lines);
        if (my_tiles) atomicAdd(&T.sc[SC_TILES], my_tiles);
    }
}

}  // namespace fastp
