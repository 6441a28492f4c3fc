// line_core.cuh -- one GAF record -> counter events, as an O(1)-state stream.
//
// This is the B200 formulation of the reference's per-line body
//   /root/reference/scripts/alignments_augmentation_from_gaf.py:142-363   (REF:n)
// The reference materialises token lists, an op list, per-node op slices
// (`align`), the cleared/compacted copy (`final_align`) and then makes three
// passes over it (NC, IL/OL, RC).  One GPU thread owns one record here, so
// nothing is materialised.  The work is split in two so that a CTA can regroup
// records between the halves (pantas_aug.cu):
//
//   front_line()    tokenise (REF:142), MAPQ / '*' / dv filters (REF:143-148,172-180),
//                   the three coordinates (REF:151-153), locate the cs tag (REF:154-160)
//                   -> a 32-byte LineRec, classified SIMPLE (cs is one ':' op, i.e. a
//                   perfect match) or GENERAL.  Emits nothing but reject()/error().
//   walk_simple()   path decode + merge walk for a single ':' op: every path node with a
//                   positive length survives, one increment each.
//   walk_general()  the full thing: streaming cs reader, cigar_clipping (REF:40-50), merge
//                   walk (REF:205-255), clear_align/compact_align (REF:63-107) folded on the fly.
//
// Both walks go through the path once with a one-node look-ahead (is this the
// last path node?  REF:217) and a one-node look-behind (was that the last
// *surviving* node?  REF:290,306) and emit every counter update as an event:
//
//   sink.id_to_idx / load_len           node id -> dense index, length       (REF:214)
//   sink.count_node(idx)                NC[idx] += 1                         (REF:263-269)
//   sink.dense(idx, il, ol, stamp, st, has_in, has_out)
//                                       IL[idx][0] += il, OL[idx][len] += ol; has_in / has_out: this occurrence is the
//                                       `to` / the `from` of a link of the same read (REF:298-313,335-351)
//   sink.sparse(idx, dir, pos, stamp)   IL/OL[idx][pos] += 1, deletion-derived keys
//                                                                            (REF:281-297,317-333)
//   sink.edge(from, to, stamp, pf)      RC[(from,to)] += 1                   (REF:357-363)
//   sink.reject()                       rej += 1                             (REF:144-146)
//   sink.error(code, file_offset)       the reference would raise / input not modelled
//
// load_len / load_stamps / prefetch_edge are issued one path step before their
// values are used, so the loads overlap the parsing of the next step.
//
// `stamp` = (file byte offset of the path step) * 4 + e orders first insertions
// exactly like Python's dict insertion order does:
// e = 0 for the j == 0 deletion key, 1 for the dense key, 2 for the j == last
// deletion key.
//
// The same header is compiled by nvcc into the kernels (pantas_aug.cu) and by
// g++ into tests/hostsim (a TEST harness that lets the CPU-only container fuzz
// this logic against the oracle; it is not part of the shipped library).
#pragma once
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define PT_HD __host__ __device__ __forceinline__
#else
#define PT_HD inline
#endif

namespace pt {

// Error codes < 20: the reference raises on this record (nothing on stdout).
// Codes >= 20: the reference would go on, relying on behaviour we refuse to
// guess at (DESIGN.md "documented deviations"); codes >= 40: capacity.
enum : int {
    PT_OK = 0,
    PT_E_COLUMNS = 1,      // < 12 whitespace separated columns   (IndexError REF:143)
    PT_E_MAPQ = 2,         // int(tokens[11]) fails               (ValueError REF:143)
    PT_E_COORD = 3,        // int(tokens[6..8]) fails             (ValueError REF:151-153)
    PT_E_NO_DV = 4,        // no dv:f:<digit> tag                 (ValueError REF:179)
    PT_E_EMPTY_PATH = 5,   // assert len(nodes) > 0               (REF:197)
    PT_E_UNKNOWN_NODE = 6, // KeyError                            (REF:214)
    PT_E_CS_SHORT = 7,     // cigar_vals[0] on an empty list      (IndexError REF:227)
    PT_U_TILDE = 20,       // '~' op: reference reuses a stale length
    PT_U_NON_ASCII = 21,   // byte >= 0x80 (Unicode whitespace/digits not modelled)
    PT_U_BARE_CR = 22,     // lone '\r' (universal-newline line break)
    PT_U_BIG_INT = 23,     // integer beyond the device range
    PT_U_UNDERSCORE = 24,  // int("1_0") is valid Python
    PT_U_POSITION = 25,    // IL/OL position outside the 31-bit key field
    PT_X_NOVEL_FULL = 40,  // novel-edge table full
    PT_X_SPARSE_FULL = 41, // sparse IL/OL table full
    PT_X_DEFER_FULL = 42,  // long-line list full
};

enum : int { LINE_DONE = 0, LINE_DEFER = 1, LINE_SIMPLE = 2, LINE_GENERAL = 3 };

constexpr uint32_t NODE_LEN_ABSENT = 0xFFFFFFFFu;

// str.isspace() for ASCII (str.split() / str.strip() / regex \s):
// \t \n \v \f \r, 0x1c..0x1f, space
PT_HD bool is_ws(uint32_t c) {
    return c == 0x20u || (c - 0x09u) <= 4u || (c - 0x1cu) <= 3u;
}
PT_HD bool is_digit(uint32_t c) { return (c - 0x30u) <= 9u; }
PT_HD bool is_cs_op(uint32_t c) {
    return c == ':' || c == '*' || c == '-' || c == '+' || c == '=' || c == '~';
}

// Fractional digits of the exact midpoint between the double nearest 0.1 and
// the next double above it: float(s) > 0.1  <=>  decimal(s) > 0.1000000000000000124900090...
// (the tie rounds to even, i.e. down to 0.1, so equality is "not greater").
// REF:179.  57 digits.
PT_HD uint32_t dv_midpoint_digit(int k) {
    const char* m = "100000000000000012490009027033011079765856266021728515625";
    return k < 57 ? (uint32_t)(m[k] - '0') : 0u;
}

// ---- SWAR helpers: four bytes per step --------------------------------------

// aligned little-endian word at byte offset w (multiple of 4); s is 4-byte aligned on the
// device, and every buffer is readable up to the next multiple of 16 past `lim`
PT_HD uint32_t ld32(const uint8_t* s, int w) {
#if defined(__CUDA_ARCH__)
    return *reinterpret_cast<const uint32_t*>(s + w);
#else
    uint32_t v;
    memcpy(&v, s + w, 4);
    return v;
#endif
}
PT_HD int first_flag_byte(uint32_t m) {          // index of the lowest byte whose 0x80 flag is set
#if defined(__CUDA_ARCH__)
    return (__ffs((int)m) - 1) >> 3;
#else
    return __builtin_ctz(m) >> 3;
#endif
}
// 0x80 in every byte that is <= 0x20 (ASCII bytes; bytes >= 0x80 may be flagged too, callers verify)
PT_HD uint32_t flag_le20(uint32_t x) { return ~((x | 0x80808080u) - 0x21212121u) & 0x80808080u; }
// 0x80 in every byte equal to the byte replicated in pat4 (exact)
PT_HD uint32_t flag_eq(uint32_t x, uint32_t pat4) {
    const uint32_t y = x ^ pat4;
    return ~(((y & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | y) & 0x80808080u;
}

// first position >= q holding a whitespace byte, or lim
PT_HD int find_ws(const uint8_t* s, int q, int lim) {
    if (q >= lim) return lim;
    int w = q & ~3;
    uint32_t x = ld32(s, w);
    uint32_t m = flag_le20(x) & (0x80808080u << ((q & 3) * 8));
    for (;;) {
        while (m) {
            const int b = first_flag_byte(m);
            if (w + b >= lim) return lim;
            if (is_ws((x >> (8 * b)) & 0xFFu)) return w + b;
            m &= m - 1;
        }
        w += 4;
        if (w >= lim) return lim;
        x = ld32(s, w);
        m = flag_le20(x);
    }
}

struct LineCtx {
    const uint8_t* s;   // base of the addressable bytes (16-byte aligned on the device)
    int lim;            // bytes [0, lim) hold data; reads may touch up to the next multiple of 16
    bool lim_final;     // true: the data really ends at lim (end of chunk)
    int64_t base_off;   // file offset of s[0]
};

// what front_line() hands to the walk
struct LineRec {
    int32_t a5, b5;      // path column [a5, b5)
    union {
        struct { int32_t c0, c1; } cs;   // GENERAL: cs match [c0, c1), c0 < 0 if the tag is absent
        int64_t n;                        // SIMPLE: length of the single ':' op
    };
    int64_t start;       // int(tokens[7])
    int64_t end_rel;     // int(tokens[6]) - int(tokens[8])
};

// [+-]?[0-9]+  -> 1 ok, 0 ValueError, PT_U_* if Python accepts what we do not
PT_HD int parse_py_int(const uint8_t* s, int a, int b, int64_t& out) {
    int q = a;
    bool neg = false;
    if (q < b && (s[q] == '+' || s[q] == '-')) { neg = s[q] == '-'; q++; }
    if (q >= b) return 0;
    int64_t v = 0;
    if (b - q <= 9) {                         // the common case fits 32 bits
        uint32_t v32 = 0;
        for (; q < b; q++) {
            const uint32_t d = (uint32_t)s[q] - '0';
            if (d > 9u) return ((uint32_t)s[q] == '_' && q > a && q + 1 < b) ? PT_U_UNDERSCORE : 0;
            v32 = v32 * 10u + d;
        }
        v = (int64_t)v32;
    } else {
        int sig = 0;
        for (; q < b; q++) {
            const uint32_t c = s[q];
            if (!is_digit(c)) return (c == '_' && q > a && q + 1 < b) ? PT_U_UNDERSCORE : 0;
            if (sig || c != '0') sig++;
            if (sig > 18) return PT_U_BIG_INT;
            v = v * 10 + (int64_t)(c - '0');
        }
    }
    out = neg ? -v : v;
    return 1;
}

// float(text) > 0.1 for the regex match dv:f:(\d+(\.\d+)?) starting at s[d0] (a digit), where the
// bytes [d0, b) are all there is (b = end of the token).  Same exact decimal rule as front_line().
PT_HD bool dv_token_greater(const uint8_t* s, int d0, int b) {
    int t = d0;
    bool int_nonzero = false;
    while (t < b && is_digit(s[t])) { int_nonzero |= (s[t] != '0'); t++; }
    if (int_nonzero) return true;
    if (!(t + 1 < b && s[t] == '.' && is_digit(s[t + 1]))) return false;
    t++;
    const uint32_t f0 = s[t];
    if (f0 == '0') return false;
    if (f0 >= '2') return true;
    int k = 0, cmp = 0;
    while (t < b && is_digit(s[t])) {
        if (cmp == 0) {
            const uint32_t fd = s[t] - '0', md = dv_midpoint_digit(k);
            cmp = fd > md ? 1 : (fd < md ? -1 : 0);
        }
        k++;
        t++;
    }
    if (cmp == 0)
        for (; k < 57; k++)
            if (dv_midpoint_digit(k) != 0) { cmp = -1; break; }
    return cmp > 0;
}

// ---- front half ----------------------------------------------------------------

// Returns LINE_DEFER (nothing emitted) when the record runs past `lim` and lim
// is not the end of the data; LINE_DONE when the record is filtered or in error;
// LINE_SIMPLE / LINE_GENERAL with `rec` filled otherwise.
template <class Sink>
PT_HD int front_line(const LineCtx& cx, int p, int64_t thr, Sink& sink, LineRec& rec) {
    const uint8_t* s = cx.s;
    const int lim = cx.lim;
    const int64_t line_off = cx.base_off + p;

    // ---- tokens 0..11 of line.strip().split()  (REF:142)
    int q = p;
    int a5 = 0, b5 = 0, a6 = 0, b6 = 0, a7 = 0, b7 = 0, a8 = 0, b8 = 0, a11 = 0, b11 = 0;
#pragma unroll 1
    for (int t = 0; t < 12; t++) {
        while (q < lim && s[q] != '\n' && is_ws(s[q])) q++;
        if (q >= lim) {
            if (!cx.lim_final) return LINE_DEFER;
            sink.error(PT_E_COLUMNS, line_off);
            return LINE_DONE;
        }
        if (s[q] == '\n') { sink.error(PT_E_COLUMNS, line_off); return LINE_DONE; }
        const int a = q;
        q = find_ws(s, q + 1, lim);
        if (q >= lim && !cx.lim_final) return LINE_DEFER;
        if (t == 5) { a5 = a; b5 = q; }
        else if (t == 6) { a6 = a; b6 = q; }
        else if (t == 7) { a7 = a; b7 = q; }
        else if (t == 8) { a8 = a; b8 = q; }
        else if (t == 11) { a11 = a; b11 = q; }
    }

    // ---- MAPQ filter, unmapped filter (REF:143-148)
    int64_t mapq;
    int r = parse_py_int(s, a11, b11, mapq);
    if (r != 1) { sink.error(r == 0 ? PT_E_MAPQ : r, line_off); return LINE_DONE; }
    if (mapq < thr) { sink.reject(); return LINE_DONE; }
    if (b5 - a5 == 1 && s[a5] == '*') return LINE_DONE;

    // ---- path length / start / end (REF:151-153)
    int64_t plen, start, pend;
    r = parse_py_int(s, a6, b6, plen);
    if (r != 1) { sink.error(r == 0 ? PT_E_COORD : r, line_off); return LINE_DONE; }
    r = parse_py_int(s, a7, b7, start);
    if (r != 1) { sink.error(r == 0 ? PT_E_COORD : r, line_off); return LINE_DONE; }
    r = parse_py_int(s, a8, b8, pend);
    if (r != 1) { sink.error(r == 0 ? PT_E_COORD : r, line_off); return LINE_DONE; }

    // ---- tags: first "cs:" (to the end of its token) and first "dv:f:<digit>"
    //      (REF:154-160, 172-180).  Both regexes run over " ".join(tokens[12:]);
    //      neither pattern contains whitespace, so a left-to-right scan of the
    //      rest of the line finds the same matches.  Both patterns have ':' as
    //      their third byte: look at ':' bytes and at whitespace only.
    int c0 = -1, c1 = -1, d0 = -1;
    if (q < lim) {
        int w = q & ~3;
        uint32_t x = ld32(s, w);
        uint32_t keep = 0x80808080u << ((q & 3) * 8);
        bool done = false;
        int t_end = lim;                       // where the scan stopped (line end)
        for (;;) {
            uint32_t m = (flag_eq(x, 0x3A3A3A3Au) | flag_le20(x)) & keep;
            while (m) {
                const int b = first_flag_byte(m);
                m &= m - 1;
                const int t = w + b;
                if (t >= lim) { m = 0; w = lim; break; }
                const uint32_t c = (x >> (8 * b)) & 0xFFu;
                if (c == ':') {
                    if (t - 2 >= q && s[t - 1] == 's' && s[t - 2] == 'c') {
                        if (c0 < 0) c0 = t - 2;
                    } else if (d0 < 0 && t - 2 >= q && s[t - 1] == 'v' && s[t - 2] == 'd') {
                        if (t + 4 > lim && !cx.lim_final) return LINE_DEFER;
                        if (t + 4 <= lim && s[t + 1] == 'f' && s[t + 2] == ':' && is_digit(s[t + 3])) d0 = t + 3;
                    }
                } else if (c == '\n') {
                    t_end = t;
                    done = true;
                    break;
                } else if (is_ws(c)) {
                    if (c0 >= 0 && c1 < 0) c1 = t;
                }
                if (c0 >= 0 && c1 >= 0 && d0 >= 0) { done = true; break; }
            }
            if (done) break;
            w += 4;
            if (w >= lim) {
                if (!cx.lim_final) return LINE_DEFER;
                break;
            }
            x = ld32(s, w);
            keep = 0x80808080u;
        }
        if (c0 >= 0 && c1 < 0) c1 = t_end;      // token ran to the end of the line
    }

    // ---- dv filter: float(dv) > 0.1 -> skip (REF:172-180), exact on the decimal text
    if (d0 < 0) { sink.error(PT_E_NO_DV, line_off); return LINE_DONE; }
    {
        int t = d0;
        bool int_nonzero = false;
        while (t < lim && is_digit(s[t])) { int_nonzero |= (s[t] != '0'); t++; }
        if (t >= lim && !cx.lim_final) return LINE_DEFER;
        bool greater = int_nonzero;
        if (!greater && t + 1 < lim && s[t] == '.' && is_digit(s[t + 1])) {
            t++;
            const uint32_t f0 = s[t];
            if (f0 == '0') {
                // 0.0xxx is never greater, whatever follows
            } else if (f0 >= '2') {
                greater = true;
            } else {
                int k = 0;
                int cmp = 0;                   // sign of (fraction - midpoint) so far
                while (t < lim && is_digit(s[t])) {
                    if (cmp == 0) {
                        const uint32_t fd = s[t] - '0', md = dv_midpoint_digit(k);
                        cmp = fd > md ? 1 : (fd < md ? -1 : 0);
                    }
                    k++;
                    t++;
                }
                if (t >= lim && !cx.lim_final) return LINE_DEFER;
                if (cmp == 0)                  // remaining midpoint digits vs implicit zeros
                    for (; k < 57; k++)
                        if (dv_midpoint_digit(k) != 0) { cmp = -1; break; }
                greater = cmp > 0;
            }
        } else if (!greater && t + 1 >= lim && !cx.lim_final && t < lim && s[t] == '.') {
            return LINE_DEFER;
        }
        if (greater) return LINE_DONE;
    }

    rec.a5 = a5;
    rec.b5 = b5;
    rec.start = start;
    rec.end_rel = plen - pend;                                        // REF:153

    // ---- perfect match?  the match is exactly "cs:Z::<digits>"  -> one op (':', n)
    if (c0 >= 0 && c1 - c0 >= 7 && s[c0 + 3] == 'Z' && s[c0 + 4] == ':' && s[c0 + 5] == ':') {
        int t = c0 + 6;
        int sig = 0;
        int64_t n = 0;
        for (; t < c1; t++) {
            const uint32_t d = (uint32_t)s[t] - '0';
            if (d > 9u) break;
            if (sig || d) sig++;
            n = n * 10 + (int64_t)d;
            if (sig > 15) break;
        }
        if (t == c1 && sig <= 15) {
            rec.n = n;
            return LINE_SIMPLE;
        }
    }
    rec.cs.c0 = c0;
    rec.cs.c1 = c1;
    return LINE_GENERAL;
}

// ---- path column ------------------------------------------------------------------

struct PathIter {
    const uint8_t* s;
    int pq, b5;
    uint32_t sep;
    bool rev;
};

// REF:185-197: path[0] == '>' ? split('>') : split('<'); the first piece is dropped
PT_HD bool path_begin(PathIter& it, const uint8_t* s, int a5, int b5) {
    it.s = s;
    it.b5 = b5;
    it.rev = s[a5] != '>';
    it.sep = it.rev ? '<' : '>';
    int pq = a5;
    while (pq < b5 && s[pq] != it.sep) pq++;
    it.pq = pq;
    return pq < b5;
}
PT_HD bool path_more(const PathIter& it) { return it.pq < it.b5; }

// One piece starting at s[pq] == sep.  Only canonical decimal ids can name a
// node (the host refuses graphs with any other S id), so every other spelling is
// the reference's KeyError.  s[b5] is whitespace, which ends the digit loop.
template <class Sink>
PT_HD bool next_piece(PathIter& it, Sink& sink, uint32_t& idx, int& at) {
    const uint8_t* s = it.s;
    int pq = it.pq + 1;
    at = pq;
    uint32_t v = 0;
    uint32_t d = (uint32_t)s[pq] - '0';
    const uint32_t d0 = d;
    while (d <= 9u && pq - at < 9) {
        v = v * 10u + d;
        pq++;
        d = (uint32_t)s[pq] - '0';
    }
    uint64_t id = v;
    if (d <= 9u) {                               // a 10th digit (ids up to 2^32 - 2)
        id = id * 10u + d;
        pq++;
        d = (uint32_t)s[pq] - '0';
    }
    const int nd = pq - at;
    const bool ends = pq >= it.b5 || (uint32_t)s[pq] == it.sep;
    if (!ends) {                                 // junk inside the piece: skip to its end, report invalid
        while (pq < it.b5 && (uint32_t)s[pq] != it.sep) pq++;
        it.pq = pq;
        return false;
    }
    it.pq = pq;
    if (nd == 0 || (nd > 1 && d0 == 0)) return false;
    return sink.id_to_idx(id, idx);
}

// ---- look-behind: the last surviving node waits to learn whether it is the last one ----

template <class Stamps>
struct PendSimple {
    bool valid, is_first;
    uint32_t idx;
    uint64_t stamp;
    Stamps st;
};

template <class Sink>
PT_HD void walk_simple(const LineCtx& cx, const LineRec& rec, Sink& sink) {
    typedef typename Sink::Stamps Stamps;
    typedef typename Sink::EdgePf EdgePf;
    const uint8_t* s = cx.s;
    const int64_t line_off = cx.base_off + rec.a5;
    PathIter it;
    if (!path_begin(it, s, rec.a5, rec.b5)) { sink.error(PT_E_EMPTY_PATH, line_off); return; }
    const bool rev = it.rev;

    uint32_t cur_idx = 0, nxt_idx = 0;
    int cur_at = 0, nxt_at = 0;
    if (!next_piece(it, sink, cur_idx, cur_at)) { sink.error(PT_E_UNKNOWN_NODE, cx.base_off + cur_at); return; }
    uint32_t cur_len = sink.load_len(cur_idx);
    uint32_t nxt_len = 0;

    int64_t rem = rec.n;          // what is left of the single ':' op
    bool head = true;             // op list not yet exhausted
    bool first_node = true;
    bool any_survivor = false;
    PendSimple<Stamps> pd;
    pd.valid = false;
    pd.is_first = false;
    pd.idx = 0;
    pd.stamp = 0;
    EdgePf pf;
    sink.edge_pf_init(pf);

    for (;;) {
        // look ahead to the next *distinct* piece (REF:188)
        bool have_nxt = false;
        while (path_more(it)) {
            if (!next_piece(it, sink, nxt_idx, nxt_at)) { sink.error(PT_E_UNKNOWN_NODE, cx.base_off + nxt_at); return; }
            if (nxt_idx != cur_idx) { have_nxt = true; break; }
        }
        if (cur_len == NODE_LEN_ABSENT) { sink.error(PT_E_UNKNOWN_NODE, cx.base_off + cur_at); return; }

        int64_t L = (int64_t)cur_len;
        if (first_node) L -= rec.start;                 // REF:215-216
        if (!have_nxt) L = L - rec.end_rel + 1;         // REF:217-218
        first_node = false;

        if (L > 0) {
            if (!head) { sink.error(PT_E_CS_SHORT, line_off); return; }   // REF:227
            if (L < rem) rem -= L;                      // REF:234-243
            else head = false;                          // op used up (REF:244-255)
            // slice is [(':', take)]: never dropped, one counting op (REF:97-107, 298)
            const uint64_t stamp = (uint64_t)(cx.base_off + cur_at) << 2;
            if (pd.valid) {
                const bool not_first = !pd.is_first;
                sink.dense(pd.idx, (rev || not_first) ? 1 : 0, (!rev || not_first) ? 1 : 0, pd.stamp | 1u, pd.st, rev || not_first,
                           !rev || not_first);
                if (rev) sink.edge(cur_idx, pd.idx, stamp, pf);
                else sink.edge(pd.idx, cur_idx, stamp, pf);
            }
            sink.count_node(cur_idx);                    // REF:263-269
            pd.valid = true;
            pd.is_first = !any_survivor;
            pd.idx = cur_idx;
            pd.stamp = stamp;
            pd.st = sink.load_stamps(cur_idx);           // compared when this node is flushed
            any_survivor = true;
        }
        if (!have_nxt) break;
        // issue the next step's loads now; they are consumed after the next piece is parsed
        nxt_len = sink.load_len(nxt_idx);
        sink.prefetch_edge(pf, rev ? nxt_idx : cur_idx);   // home slots of edge (cur,nxt) / (nxt,cur)
        cur_idx = nxt_idx;
        cur_len = nxt_len;
        cur_at = nxt_at;
    }
    if (pd.valid) {
        // last surviving node: i == last; i != 0 unless it is also the first
        const bool not_first = !pd.is_first;
        sink.dense(pd.idx, (!rev && not_first) ? 1 : 0, (rev && not_first) ? 1 : 0, pd.stamp | 1u, pd.st, !rev && not_first, rev && not_first);
    }
}

// ---- general walk -------------------------------------------------------------------

// Streaming reader of the cs difference string, REF:154-167 + parse_cigar REF:10-37.
struct OpReader {
    const uint8_t* s;
    int q, end;          // unread part of the cs token (stream mode)
    int n_pre;           // ops held in registers (absent tag / exactly-two-op case)
    int i_pre;
    uint8_t pre_op[2];
    int64_t pre_len[2];

    // the cs token with every "cs:Z:" removed (.replace("cs:Z:", ""), REF:158)
    PT_HD bool next_char(uint32_t& c) {
        while (q < end) {
            if (s[q] == 'c' && q + 5 <= end && s[q + 1] == 's' && s[q + 2] == ':' && s[q + 3] == 'Z' &&
                s[q + 4] == ':') {
                q += 5;
                continue;
            }
            c = s[q];
            return true;
        }
        return false;
    }

    // one (op, len) from the string; 1 ok, 0 exhausted, PT_U_BIG_INT
    PT_HD int parse_one(uint8_t& op, int64_t& len) {
        uint32_t c;
        for (;;) {                       // text before the first operator is ignored (curr_op is None)
            if (!next_char(c)) return 0;
            q++;
            if (is_cs_op(c)) break;
        }
        op = (uint8_t)c;
        int64_t count = 0, val = 0;
        int sig = 0;
        bool alldig = true;
        while (next_char(c)) {
            if (is_cs_op(c)) break;
            q++;
            count++;
            if (is_digit(c)) {
                if (sig || c != '0') sig++;
                if (sig <= 15) val = val * 10 + (int64_t)(c - '0');
            } else {
                alldig = false;
            }
        }
        if (op == '*') len = 1;                          // REF:28-29
        else if (alldig && count > 0) {                  // REF:31-32
            if (sig > 15) return PT_U_BIG_INT;
            len = val;
        } else len = count;                              // REF:33-34
        return 1;
    }

    PT_HD int fetch(uint8_t& op, int64_t& len) {
        if (i_pre < n_pre) {
            op = pre_op[i_pre];
            len = pre_len[i_pre];
            i_pre++;
            return 1;
        }
        if (n_pre) return 0;
        return parse_one(op, len);
    }
};

template <class Stamps>
struct Pending {
    bool valid;
    bool is_first;       // i == 0 in final_align
    bool first_del, last_del;
    uint32_t idx, len;
    int64_t n_count;     // ops that are neither '-' nor '*' in the compacted slice
    int64_t first_len, last_len;
    uint64_t stamp;      // file offset of the step << 2
    Stamps st;
};

template <class Sink, class Stamps>
PT_HD void flush_pending(const Pending<Stamps>& p, bool is_last, bool rev, Sink& sink) {
    const bool not_first = !p.is_first, not_last = !is_last;
    // which end condition guards which dictionary (REF:280-353)
    const bool il_cond = rev ? not_last : not_first;
    const bool ol_cond = rev ? not_first : not_last;
    int64_t il_touch = il_cond ? p.n_count : 0;
    int64_t ol_touch = ol_cond ? p.n_count : 0;
    if (!rev) {
        if (p.first_del && not_first) {                                   // REF:282-289
            if (p.first_len == 0) il_touch++;                             // same key as the dense one
            else sink.sparse(p.idx, 0, p.first_len, p.stamp | 0u);
        }
        if (p.last_del && not_last)                                       // REF:290-297
            sink.sparse(p.idx, 1, (int64_t)p.len - p.last_len - 1, p.stamp | 2u);
    } else {
        if (p.first_del && not_first)                                     // REF:318-325
            sink.sparse(p.idx, 1, (int64_t)p.len - 1 - p.first_len, p.stamp | 0u);
        if (p.last_del && not_last) {                                     // REF:326-333
            if (p.last_len == 0) il_touch++;
            else sink.sparse(p.idx, 0, p.last_len, p.stamp | 2u);
        }
    }
    sink.dense(p.idx, il_touch, ol_touch, p.stamp | 1u, p.st, il_cond, ol_cond);
}

template <class Sink>
PT_HD void walk_general(const LineCtx& cx, const LineRec& rec, Sink& sink) {
    typedef typename Sink::Stamps Stamps;
    typedef typename Sink::EdgePf EdgePf;
    const uint8_t* s = cx.s;
    const int64_t line_off = cx.base_off + rec.a5;
    const int c0 = rec.cs.c0, c1 = rec.cs.c1;

    // ---- cs ops (REF:154-167)
    OpReader ops;
    ops.s = s;
    ops.n_pre = 0;
    ops.i_pre = 0;
    ops.pre_op[0] = ops.pre_op[1] = 0;
    ops.pre_len[0] = ops.pre_len[1] = 0;
    int64_t start_pos = rec.start;
    if (c0 < 0) {                                   // cigar = "*" -> [('*', 1)]  (REF:160)
        ops.q = ops.end = 0;
        ops.n_pre = 1;
        ops.pre_op[0] = '*';
        ops.pre_len[0] = 1;
    } else {
        ops.q = c0;
        ops.end = c1;
        int nops = 0;
        uint32_t c;
        while (ops.next_char(c)) { nops += is_cs_op(c) ? 1 : 0; ops.q++; }
        ops.q = c0;
        if (nops == 2) {                            // cigar_clipping, REF:40-50,164-167
            int r0 = ops.parse_one(ops.pre_op[0], ops.pre_len[0]);
            int r1 = ops.parse_one(ops.pre_op[1], ops.pre_len[1]);
            if (r0 != 1 || r1 != 1) { sink.error(PT_U_BIG_INT, line_off); return; }
            ops.n_pre = 2;
            if (ops.pre_op[0] == '+' && ops.pre_op[1] == ':') {
                start_pos += ops.pre_len[0];
                ops.pre_op[0] = ops.pre_op[1];
                ops.pre_len[0] = ops.pre_len[1];
                ops.n_pre = 1;
            } else if (ops.pre_op[0] == ':' && ops.pre_op[1] == '+') {
                ops.n_pre = 1;
            }
        } else if (nops == 0) {
            ops.q = ops.end;                        // empty op list
        }
    }

    // ---- path decode (REF:185-197) fused with the merge walk (REF:205-255),
    //      clear_align/compact_align (REF:63-107) and the three accumulations.
    PathIter it;
    if (!path_begin(it, s, rec.a5, rec.b5)) { sink.error(PT_E_EMPTY_PATH, line_off); return; }
    const bool rev = it.rev;
    uint32_t cur_idx = 0, nxt_idx = 0;
    int cur_at = 0, nxt_at = 0;
    if (!next_piece(it, sink, cur_idx, cur_at)) { sink.error(PT_E_UNKNOWN_NODE, cx.base_off + cur_at); return; }
    uint32_t cur_len = sink.load_len(cur_idx);
    uint32_t nxt_len = 0;

    bool head_valid = false;
    uint8_t head_op = 0;
    int64_t head_rem = 0;

    Pending<Stamps> pend_node;
    pend_node.valid = false;
    pend_node.is_first = false;
    pend_node.first_del = pend_node.last_del = false;
    pend_node.idx = pend_node.len = 0;
    pend_node.n_count = pend_node.first_len = pend_node.last_len = 0;
    pend_node.stamp = 0;
    bool any_survivor = false;
    bool first_node = true;
    EdgePf pf;
    sink.edge_pf_init(pf);

    for (;;) {
        bool have_nxt = false;
        while (path_more(it)) {
            if (!next_piece(it, sink, nxt_idx, nxt_at)) { sink.error(PT_E_UNKNOWN_NODE, cx.base_off + nxt_at); return; }
            if (nxt_idx != cur_idx) { have_nxt = true; break; }
        }
        if (have_nxt) nxt_len = sink.load_len(nxt_idx);
        if (cur_len == NODE_LEN_ABSENT) { sink.error(PT_E_UNKNOWN_NODE, cx.base_off + cur_at); return; }

        int64_t L = (int64_t)cur_len;
        if (first_node) L -= start_pos;                 // REF:215-216
        if (!have_nxt) L = L - rec.end_rel + 1;         // REF:217-218
        first_node = false;

        if (L > 0) {
            int nP = 0, nQ = 0;
            uint8_t p0_op = 0, qlast_op = 0, first_op = 0;
            int64_t qlast_len = 0, first_len = 0, n_count = 0;
            while (L > 0) {
                if (!head_valid) {
                    int fr = ops.fetch(head_op, head_rem);
                    if (fr == 0) {
                        if (nP == 0) { sink.error(PT_E_CS_SHORT, line_off); return; }   // REF:227
                        break;                                                           // REF:252-255
                    }
                    if (fr != 1) { sink.error(fr, line_off); return; }
                    head_valid = true;
                }
                if (head_op == '~') { sink.error(PT_U_TILDE, line_off); return; }
                int64_t take;
                if (L <= head_rem) {                    // REF:234-243
                    take = L;
                    head_rem -= L;
                    if (head_rem == 0) head_valid = false;
                    L = 0;
                } else {                                // REF:244-251
                    take = head_rem;
                    L -= head_rem;
                    head_valid = false;
                }
                // compact_align as a running fold (REF:63-94)
                bool push = false;
                int64_t push_len = take;
                if (nP == 0) {
                    p0_op = head_op;
                    push = head_op != '*';
                } else if (nQ == 0) {
                    push = true;
                    push_len = take + 1;
                } else if (head_op == qlast_op || head_op == '*') {
                    qlast_len += take;
                } else {
                    push = true;
                }
                if (push) {
                    if (nQ == 1) { first_op = qlast_op; first_len = qlast_len; }
                    qlast_op = head_op;
                    qlast_len = push_len;
                    nQ++;
                    if (head_op != '-' && head_op != '*') n_count++;
                }
                nP++;
            }
            if (nQ == 1) { first_op = qlast_op; first_len = qlast_len; }
            const bool dropped = (nP == 1 && (p0_op == '-' || p0_op == '+'));   // REF:101-102
            if (!dropped) {
                const uint64_t stamp = (uint64_t)(cx.base_off + cur_at) << 2;
                if (pend_node.valid) {
                    flush_pending(pend_node, false, rev, sink);
                    if (rev) sink.edge(cur_idx, pend_node.idx, stamp, pf);          // REF:357-363
                    else sink.edge(pend_node.idx, cur_idx, stamp, pf);
                }
                sink.count_node(cur_idx);                                           // REF:263-269
                pend_node.valid = true;
                pend_node.is_first = !any_survivor;
                pend_node.idx = cur_idx;
                pend_node.len = cur_len;
                pend_node.n_count = n_count;
                pend_node.first_del = nQ > 0 && first_op == '-';
                pend_node.last_del = nQ > 0 && qlast_op == '-';
                pend_node.first_len = first_len;
                pend_node.last_len = qlast_len;
                pend_node.stamp = stamp;
                pend_node.st = sink.load_stamps(cur_idx);
                any_survivor = true;
            }
        }
        if (!have_nxt) break;
        cur_idx = nxt_idx;
        cur_len = nxt_len;
        cur_at = nxt_at;
    }
    if (pend_node.valid) flush_pending(pend_node, true, rev, sink);
}

// front + walk in one go (records taken one at a time: the long-record kernel, tests)
template <class Sink>
PT_HD int process_line(const LineCtx& cx, int p, int64_t thr, Sink& sink) {
    LineRec rec;
    const int r = front_line(cx, p, thr, sink, rec);
    if (r == LINE_SIMPLE) { walk_simple(cx, rec, sink); return LINE_DONE; }
    if (r == LINE_GENERAL) { walk_general(cx, rec, sink); return LINE_DONE; }
    return r;
}

}  // namespace pt
