// gfa_kernels.cuh -- the two GFA passes of `pantas augment` on the device (included by pantas_aug.cu).
//
// Reference: /root/reference/scripts/alignments_augmentation_from_gaf.py (REF:n)
//   pass 1  REF:121-126  nodes_info from the S lines          -> gfa_parse_kernel (per line: kind, tokens, id, sequence length)
//   pass 2  REF:377-427  echo every line with its tags        -> gfa_measure_kernel (bytes each line prints), gfa_format_kernel
//
// One thread per GFA line (lines are ~35 bytes: the passes are latency- and launch-bound, not bandwidth-bound; they
// matter because after the GAF loop moved to the GPU they were > 99 % of the wall time of `./pantas augment`).
// Line boundaries, prefix sums and the de-duplication of link keys are index plumbing done with torch by the caller
// (pantas_b200/gfa_device.py); everything that looks at GFA bytes or formats output bytes is here.
//
// Python semantics kept: str.strip() / str.split() on ASCII whitespace (\t \n \v \f \r 0x1c-0x1f space); pass 1 tests the
// RAW line for a leading "S" (REF:123), pass 2 the stripped one (REF:379); a segment id is looked up as a string, so only
// canonical decimal spellings can name a node (anything else in an S line: UnsupportedInput, like the host loader).
#pragma once

namespace gfa {

enum : uint32_t {
    K_RAW_S = 1u,        // raw line starts with 'S'            (REF:123: enters nodes_info)
    K_STR_S = 2u,        // stripped line starts with 'S'       (REF:379)
    K_STR_L = 4u,        // stripped line starts with 'L'       (REF:417)
    K_NTOK_SHIFT = 4,    // bits 4..6: number of whitespace-separated tokens, capped at 4
    K_NTOK_MASK = 7u << 4,
};
constexpr uint32_t ID_INVALID = 0xFFFFFFFFu;     // token is not a canonical decimal <= 0xFFFFFFFE
// error codes of the GFA passes (reported as line_number << 8 | code, smallest line first)
enum : int { GE_S_FIELDS = 1 /* REF:124-125 IndexError */, GE_S_ID = 2 /* id spelling not supported */, GE_S_LONG = 3 /* >= 2^30 bases */,
             GE_W_S_ID = 4 /* REF:381 IndexError */, GE_W_S_KEY = 5 /* REF:382 KeyError */, GE_W_L_FIELDS = 6 /* REF:421 IndexError */ };

__device__ __forceinline__ bool py_ws(uint32_t c) { return c == 0x20u || (c - 0x09u) <= 4u || (c - 0x1cu) <= 3u; }

// canonical decimal in [a, b): digits only, no leading zero unless "0", at most 10 digits, value <= 0xFFFFFFFE
__device__ __forceinline__ uint32_t canonical_id(const uint8_t* s, uint64_t a, uint64_t b) {
    const uint64_t n = b - a;
    if (n == 0 || n > 10) return ID_INVALID;
    if (n > 1 && s[a] == '0') return ID_INVALID;
    uint64_t v = 0;
    for (uint64_t q = a; q < b; q++) {
        const uint32_t d = (uint32_t)s[q] - '0';
        if (d > 9u) return ID_INVALID;
        v = v * 10u + d;
    }
    return v <= 0xFFFFFFFEull ? (uint32_t)v : ID_INVALID;
}

struct LineArrays {
    const long long* start;     // [n_lines + 1] byte offset of every line; start[n_lines] = one past the last line's break
    const long long* end;       // [n_lines] end of the line's text (before its line break)
    uint32_t* a_rel;            // stripped text starts at start + a_rel
    uint32_t* slen;             // ... and is slen bytes long
    uint32_t* kind;             // K_* flags
    uint32_t* v1;               // token 1 as a canonical id (S: the segment, L: from)
    uint32_t* v2;               // S: len(token 2); L: token 3 as a canonical id
};

__global__ void gfa_parse_kernel(const uint8_t* s, LineArrays L, uint64_t n_lines, unsigned long long* err) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n_lines; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t ls = (uint64_t)L.start[i], le = (uint64_t)L.end[i];
        uint64_t a = ls, b = le;
        while (a < b && py_ws(s[a])) a++;
        while (b > a && py_ws(s[b - 1])) b--;
        uint32_t kind = 0;
        if (ls < le && s[ls] == 'S') kind |= K_RAW_S;
        if (a < b && s[a] == 'S') kind |= K_STR_S;
        if (a < b && s[a] == 'L') kind |= K_STR_L;
        // tokens of the stripped text -- of S and L lines only: nothing reads them for other lines, and the P line of a
        // reference path is one token of up to 100 MB
        uint64_t ta[4] = {0, 0, 0, 0}, tb[4] = {0, 0, 0, 0};
        uint32_t nt = 0;
        uint64_t q = a;
        while (kind != 0u && q < b && nt < 4u) {
            ta[nt] = q;
            while (q < b && !py_ws(s[q])) q++;
            tb[nt] = q;
            nt++;
            while (q < b && py_ws(s[q])) q++;
        }
        kind |= nt << K_NTOK_SHIFT;
        uint32_t v1 = ID_INVALID, v2 = 0;
        if (kind & (K_RAW_S | K_STR_S)) {
            if (nt >= 2u) v1 = canonical_id(s, ta[1], tb[1]);
            if (nt >= 3u) {
                const uint64_t sl = tb[2] - ta[2];
                v2 = sl >= (1ull << 30) ? 0xFFFFFFFFu : (uint32_t)sl;
            }
            if (kind & K_RAW_S) {                                        // REF:123-126
                if (nt < 3u) atomicMin(err, ((unsigned long long)i << 8) | GE_S_FIELDS);
                else if (v1 == ID_INVALID) atomicMin(err, ((unsigned long long)i << 8) | GE_S_ID);
                else if (v2 == 0xFFFFFFFFu) atomicMin(err, ((unsigned long long)i << 8) | GE_S_LONG);
            }
        } else if (kind & K_STR_L) {
            if (nt >= 2u) v1 = canonical_id(s, ta[1], tb[1]);
            v2 = nt >= 4u ? canonical_id(s, ta[3], tb[3]) : ID_INVALID;
        }
        L.a_rel[i] = (uint32_t)(a - ls);
        L.slen[i] = (uint32_t)(b - a);
        L.kind[i] = kind;
        L.v1[i] = v1;
        L.v2[i] = v2;
    }
}

// what the writer needs besides the parsed lines
struct WriterArgs {
    const uint8_t* s;             // GFA bytes
    const long long* start;
    const uint32_t* a_rel;
    const uint32_t* slen;
    const uint32_t* kind;
    const uint32_t* v1;
    const int32_t* link_edge;     // per line: edge index an L line prints, -1: prints 0 (REF:421 weights.pop)
    const uint32_t* node_len;     // [n_nodes], 0xFFFFFFFF = no such node
    const long long* sums;        // pt_export_dense layout: [nc | il_adj | ol_adj | rc | ...]
    const int32_t* sp_slot;       // [n_nodes] index of the node's preformatted IL/OL tags (deletion-derived keys), -1: none
    const long long* sp_off;      // [n_sp + 1] offsets into sp_text
    const uint8_t* sp_text;
    uint64_t n_nodes;
    uint64_t n_lines;
    uint32_t min_id;
};

__device__ __forceinline__ uint32_t dec_digits(unsigned long long v) {
    uint32_t n = 1;
    while (v >= 10ull) { v /= 10ull; n++; }
    return n;
}
__device__ __forceinline__ uint8_t* put_dec(uint8_t* o, unsigned long long v) {
    const uint32_t n = dec_digits(v);
    for (uint32_t k = n; k-- > 0;) { o[k] = (uint8_t)('0' + v % 10ull); v /= 10ull; }
    return o + n;
}
__device__ __forceinline__ uint8_t* put_str(uint8_t* o, const char* t, int n) {
    for (int k = 0; k < n; k++) o[k] = (uint8_t)t[k];
    return o + n;
}

// other lines (REF:424: echoed stripped) longer than this are copied by a whole block, not by the line's one thread
constexpr uint32_t LONG_LINE = 4096;

// bytes line i prints (0: the line is dropped); errors like the reference's second pass (REF:377-424)
template <bool WRITE>
__device__ __forceinline__ uint64_t gfa_line_out(const WriterArgs& W, uint64_t i, uint8_t* o, unsigned long long* err) {
    const uint32_t kind = W.kind[i], slen = W.slen[i], nt = (kind & K_NTOK_MASK) >> K_NTOK_SHIFT;
    const uint8_t* src = W.s + (uint64_t)W.start[i] + W.a_rel[i];
    uint8_t* const o0 = o;
    if (kind & K_STR_S) {
        if (nt < 2u) { atomicMin(err, ((unsigned long long)i << 8) | GE_W_S_ID); return 0; }
        const uint32_t id = W.v1[i];
        const uint64_t idx = (uint64_t)id - W.min_id;
        if (id == ID_INVALID || id < W.min_id || idx >= W.n_nodes || W.node_len[idx] == 0xFFFFFFFFu) {
            atomicMin(err, ((unsigned long long)i << 8) | GE_W_S_KEY);
            return 0;
        }
        if (nt < 3u) return 0;                                           // REF:395-416: neither branch prints
        const unsigned long long nc = (unsigned long long)W.sums[idx];
        const unsigned long long il = (unsigned long long)(W.sums[idx] + W.sums[W.n_nodes + idx]);
        const unsigned long long ol = (unsigned long long)(W.sums[idx] + W.sums[2 * W.n_nodes + idx]);
        const int32_t sp = W.sp_slot[idx];
        uint64_t n = (uint64_t)slen + 6u + dec_digits(nc) + 1u;
        if (sp >= 0) n += (uint64_t)(W.sp_off[sp + 1] - W.sp_off[sp]);
        else {
            if (il) n += 8u + dec_digits(il);                            // "\tIL:Z:0." + count
            if (ol) n += 7u + dec_digits(W.node_len[idx]) + dec_digits(ol);
        }
        if (WRITE) {
            for (uint32_t k = 0; k < slen; k++) o[k] = src[k];
            o = put_str(o + slen, "\tNC:i:", 6);
            o = put_dec(o, nc);
            if (sp >= 0) {
                const uint8_t* t = W.sp_text + W.sp_off[sp];
                const uint64_t tn = (uint64_t)(W.sp_off[sp + 1] - W.sp_off[sp]);
                for (uint64_t k = 0; k < tn; k++) o[k] = t[k];
                o += tn;
            } else {
                if (il) { o = put_str(o, "\tIL:Z:0.", 8); o = put_dec(o, il); }
                if (ol) { o = put_str(o, "\tOL:Z:", 6); o = put_dec(o, W.node_len[idx]); *o++ = '.'; o = put_dec(o, ol); }
            }
            *o++ = '\n';
            (void)o0;
        }
        return n;
    }
    if (kind & K_STR_L) {
        if (slen == 1u) return 0;                                        // REF:418-419
        if (nt < 4u) { atomicMin(err, ((unsigned long long)i << 8) | GE_W_L_FIELDS); return 0; }
        const int32_t e = W.link_edge[i];
        const unsigned long long w = e >= 0 ? (unsigned long long)W.sums[3 * W.n_nodes + (uint64_t)e] : 0ull;
        if (WRITE) {
            for (uint32_t k = 0; k < slen; k++) o[k] = src[k];
            o = put_str(o + slen, "\tRC:i:", 6);
            o = put_dec(o, w);
            *o++ = '\n';
        }
        return (uint64_t)slen + 6u + dec_digits(w) + 1u;
    }
    if (WRITE) {
        if (slen <= LONG_LINE)                                           // (longer: gfa_format_kernel's second loop)
            for (uint32_t k = 0; k < slen; k++) o[k] = src[k];
        o[slen] = '\n';
    }
    return (uint64_t)slen + 1u;                                          // REF:424
}

__global__ void gfa_measure_kernel(WriterArgs W, long long* out_len, unsigned long long* err) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < W.n_lines; i += (uint64_t)gridDim.x * blockDim.x)
        out_len[i] = (long long)gfa_line_out<false>(W, i, nullptr, err);
}
// out_off[i] = exclusive prefix sum of out_len (the caller's cumsum)
__global__ void gfa_format_kernel(WriterArgs W, const long long* out_off, uint8_t* out, unsigned long long* err) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < W.n_lines; i += (uint64_t)gridDim.x * blockDim.x)
        gfa_line_out<true>(W, i, out + out_off[i], err);
    // long echoed lines (the P line of a reference path can be 100 MB): the block that owns the line copies it together
    __shared__ uint32_t long_n;
    __shared__ uint32_t long_list[256];
    for (uint64_t base = blockIdx.x * (uint64_t)blockDim.x; base < W.n_lines; base += (uint64_t)gridDim.x * blockDim.x) {
        if (threadIdx.x == 0) long_n = 0;
        __syncthreads();
        const uint64_t i = base + threadIdx.x;
        if (i < W.n_lines && (W.kind[i] & (K_STR_S | K_STR_L)) == 0u && W.slen[i] > LONG_LINE) long_list[atomicAdd(&long_n, 1u)] = threadIdx.x;
        __syncthreads();
        const uint32_t n = long_n;
        for (uint32_t k = 0; k < n; k++) {
            const uint64_t j = base + long_list[k];
            const uint8_t* src = W.s + (uint64_t)W.start[j] + W.a_rel[j];
            uint8_t* dst = out + out_off[j];
            const uint32_t len = W.slen[j];
            for (uint32_t q = threadIdx.x; q < len; q += blockDim.x) dst[q] = src[q];
        }
        __syncthreads();
    }
}

}  // namespace gfa
