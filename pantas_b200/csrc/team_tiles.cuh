// team_tiles.cuh -- the fast path of the augment kernel (included by aug_kernels.cuh after tables.cuh, the TMA
// helpers, ChunkArgs and defer_line(); compiled for sm_100a by pantas_aug.cu and for the CPU emulator by
// tests/hostsim/fastsim.cpp).
//
// Reference loop body: /root/reference/scripts/alignments_augmentation_from_gaf.py:142-363 (REF:n).
//
// The unit of execution is a TEAM: two warps that own tiles of the GAF chunk -- ChunkArgs::tile_bytes each, about 30 records,
// at most G::TILE = 9 KiB (+ 1 KiB of look-ahead: a record belongs to the tile it starts in) -- and their own slice of shared memory.  G::NT teams form one CTA (default: the ten
// teams of an SM are ONE 640-thread CTA) and share nothing but the global tables and ONE CTA-wide barrier per tile,
// after `scan`: it keeps every team of the SM in (nearly) the same phase, so that the SM's instruction caches hold one
// or two phases of this 100 KB kernel instead of all of them (independent 64-thread CTAs starved on instruction fetch).
// Inside a tile the phases are separated by team-level named barriers (bar.sync team + 1, 64; ChunkArgs::loose = 0 makes
// every barrier CTA-wide).  One elected thread per team moves the tile with a 1-D TMA bulk copy (UBLKCP, L2 evict-first);
// the copy of the team's next tile is issued as soon as the bytes are dead (after `ids`).  Phases over a tile:
//
//   scan     every thread takes 4 x 16 bytes per round (LDS.128, conflict-free), branch-free SWAR on 32-bit words: per
//            16-byte vector a 16-bit whitespace mask (bytes <= 0x20) and a 16-bit mask of path separators ('>' '<') OR
//            non-tab whitespace ('\n' rides along for free: sep & ws = record-end candidates).  Record starts are read
//            back from the mask words and ranked with a warp scan (no atomics) into a two-ended list: warp 0 fills it
//            from the front, warp 1 from the back.
//   records  warp 0 = role B for every record, warp 1 = role A (a warp takes as long for one record as for 32; ~27
//            records per tile keep both warps' lanes busy):
//              B  the 12 column boundaries (single tabs, no empty column), MAPQ and '*' filters (REF:143-148), the three
//                 coordinates (REF:151-153), then one step-list entry per separator bit of the path column (+ a
//                 sentinel); list space comes from a warp prefix sum;
//              A  skips ten boundaries by popcount; first cs token, first dv:f: token (REF:154-160,172-180), dv filter,
//                 cs string parsed into the tile's op pool (REF:10-50, incl. cigar_clipping).
//            Anything unusual hands the record to the exact per-record path (line_core.cuh, augment_deferred_kernel).
//   ids      one thread per path step: SWAR decimal parse of the id out of shared memory -> node index -> the node's
//            meta word from the L2 (four loads in flight per thread); keeps index, meta word and the node's length in
//            shared memory, flags consecutive duplicates (REF:190-196) and adds the length to the record's sum.
//   walk     warp 0 the even records, warp 1 the odd ones, O(1) for a record whose cs string is one op: first / last node
//            shortened (REF:215-218), the sum of the node shares against the cs length (IndexError REF:227), dropped
//            ends.  Records with several ops get the prefix sums the merge walk needs (REF:205-255) from a warp scan and
//            are listed for `fold`.  Warp 1 first drains the previous tile's list of links that are not inline (hash probes); warp 0
//            takes what is left of it when its records are done.
//   fold     the listed steps of multi-op records, half of them per warp: clear_align / compact_align (REF:63-107) folded
//            over the op pieces that overlap the node: dropped or not, counting ops, deletion-derived IL/OL keys.
//   count    one thread per surviving step, no dependent loads from the tables: ONE 32-bit RED (tables.cuh), stamps only
//            while the node's settled bit is clear.  Links that are not inline are listed for the next tile's walk phase.
//            No barrier closes the tile: a warp that is done starts scanning the next one.
#pragma once

namespace teamp {

constexpr uint32_t NONE32 = 0xffffffffu;
constexpr uint32_t FULL = 0xffffffffu;
constexpr uint32_t THREADS = 64;
constexpr int MAX_STEPS = 250;                // longer paths take the exact path
constexpr int MAX_OPS = 48;                   // more cs ops: exact path
constexpr int32_t MAX_NTOT = 60000;           // longer cs strings: exact path (cs coordinates are kept in 16 bits, saturating)
constexpr uint32_t SL_BAD = 0xFFFFu;          // sL[] entry: unknown node / share that does not fit 16 bits

enum : uint8_t { ST_FAST = 0, ST_DONE = 1, ST_DEFER = 2 };      // per role; a record's status is the maximum
enum : uint32_t { OP_MATCH = 0, OP_SUB = 1, OP_DEL = 2, OP_INS = 3, OP_EQ = 4 };   // ':' '*' '-' '+' '='  (op = kind | len << 3)

struct __align__(4) Rec {
    int32_t start;        // int(tokens[7])                                    (role B)
    int32_t end_rel1;     // int(tokens[6]) - int(tokens[8]) - 1               (role B)
    uint16_t n_tot;       // sum of the cs op lengths (<= MAX_NTOT)            (role A)
    uint16_t start_add;   // cigar_clipping: start_pos += len of a leading '+' (role A, REF:46-47)
    uint32_t sum;         // sum of the node lengths of the record's steps     (role B zeroes, ids adds)
    uint16_t s0;          // first entry of the record in the step list        (role B)
    uint16_t nsteps;      //                                                   (role B)
    uint16_t ls;          // buffer position of the record's first byte        (role B)
    uint16_t op_off;      // first op of the record in the op pool             (role A)
    uint8_t nops;         //                                                   (role A)
    uint8_t stA, stB;     // ST_* per role; walk raises stB
    uint8_t whyA;         // WHY_* when role A says ST_DEFER
    uint16_t spare;
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

template <int TILE_, int OV_, int STEP_CAP_, int NT_, int CTAS_>
struct Geo {
    static constexpr int TILE = TILE_;
    static constexpr int OV = OV_;
    static constexpr int BUF = 16 + TILE + OV + 16;               // [pre 16][tile][look-ahead][pad 16]
    static constexpr int NV = ((16 + TILE + OV) / 16 + 14 + 3) & ~3;  // 16-byte vectors = 16-bit mask words (+ the walkers' end marks, padded)
    static constexpr int LINE_CAP = 64;                           // records per tile (slots: 6 bits)
    static constexpr int STEP_CAP = STEP_CAP_;                    // typical: 15 entries per 300-byte record
    static constexpr int OPS_CAP = 192;
    static constexpr int HEAVY_CAP = 160;                         // steps of multi-op records that an op boundary / mismatch / indel falls into
    static constexpr int FAR_CAP = 48;                            // links that are not inline: typically 1 per record
    static constexpr int DEL_CAP = 24;                            // steps with deletion-derived keys
    static constexpr int NT = NT_;                                // teams per CTA: they walk through the phases together (one instruction working set)
    static constexpr int CTAS = CTAS_;                            // CTAs per SM
    // masks are dead after `records`; sL / prefix pool / heavy list live from `ids` to `fold` in the same bytes
    static constexpr int MASK_BYTES = 4 * NV;
    static constexpr int WALK_BYTES = 2 * STEP_CAP + 2 * STEP_CAP + 2 * HEAVY_CAP;
    static constexpr int OFF_CTL = (BUF + 15) & ~15;              // TeamCtl
    static constexpr int OFF_WM = (OFF_CTL + 64 + 127) & ~127;
    static constexpr int OFF_SM = OFF_WM + 2 * NV;
    static constexpr int OFF_SL = OFF_WM;
    static constexpr int OFF_SR = OFF_SL + 2 * STEP_CAP;
    static constexpr int OFF_HEAVY = OFF_SR + 2 * STEP_CAP;
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
    static_assert(2 * STEP_CAP <= 2 * NV, "`ids` writes sL while it still reads the separator masks: sL must stay inside the whitespace masks");
    static_assert(LINE_CAP <= 64, "step entries hold 6-bit record slots");
    static_assert(STEP_CAP < 65536 && OPS_CAP < 65536 && HEAVY_CAP < 65536, "records hold 16-bit list offsets");
    static_assert((SMEM_BYTES * NT + 1024 + 64) * CTAS <= 227 * 1024, "CTAS x NT teams must fit one SM's shared memory");
};

// one LOP3 for any three-input bitwise function (LUT = f(0xF0, 0xCC, 0xAA)); written out because the compiler spends two
// instructions on an expression with two different constants
template <int LUT>
__device__ __forceinline__ uint32_t lop3(uint32_t a, uint32_t b, uint32_t c) {
#ifndef PT_EMU
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, %4;" : "=r"(d) : "r"(a), "r"(b), "r"(c), "n"(LUT));
    return d;
#else
    uint32_t d = 0;
    for (int k = 0; k < 8; k++)
        if ((LUT >> k) & 1) d |= ((k & 4) ? a : ~a) & ((k & 2) ? b : ~b) & ((k & 1) ? c : ~c);
    return d;
#endif
}

// no "s:" / "v:" byte pair inside: neither regex of REF:154-156,172-174 can start in this token.  Short tokens are checked
// byte by byte; a long one (bq:Z:<one quality character per base>, what vg writes for FASTQ reads) a word at a time for a
// sufficient condition: no byte in 0x60..0x7F after the "xx:T:" prefix, i.e. no lower-case letter at all (Phred+33 qualities
// end at '~' = Q93, real ones far below 's' = Q82).
__device__ __forceinline__ bool token_is_inert(const uint8_t* s, uint32_t a, uint32_t b) {
    if (b - a > 48u) {
        if (b - a > 4096u) return false;
        // the tag's own name is lower case ("bq:Z:"): its first bytes get the exact pair test, the rest the word test
        uint32_t prev = 0;
        for (uint32_t q = a; q < a + 8u; q++) {
            const uint32_t c = s[q];
            if (c == ':' && (prev == 's' || prev == 'v')) return false;
            prev = c;
        }
        a += 5u;
        const uint32_t* w = reinterpret_cast<const uint32_t*>(s);
        const uint32_t wa = a >> 2, wb = (b - 1u) >> 2;                       // first / last word holding a token byte
        uint32_t acc = 0;
        {
            const uint32_t x = w[wa] & (0xFFFFFFFFu << (8u * (a & 3u)));        // bytes before a: dropped
            acc |= x & (x << 1);
        }
        for (uint32_t i = wa + 1u; i < wb; i++) {
            const uint32_t x = w[i];
            acc |= x & (x << 1);
        }
        {
            uint32_t x = w[wb] & (0xFFFFFFFFu >> (8u * (3u - ((b - 1u) & 3u))));   // bytes from b on: dropped
            if (wb == wa) x &= 0xFFFFFFFFu << (8u * (a & 3u));
            acc |= x & (x << 1);
        }
        return (acc & 0x40404040u) == 0u;                                    // bit 6 and bit 5 of one byte both set
    }
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

// The walkers read the mask arrays as 32-bit words (32 bytes of the tile each).
// Next whitespace bit at or after the walker's position.  No bounds checks: the scan leaves four zero words and then two
// all-ones words behind the data, so a walker that runs off the record's end finds positions >= lim (checked once by the
// caller).  Up to two empty words (a read name, a long tag) are skipped without a loop: the lanes of a role warp stay together.
__device__ __forceinline__ void next_ws(const uint32_t* wm32, uint32_t& wi, uint32_t& m, uint32_t& pos) {
    if (m == 0u) m = wm32[++wi];
    if (m == 0u) m = wm32[++wi];
    while (m == 0u) m = wm32[++wi];
    pos = 32u * wi + (uint32_t)(__ffs((int)m) - 1);
    m &= m - 1u;
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

// per-team control block (shared memory)
struct __align__(8) TeamCtl {
    uint64_t mbar;               // the tile's TMA copy has landed
    uint32_t cnt[2];             // record starts found by each warp
    uint32_t nent, nheavy[2], nfar, ndel, far_take, lwm;
};
static_assert(sizeof(TeamCtl) <= 64, "OFF_WM leaves 64 bytes for the control block");

template <class G>
__global__ void __launch_bounds__(THREADS * G::NT, G::CTAS) augment_team_kernel(ChunkArgs A, Tables T) {
    PT_DYNAMIC_SMEM(smem_cta);
    // The CTA is G::NT independent teams that only share the barriers: all teams of an SM are then in (nearly) the same
    // phase, so the SM's instruction caches hold one or two phases instead of all of them.
    const uint32_t team = threadIdx.x / THREADS, tid = threadIdx.x % THREADS, lane = tid & 31u, warp = tid >> 5;
    uint8_t* const smem = smem_cta + team * (uint32_t)G::SMEM_BYTES;
    TeamCtl& C = *reinterpret_cast<TeamCtl*>(smem + G::OFF_CTL);
    uint64_t& mbar = C.mbar;
    uint32_t* const s_cnt = C.cnt;
    uint32_t* const s_nheavy2 = C.nheavy;
    uint32_t &s_nent = C.nent, &s_nfar = C.nfar, &s_ndel = C.ndel, &s_far_take = C.far_take, &s_lwm = C.lwm;
    const uint32_t gteam = blockIdx.x * (uint32_t)G::NT + team, nteams = gridDim.x * (uint32_t)G::NT;
    // inside a tile the phases only need the team's own two warps; the CTA-wide barrier after `scan` re-aligns the teams
    auto team_sync = [&]() { if (A.loose) named_barrier(team + 1u, THREADS); else __syncthreads(); };
    const uint32_t lt_mask = (1u << lane) - 1u;
    uint8_t* const buf = smem;
    uint16_t* const wm16 = reinterpret_cast<uint16_t*>(smem + G::OFF_WM);      // whitespace mask, 16 bits per 16-byte vector
    uint16_t* const sm16 = reinterpret_cast<uint16_t*>(smem + G::OFF_SM);      // separators + non-tab whitespace
    const uint32_t* const wm32 = reinterpret_cast<const uint32_t*>(smem + G::OFF_WM);   // the same masks, as half words
    const uint32_t* const sm32 = reinterpret_cast<const uint32_t*>(smem + G::OFF_SM);
    uint16_t* const sL = reinterpret_cast<uint16_t*>(smem + G::OFF_SL);        // share of the query per step (REF:215-218), SL_BAD
    uint16_t* const sR = reinterpret_cast<uint16_t*>(smem + G::OFF_SR);        // cs coordinate where the step's node starts (saturating)
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
        s_nheavy2[0] = 0;
        s_nheavy2[1] = 0;
        s_nent = 0;
    }
    __syncthreads();

    DevSink sink(T);
    const uint32_t ablate = A.ablate;            // diagnostics: 0 = everything, k = stop every tile after phase k (profiles/ ablation ladder)
    const uint64_t nbytes16 = (A.nbytes + 15ull) & ~15ull;
    uint32_t parity = 0;
    unsigned long long my_tiles = 0;
    uint32_t my_real = 0;                        // records this thread has listed
    int64_t far_base = 0;                        // file offset of buf[0] of the tile that filled the far-link list

    // bytes per tile: G::TILE is what the shared-memory layout holds, the host picks the length (a multiple of 16) so that a
    // tile carries just under 32 records -- one lane each in `records` and `walk`
    const uint32_t tile_b = A.tile_bytes;
    auto issue_load = [&](uint32_t tile) {
        const uint64_t t0 = (uint64_t)tile * tile_b;
        const uint64_t lo = tile ? t0 - 16 : 0;
        const uint64_t hi = min(t0 + tile_b + G::OV, nbytes16);
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

    if (gteam < A.n_tiles && tid == 0) issue_load(gteam);

    // every team of the CTA makes the same number of rounds (the barriers are CTA-wide); a team without a tile idles through
    const uint32_t rounds = (A.n_tiles + nteams - 1u) / nteams;
    for (uint32_t round = 0; round < rounds; round++) {
        const uint32_t tile = gteam + round * nteams;
        const bool active = tile < A.n_tiles;
        const uint64_t t0 = (uint64_t)tile * tile_b;
        const uint32_t owned = active ? (uint32_t)min((uint64_t)tile_b, A.nbytes - t0) : 0u;
        const uint64_t hi = min(t0 + tile_b + G::OV, nbytes16);
        const uint32_t lim = active ? 16u + (uint32_t)(min(hi, A.nbytes) - t0) : 0u;     // data ends here in the buffer
        const int64_t base_off = A.file_off + (int64_t)t0 - 16;            // file offset of buf[0]
        const uint32_t own_end = active ? 16u + owned : 0u;                 // records starting before this are ours
        const uint32_t nvec = (lim + 15u) >> 4;                             // vectors holding data
        const uint32_t nvec2 = (nvec + 1u) & ~1u;                           // whole 32-bit mask words
        const uint32_t nhalf = nvec2 >> 1;
        const uint32_t nxt_tile = tile + nteams;
        if (tid == 0 && active) {
            // low-water mark of the running kernel (tables.cuh): every tile below it is complete
            const unsigned long long lw = *(volatile unsigned long long*)&T.sc[SC_LWM];
            const int64_t rel = A.file_off - T.epoch_base + (int64_t)(lw * (unsigned long long)tile_b);
            s_lwm = rel <= 0 ? 0u : (rel > 0xFFFFFFF0ll ? 0xFFFFFFF0u : (uint32_t)rel);
        }
        if (active) {
            mbar_wait(&mbar, parity);
            parity ^= 1;
        }
        if (ablate == 1u) {                                                 // TMA only
            __syncthreads();
            if (tid == 0 && nxt_tile < A.n_tiles) issue_load(nxt_tile);
            continue;
        }

        // ================= scan: whitespace / separator masks, record starts =================
        // A warp iteration covers 2 KiB: lane L takes the 16-byte vectors L, L + 64, L + 128, L + 192 of the team's 4 KiB
        // (consecutive lanes, consecutive vectors: no bank conflicts) and stores one 16-bit mask pair per vector; the
        // walkers read the mask arrays as 32-bit words (32 bytes of the tile each).
        {
            auto scan_vec = [&](uint32_t v, uint32_t& hib, auto tail) {
                const uint4 q = *reinterpret_cast<const uint4*>(buf + 16u * v);
                const uint32_t x4[4] = {q.x, q.y, q.z, q.w};
                uint32_t bw[4], sv[4];
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const uint32_t x = x4[k];
                    // exact for ASCII bytes; a byte >= 0x80 is a fatal PT_U_NON_ASCII error anyway
                    bw[k] = x + 0x5F5F5F5Fu;                                      // bit 7 clear: x <= 0x20
                    const uint32_t ts = lop3<0x56>(x, 0x02020202u, 0x3E3E3E3Eu) + 0x7F7F7F7Fu;   // (x | 2) ^ '>': bit 7 clear: '>' or '<'
                    const uint32_t t = x + 0x76767676u;                           // bit 7 set: x >= 0x0A
                    sv[k] = lop3<0x2F>(ts, bw[k], t);                             // bit 7: separator, or whitespace that is not a tab
                }
                // two words per multiply: flags of word 2k+1 at bit 7, of word 2k at bit 3 of every byte -> the top byte of
                // the product holds the eight flags in byte order (partial products land on distinct bits: no carries)
                const uint32_t w01 = ~lop3<0xCA>(0x80808080u, bw[1], bw[0] >> 4) & 0x88888888u;
                const uint32_t w23 = ~lop3<0xCA>(0x80808080u, bw[3], bw[2] >> 4) & 0x88888888u;
                const uint32_t s01 = lop3<0xCA>(0x80808080u, sv[1], sv[0] >> 4) & 0x88888888u;
                const uint32_t s23 = lop3<0xCA>(0x80808080u, sv[3], sv[2] >> 4) & 0x88888888u;
                uint32_t wm = __byte_perm(w01 * 0x00204081u, w23 * 0x00204081u, 0x4473);   // byte 3 of each product -> 16-bit mask
                uint32_t sm = __byte_perm(s01 * 0x00204081u, s23 * 0x00204081u, 0x4473);
                const bool last = decltype(tail)::value && v + 2u >= nvec;   // the last vectors: nothing past the data
                if (last) {
                    const uint32_t room = lim > 16u * v ? lim - 16u * v : 0u;
                    const uint32_t keep = room < 16u ? ~(~0u << room) : 0xFFFFu;
                    wm &= keep;
                    sm &= keep;
                }
                wm16[v] = (uint16_t)wm;
                sm16[v] = (uint16_t)sm;
                if (last) {                                                 // (bytes past the data are whatever the buffer held before)
                    for (uint32_t p = max(16u * v, 16u); p < min(16u * v + 16u, min(lim, own_end)); p++)
                        if (buf[p] >= 0x80u) hib |= 0x80u;
                } else {
                    hib |= (q.x | q.y) | (q.z | q.w);
                }
            };
            const uint32_t nfull = nvec > 2u ? (nvec - 2u) / (4u * THREADS) : 0u;     // rounds in which every vector is whole
            uint32_t hib = 0;
            for (uint32_t r = 0; r < nfull; r++) {
#pragma unroll
                for (int u = 0; u < 4; u++) scan_vec(4u * THREADS * r + THREADS * u + tid, hib, std::false_type());
            }
            for (uint32_t v0 = 4u * THREADS * nfull; v0 < nvec2; v0 += 4u * THREADS) {
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const uint32_t v = v0 + THREADS * u + tid;
                    if (v < nvec2) scan_vec(v, hib, std::true_type());
                }
            }
            if ((hib & 0x80808080u) != 0u) {                                // non-ASCII byte: not modelled.  (Rare: look again.)
                for (uint32_t v = tid; v < nvec; v += THREADS)
                    for (uint32_t p = max(16u * v, 16u); p < min(16u * v + 16u, min(lim, own_end)); p++)
                        if (buf[p] >= 0x80u) { report_error(T, pt::PT_U_NON_ASCII, base_off + (int64_t)p); break; }
            }
        }
        __syncwarp();
        // ---- record starts.  A warp reads back the mask words it wrote itself (16 words of every 32): whitespace that is
        // not a tab (sep & ws) is a '\n' (record end), a '\r' or something odd; ranked by one warp scan, no atomics.
        uint32_t my_cnt = 0;                                                // warp-uniform: list slots this warp has handed out
        {
            const uint32_t p_min = tile ? 15u : 16u;                        // is the byte before the tile a newline? (tile 0: nothing there)
            const uint32_t w_end = (own_end + 31u) >> 5;                    // record starts are ours up to own_end
            const uint32_t w_first = 32u * (lane >> 4) + 16u * warp + (lane & 15u);
            uint32_t cnt = 0;
            for (uint32_t w = w_first; w < w_end; w += 64u) cnt += (uint32_t)__popc(wm32[w] & sm32[w]);
            uint32_t total;
            uint32_t j = warp_excl_scan(cnt, lane, total);
            if (active && tile == 0 && warp == 0 && owned > 0u) {           // the chunk starts at a record start
                if (lane == 0) lines[0] = 16;
                j += 1u;
                total += 1u;
                my_real += lane == 0 ? 1u : 0u;
            }
            my_cnt = total;
            for (uint32_t w = w_first; w < w_end && cnt != 0u; w += 64u) {
                uint32_t c = wm32[w] & sm32[w];
                while (c) {
                    const uint32_t p = 32u * w + (uint32_t)(__ffs((int)c) - 1);
                    c &= c - 1u;
                    cnt--;
                    const uint32_t ch = buf[p];
                    const bool is_start = ch == '\n' && p + 1u < own_end && p >= p_min;
                    if (ch == '\r' && p >= 16u && p < own_end) {            // lone '\r': a line break for the reference's text mode
                        const uint64_t abs_pos = t0 + p - 16u;
                        if (abs_pos + 1 < A.nbytes && buf[p + 1] != '\n') report_error(T, pt::PT_U_BARE_CR, base_off + (int64_t)p);
                    }
                    // (a candidate that is no record start keeps its slot: 0 = no record)
                    if (j < (uint32_t)G::LINE_CAP) lines[warp ? (uint32_t)G::LINE_CAP - 1u - j : j] = is_start ? (uint16_t)(p + 1u) : (uint16_t)0;
                    j++;
                    my_real += is_start ? 1u : 0u;
                }
            }
        }
        if (tid < 12u) wm16[nvec2 + tid] = tid < 8u ? 0u : 0xFFFFu;         // end marks for the walkers (next_ws)
        if (lane == 0) s_cnt[warp] = my_cnt;
        __syncthreads();                                                    // ---- B1: masks + record lists complete; everyone is through the previous tile
        const uint32_t n0 = s_cnt[0], n1 = s_cnt[1];
        const uint32_t lwm_rel = s_lwm;
        if (tid == 0 && active) {
            my_tiles++;
            *(volatile uint32_t*)&T.team_tile[gteam] = tile;                // my tiles before this one are complete
        }
        if (gteam == 0 && warp == 1) {
            // one team keeps the kernel's low-water mark: the smallest tile any team is still working on
            uint32_t m = 0xFFFFFFFFu;
            for (uint32_t i = lane; i < nteams; i += 32u) m = min(m, *(volatile uint32_t*)&T.team_tile[i]);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) m = min(m, __shfl_xor_sync(FULL, m, o));
            if (lane == 0 && m != 0xFFFFFFFFu) atomicMax(&T.sc[SC_LWM], (unsigned long long)m);
        }
        if (ablate == 2u) {
            __syncthreads();
            if (tid == 0 && nxt_tile < A.n_tiles) issue_load(nxt_tile);
            continue;
        }

        uint32_t n_lines = n0 + n1;
        if (n_lines > (uint32_t)G::LINE_CAP) {
            // more records than the list holds (pathological input): all of them take the exact path
            for (uint32_t p = 15u + tid; p + 1u < own_end; p += THREADS) {
                const bool nl = p == 15u ? (tile == 0 || buf[p] == '\n') : buf[p] == '\n';
                if (nl) defer_line(T, t0 + p + 1u - 16u, A.file_off, WHY_LINES_FULL);
            }
            n_lines = 0;
        }

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
                }
                if (have && ls != 0u) {
                    uint32_t wi = ls >> 5;
                    uint32_t wmk = wm32[wi] & (~0u << (ls & 31u));
                    uint32_t e[13];
                    e[0] = ls - 1u;
                    uint32_t gaps_ok = 1u, tabs = 1u;         // AND over the boundaries: no empty column / a tab (the first 11)
#pragma unroll
                    for (int j = 1; j <= 12; j++) {
                        next_ws(wm32, wi, wmk, e[j]);
                        gaps_ok &= e[j] - e[j - 1] >= 2u ? 1u : 0u;
                        if (j < 12) tabs &= buf[e[j]] == '\t' ? 1u : 0u;
                    }
                    bool slow = e[12] >= lim, done = false, no_tags = false;      // record runs past the look-ahead
                    int32_t mapq = 0;
                    if (!slow) {
                        // 11 single tabs, then a tab (tags follow) or the end of a 12-column record
                        const uint32_t c12 = buf[e[12]];
                        no_tags = c12 == '\n';
                        if ((gaps_ok & tabs) == 0u || (c12 != '\t' && !no_tags)) { slow = true; why = WHY_COLUMNS; }
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
                    R.sum = 0;
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
            if (lane == 0) s_nent = min(step_base, (uint32_t)G::STEP_CAP);
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
                const uint32_t ls = have ? lines[l < n0 ? l : (uint32_t)G::LINE_CAP - 1u - (l - n0)] : 0u;
                if (ls != 0u) {                                             // (0: a candidate that was no record start)
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
                        if (++wi >= nhalf) { ran_off = true; break; }
                        wmk = wm32[wi];
                    }
                    for (; skip != 0u && !ran_off; skip--) wmk &= wmk - 1u;
                    if (!ran_off) {
                        next_ws(wm32, wi, wmk, e11);
                        next_ws(wm32, wi, wmk, e12);
                        ran_off = e12 >= lim;
                    }
                    // role B decides about everything up to column 12; here: is there anything left to do?
                    int32_t mapq = 0;
                    idle = ran_off || buf[e12] != '\t' || !small_uint(buf, e11 + 1u, e12, mapq) || (int64_t)mapq < A.thr;
                    uint32_t dv_a = 0, dv_b = 0, dv3 = 0;
                    unsigned long long cs8 = 0;
                    if (!idle) {
                        // ---- tags: [inert]* cs [inert]* dv in any order, within the first few tags
                        why = WHY_TAGS;
                        uint32_t a = e12 + 1u, b = 0;
                        next_ws(wm32, wi, wmk, b);
                        if (b >= lim) slow = true;
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
                            next_ws(wm32, wi, wmk, b);
                            if (b >= lim) { slow = true; break; }
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
                            R.n_tot = (uint16_t)n_tot;
                            R.op_off = (uint16_t)op_off;
                            R.start_add = (uint16_t)start_add;
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
        team_sync();                                                          // ---- B2: records, ops, step list complete
        const uint32_t n_ent = s_nent;                                      // step entries incl. sentinels
        if (ablate == 3u) {
            __syncthreads();
            if (tid == 0 && nxt_tile < A.n_tiles) issue_load(nxt_tile);
            continue;
        }

        // ================= ids: one thread per path step: id -> node index -> the node's hot record =================
        // UI steps per thread and iteration: their 16-byte loads are all in flight before the first is used.  A warp holds 32
        // consecutive entries, so the collapsible duplicates of REF:188 show up in a shuffle; the record's sum of node lengths
        // is collected on the way.
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
                    idx_[u] = idx;                                          // NONE32: KeyError in the reference, the exact path reports it
                    se_[u] = se;
                    hot_[u] = make_uint4(0u, 0u, 0u, 0u);
                    if (idx != NONE32) hot_[u] = sink.load_hot(idx);        // issued at once: in flight while the next id is parsed
                }
#pragma unroll
                for (int u = 0; u < UI; u++) {
                    const uint32_t s = s00 + THREADS * u + tid;
                    const uint32_t se = se_[u];
                    const uint32_t prev = __shfl_up_sync(FULL, idx_[u], 1);  // (lane 0's predecessor sits in the other warp: `walk` checks those pairs)
                    if (s < n_ent) {
                        const uint32_t meta = hot_[u].x, len = meta & META_LEN_MASK;
                        const bool step = se != SE_INVALID && !(se & SE_SENT);
                        const bool ok = idx_[u] != NONE32 && len - 1u < META_LEN_ESC - 1u;   // absent node, >= 1023 bases, unknown id: exact path
                        sidx[s] = idx_[u];
                        smeta[s] = meta;
                        sL[s] = (uint16_t)(ok ? len : SL_BAD);
                        if (step) {
                            Rec& R = recs[(se >> SE_SLOT_SHIFT) & SE_SLOT_MASK];
                            if (!ok || (lane != 0u && !(se & SE_FIRST) && idx_[u] == prev)) {
                                R.stB = ST_DEFER;                           // (a FAST record: role B listed the step)
                                R.whyB = WHY_WALK;
                            } else {
                                atomicAdd(&R.sum, len);
                            }
                        }
                    }
                }
            }
        }
        team_sync();                                                          // ---- B3: node indices complete; the bytes and the masks are dead
        if (tid == 0 && nxt_tile < A.n_tiles) issue_load(nxt_tile);        // overlaps walk + fold + count
        if (ablate == 4u) continue;

        // ================= walk: warp 0 the even records, warp 1 the previous tile's far links and the odd records =================
        // (a) One thread per record, no loop: the two ends of the path are shortened (REF:215-218); every other node keeps its
        //     whole length, so the record's sum of shares follows from the sum `ids` collected, and with it the one check the
        //     merge walk needs for a single-op record: a node with bases left but no cs left is an IndexError (REF:227).
        // (b) Records whose cs string has several ops, one at a time, one lane per step: prefix sum of the shares = the cs
        //     coordinate of every node (REF:205-255); a node inside one ':' / '=' op has one counting op like every node of a
        //     single-op record, the nodes an op boundary or a mismatch / indel falls into are listed for `fold`.
        if (warp == 1) {
            drain_far();
            if (lane == 0) s_ndel = 0;
        }
        {
            constexpr uint32_t HALF = (uint32_t)G::HEAVY_CAP / 2u;          // every warp lists into its own half
            uint32_t n_heavy_w = 0;                                         // warp-uniform: entries of this warp's fold list
            for (uint32_t l0 = 0; 2u * l0 < n_lines; l0 += 32u) {
                const uint32_t l = 2u * (l0 + lane) + warp;
                bool multi = false;
                if (l < n_lines) {
                    Rec& R = recs[l];
                    if (rec_status(R) == ST_FAST) {
                        const uint32_t ns = R.nsteps, s0 = R.s0, sl = s0 + ns - 1u, n_tot = R.n_tot;
                        const uint32_t raw0 = sL[s0], raw1 = sL[sl];        // (not SL_BAD: `ids` would have handed the record over)
                        // first node: L -= start_pos (REF:215-216); last node: L = L - end_pos_rel + 1 (REF:217-218)
                        int64_t L0 = (int64_t)raw0 - ((int64_t)R.start + R.start_add), L1 = (int64_t)raw1 - R.end_rel1;
                        if (ns == 1u) { L0 -= R.end_rel1; L1 = L0; }
                        const uint32_t v0 = (uint32_t)(L0 <= 0 ? 0 : (L0 >= (int64_t)SL_BAD ? (int64_t)SL_BAD : L0));
                        const uint32_t v1 = (uint32_t)(L1 <= 0 ? 0 : (L1 >= (int64_t)SL_BAD ? (int64_t)SL_BAD : L1));
                        // pairs of steps that `ids` could not compare by shuffle (the later one sits in lane 0 of its warp)
                        bool bad = v0 == SL_BAD || v1 == SL_BAD;
                        for (uint32_t q = (s0 + 32u) & ~31u; q <= sl; q += 32u) bad |= sidx[q] == sidx[q - 1u];
                        // shares: v0, the interior nodes' whole lengths, v1; the last node with bases left starts at a_last
                        const uint32_t total = ns == 1u ? v0 : R.sum - raw0 - raw1 + v0 + v1;
                        uint32_t a_last = 0;
                        if (ns > 1u) a_last = v1 ? total - v1 : (ns > 2u ? total - sL[sl - 1u] : 0u);
                        if (total != 0u && a_last >= n_tot) bad = true;    // IndexError (REF:227): the exact path reports it
                        if (bad) {
                            R.stB = ST_DEFER;
                            R.whyB = WHY_WALK;
                        } else {
                            sL[s0] = (uint16_t)v0;
                            sL[sl] = (uint16_t)v1;
                            // a node is not in `align` iff no bases are left for it: only the two ends can run out
                            if (v0 == 0u) steps[s0] |= SE_DROPPED;
                            if (v1 == 0u) steps[sl] |= SE_DROPPED;
                            multi = R.single == 0;
                        }
                    }
                }
                // ---- (b) the multi-op records of this batch, one after the other, the warp's lanes on the record's steps
                uint32_t todo = __ballot_sync(FULL, multi);
                while (todo) {
                    const uint32_t lr = 2u * (l0 + (uint32_t)(__ffs((int)todo) - 1)) + warp;
                    todo &= todo - 1u;
                    Rec& R = recs[lr];
                    const uint32_t ns = R.nsteps, s0 = R.s0, n_tot = R.n_tot;
                    const uint32_t* op = ops + R.op_off;
                    uint32_t carry = 0;
                    bool full = false;
                    for (uint32_t k0 = 0; k0 < ns; k0 += 32u) {
                        const uint32_t k = k0 + lane;
                        const uint32_t v = k < ns ? sL[s0 + k] : 0u;
                        uint32_t incl = v;
#pragma unroll
                        for (int o = 1; o < 32; o <<= 1) {
                            const uint32_t y = __shfl_up_sync(FULL, incl, o);
                            if (lane >= (uint32_t)o) incl += y;
                        }
                        const uint32_t run = carry + incl - v;              // cs coordinate where this node starts
                        carry += __shfl_sync(FULL, incl, 31);
                        bool list = false;
                        if (k < ns && v != 0u && run < n_tot) {             // (run >= n_tot with bases left: excluded in (a))
                            uint32_t j = 0, o_end = op[0] >> 3;
                            while (o_end <= run) o_end += op[++j] >> 3;     // (run < n_tot = sum of the op lengths)
                            const uint32_t kind = op[j] & 7u;
                            list = min(run + v, n_tot) > o_end || (kind != OP_MATCH && kind != OP_EQ);
                            sR[s0 + k] = (uint16_t)run;
                        }
                        const uint32_t bal = __ballot_sync(FULL, list);
                        if (list) {
                            const uint32_t h = n_heavy_w + (uint32_t)__popc(bal & lt_mask);
                            if (h < HALF) heavy[warp * HALF + h] = (uint16_t)(s0 + k);
                        }
                        n_heavy_w += (uint32_t)__popc(bal);
                        full |= n_heavy_w > HALF;
                    }
                    if (full && lane == 0) {                                // list full: exact path (nothing counted yet)
                        R.stB = ST_DEFER;
                        R.whyB = WHY_WALK;
                    }
                }
            }
            if (lane == 0) s_nheavy2[warp] = min(n_heavy_w, HALF);
        }
        drain_far();                                                        // (whatever warp 1 has not taken yet: warp 0 is usually here first)
        team_sync();                                                          // ---- B4: hand-over decisions of `walk`, prefix pool, heavy list; the far-link list is drained
        if (active) far_base = base_off;                                    // the list `count` fills below belongs to this tile
        if (tid == 0) {                                                     // (`count` appends after B5)
            s_nfar = 0;
            s_far_take = 0;
        }
        if (ablate == 5u) continue;

        // ================= fold: every step of a multi-op record folds the cs ops that overlap its node =================
        const uint32_t nh0 = s_nheavy2[0], n_heavy = nh0 + s_nheavy2[1];
        {
            for (uint32_t h = tid; h < n_heavy; h += THREADS) {
                const uint32_t s = heavy[h < nh0 ? h : (uint32_t)G::HEAVY_CAP / 2u + (h - nh0)];
                const uint32_t se = steps[s];
                Rec& R = recs[(se >> SE_SLOT_SHIFT) & SE_SLOT_MASK];
                if (rec_status(R) != ST_FAST) continue;
                const uint32_t Lk = sL[s];                                  // > 0 (walk)
                const uint32_t Ak = sR[s], n_tot = R.n_tot;                 // Ak < n_tot (walk)
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
            team_sync();                                                      // ---- B5: every hand-over decision is made; nothing counted so far
        }
        for (uint32_t l = tid; l < n_lines; l += THREADS) {
            const Rec& R = recs[l];
            if (rec_status(R) == ST_DEFER) defer_line(T, t0 + R.ls - 16u, A.file_off, R.stB == ST_DEFER ? R.whyB : R.whyA);
        }

        // surviving neighbours of step s inside its record (dropped nodes are skipped)
        auto prev_survivor = [&](uint32_t s, uint32_t se) -> uint32_t {     // NONE32: s is the first survivor (se = steps[s])
            uint32_t t = s, e = se;
            while (!(e & SE_FIRST)) {
                e = steps[--t];
                if (!(e & SE_DROPPED)) return t;
            }
            return NONE32;
        };
        auto next_survivor = [&](uint32_t s, uint32_t se) -> uint32_t {     // NONE32: s is the last survivor
            uint32_t t = s, e = se;
            while (!(e & SE_LAST)) {
                e = steps[++t];
                if (!(e & SE_DROPPED)) return t;
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
            const uint32_t ps = prev_survivor(s, se), nx = next_survivor(s, se);
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
                const bool not_first = prev_survivor(s, se) != NONE32, not_last = next_survivor(s, se) != NONE32;
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
    if (tid == 0) *(volatile uint32_t*)&T.team_tile[gteam] = 0xFFFFFFFFu;   // nothing of mine is pending any more

    // rejected-record count: warp reduce, one RED per warp
    uint32_t r = sink.rej;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(FULL, r, o);
    if (lane == 0u && r) atomicAdd(&T.sc[SC_REJ], (unsigned long long)r);
    uint32_t nl = my_real;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) nl += __shfl_xor_sync(FULL, nl, o);
    if (lane == 0u && nl) atomicAdd(&T.sc[SC_LINES], (unsigned long long)nl);
    if (tid == 0) {
        if (my_tiles) atomicAdd(&T.sc[SC_TILES], my_tiles);
    }
}

}  // namespace teamp
