/*
 * pantas_aug.h -- C ABI of libpantas_aug.so, the B200 (sm_100a) implementation
 * of pantas' `augment` hot loop.
 *
 * The reference has no FFI for this path: its boundary is a process contract
 * (pantas:132 runs `python3 scripts/alignments_augmentation_from_gaf.py GAF GFA`).
 * This ABI is what a ctypes binding inside that script binds instead of running
 * the per-line Python loop; each entry point cites the reference code it
 * replaces (REF:n = scripts/alignments_augmentation_from_gaf.py line n).
 * INTEGRATION.md shows the binding.
 *
 * Conventions
 *   - plain C types only; no C++ exceptions cross the boundary;
 *   - every function returns 0 on success or a negative PT_ERR_* code;
 *     pt_last_error() gives the text of the last failure on that context;
 *   - DATA errors (a record on which the reference would raise, or input the
 *     device parser refuses to guess at) are sticky on the device and are
 *     reported by pt_error(); counters are meaningless after one;
 *   - a context belongs to one GPU and one host thread at a time;
 *   - "dev" pointers are device pointers on the context's GPU, 16-byte aligned.
 *
 * Result layout (pt_export_dense / pt_export_side), N = n_nodes, E = n_edges:
 *   sums   int64[3N + E + 4] = [ NC(N) | IL0adj(N) | OLadj(N) | RC(E) | rej, n_lines, 0, 0 ]
 *          IL[v][0]      = NC[v] + IL0adj[v]        (REF:298-305, 335-342)
 *          OL[v][len(v)] = NC[v] + OLadj[v]         (REF:306-313, 344-351)
 *   stamps int64[2N]         = [ first-touch stamp of IL[v][0] | of OL[v][len(v)] ]
 *          (INT64_MAX = never touched).  A stamp is (file byte offset of the
 *          path step << 2) | e and reproduces Python dict insertion order.
 *   novel  uint64[3 * n_novel]  rows {key = from_idx << 32 | to_idx, count, stamp}
 *          links absent from the GFA (REF:426-427), unordered
 *   sparse uint64[3 * n_sparse] rows {key = idx << 32 | dir << 31 | (pos + 2^30), count, stamp}
 *          deletion-derived IL (dir 0) / OL (dir 1) keys (REF:281-297, 317-333), unordered
 * sums add and stamps min across shards/ranks; side rows merge by key.
 */
#ifndef PANTAS_AUG_H
#define PANTAS_AUG_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pt_ctx pt_ctx;

#define PT_ABI_VERSION 1

/* API status codes (negative).  Data error codes (positive) are in pt_error(). */
#define PT_ERR_CUDA (-1)      /* a CUDA runtime call failed */
#define PT_ERR_ARG (-2)       /* bad argument */
#define PT_ERR_STATE (-3)     /* call out of order (e.g. no graph set) */
#define PT_ERR_NOMEM (-4)     /* allocation failed */
#define PT_ERR_NODEVICE (-5)  /* no CUDA device / wrong architecture */

/* Data error codes reported by pt_error() (see line_core.cuh):
 *   1..19  the reference raises on this record (IndexError/ValueError/KeyError/assert)
 *   20..39 input relies on Python behaviour the device parser does not model
 *   40..   internal capacity (novel-edge / sparse table full) */

int pt_abi_version(void);
const char* pt_strerror(int code);      /* API or data code -> static text */

/* Create / destroy a context on CUDA device `device`.  Fails with
 * PT_ERR_NODEVICE unless the device is compute capability 10.x. */
int pt_create(int device, pt_ctx** out);
void pt_destroy(pt_ctx* ctx);
const char* pt_last_error(pt_ctx* ctx);

/* Run all work of this context on an existing CUDA stream (a cudaStream_t,
 * e.g. torch.cuda.current_stream().cuda_stream).  Default: a private stream. */
int pt_set_stream(pt_ctx* ctx, void* cuda_stream);

/* Replaces REF:121-126 (nodes_info) and the implicit key set of REF:421.
 * node_len[i] is the sequence length of node id (min_id + i), 0xFFFFFFFF where
 * no such S line exists.  edge_keys[e] = from_idx << 32 | to_idx for the e-th
 * DISTINCT (from, to) pair of the GFA's L lines, in first-occurrence order.
 * Pointers may be host or device memory; the library copies.
 * novel_cap / sparse_cap: slots of the side tables (0 = default). */
int pt_set_graph(pt_ctx* ctx, const uint32_t* node_len, uint64_t n_nodes, uint32_t min_id,
                 const uint64_t* edge_keys, uint64_t n_edges, uint64_t novel_cap, uint64_t sparse_cap);

/* Zero all counters, stamps, side tables and the sticky error (graph stays). */
int pt_reset_counts(pt_ctx* ctx);

/* Replaces the loop REF:138-371 for the records in gaf_dev[0, nbytes).
 * The chunk must start at a line start and end at a line end or at EOF;
 * gaf_dev must be 16-byte aligned and readable up to nbytes rounded up to 16.
 * file_offset = byte offset of gaf_dev[0] in the whole GAF (stamps use it, so
 * results do not depend on how the file was chunked or sharded).
 * mapq_thr = argv[2] of the reference (default 20, REF:113).
 * Asynchronous: returns after enqueueing on the context's stream. */
int pt_process_chunk(pt_ctx* ctx, const uint8_t* gaf_dev, uint64_t nbytes, uint64_t file_offset,
                     int64_t mapq_thr);

/* Same, from HOST memory (pinned for full speed): copies through two
 * library-owned device staging buffers so the copy of one chunk overlaps the
 * kernels of the previous one.  nbytes <= pt_stage_bytes().  Returns a ticket
 * >= 0; the host buffer may be reused once pt_wait_copy(ticket) returns.  (The first call looks at the head of the buffer
 * for the average record length: the kernel's tiles are sized to carry about 30 records.) */
int64_t pt_process_host(pt_ctx* ctx, const uint8_t* gaf_host, uint64_t nbytes, uint64_t file_offset,
                        int64_t mapq_thr);
int pt_wait_copy(pt_ctx* ctx, int64_t ticket);
int pt_set_stage_bytes(pt_ctx* ctx, uint64_t bytes);   /* before first pt_process_host */
uint64_t pt_stage_bytes(pt_ctx* ctx);

int pt_sync(pt_ctx* ctx);

/* Sticky data error: *code = 0 if none, else the code and the file byte offset
 * of the first (lowest offset) offending record.  Synchronises. */
int pt_error(pt_ctx* ctx, uint64_t* bad_offset, int* code);

/* After the last chunk: compacts the side tables.  Synchronises. */
int pt_finalize(pt_ctx* ctx, uint64_t* n_novel, uint64_t* n_sparse);

/* Write results into caller-owned DEVICE buffers (layout above).  The caller
 * (Python, torch tensors) then reduces across ranks: REF has no counterpart,
 * see DESIGN.md "multi-GPU". */
int pt_export_dense(pt_ctx* ctx, int64_t* sums_dev, uint64_t sums_len, int64_t* stamps_dev, uint64_t stamps_len);
int pt_export_side(pt_ctx* ctx, uint64_t* novel_dev, uint64_t novel_rows, uint64_t* sparse_dev, uint64_t sparse_rows);

/* CUDA-event timer on the context's stream (bench.py). */
int pt_timer_start(pt_ctx* ctx);
int pt_timer_stop(pt_ctx* ctx, float* ms);     /* synchronises on the stop event */

/* Timing of the augment pass alone (augment_team_kernel + the exact per-record
 * kernel for the records it hands over): when enabled, every chunk
 * records a CUDA event pair around those launches on the context's stream.  pt_kernel_time() synchronises and returns the summed
 * duration and the number of launches since the last call (then clears them). */
int pt_profile_enable(pt_ctx* ctx, int on);
int pt_kernel_time(pt_ctx* ctx, float* ms_total, uint64_t* launches);
/* Same, split into the fast-path kernel and the exact per-record kernel (+ the chunk epilogue). */
int pt_kernel_time_split(pt_ctx* ctx, float* ms_fast, float* ms_slow, uint64_t* launches);

/* ---- The two GFA passes on the device.
 * Replaces REF:121-126 (pass 1: nodes_info) and REF:377-424 (pass 2: the writer); the caller (pantas_b200/gfa_device.py)
 * supplies the line index and does the prefix sums / key de-duplication with device-side index ops.
 *
 * pt_gfa_parse: one GFA line per thread.  start[i] / end[i] = byte range of line i's text (without its line break) in
 * gfa_dev.  Outputs per line: a_rel / slen = the stripped text (str.strip()), kind = flags (1: raw line starts with 'S',
 * 2: stripped line starts with 'S', 4: stripped line starts with 'L', bits 4..6: tokens of str.split(), capped at 4),
 * v1 = token 1 as a canonical decimal id (0xFFFFFFFF: any other spelling), v2 = len(token 2) for S lines, token 3 as an id
 * for L lines.  *err_dev (initialise to ~0) receives min(line << 8 | code): 1 S line with < 3 fields (reference:
 * IndexError), 2 S id spelling not supported, 3 sequence of >= 2^30 bases. */
int pt_gfa_parse(pt_ctx* ctx, const uint8_t* gfa_dev, const int64_t* start_dev, const int64_t* end_dev, uint64_t n_lines,
                 uint32_t* a_rel_dev, uint32_t* slen_dev, uint32_t* kind_dev, uint32_t* v1_dev, uint32_t* v2_dev, uint64_t* err_dev);

/* Inputs of the writer, all device pointers.  link_edge[i] = index (into the RC block of `sums`) of the count L line i
 * prints, -1 if it prints 0 (REF:421: weights.pop -- only the first L line of a key gets the count).  sums = what
 * pt_export_dense wrote (after the cross-rank reduction, if any).  Nodes with deletion-derived IL/OL keys print
 * sp_text[sp_off[k] .. sp_off[k+1]) with k = sp_slot[idx] >= 0 (the caller orders those few by first-touch stamp). */
typedef struct pt_gfa_writer {
    const uint8_t* gfa;
    const int64_t* start;
    const uint32_t* a_rel;
    const uint32_t* slen;
    const uint32_t* kind;
    const uint32_t* v1;
    const int32_t* link_edge;
    const uint32_t* node_len;
    const int64_t* sums;
    const int32_t* sp_slot;
    const int64_t* sp_off;
    const uint8_t* sp_text;
    uint64_t n_nodes;
    uint64_t n_lines;
    uint32_t min_id;
} pt_gfa_writer;

/* out_len[i] = bytes line i prints (REF:377-424; 0 for dropped lines).  Errors in *err_dev: 4 stripped S line without an
 * id (IndexError), 5 S id not in the node table (KeyError), 6 L line with < 4 fields (IndexError). */
int pt_gfa_measure(pt_ctx* ctx, const pt_gfa_writer* w, int64_t* out_len_dev, uint64_t* err_dev);
/* out_off = exclusive prefix sum of out_len; writes every line with its NC / IL / OL / RC tags at out_dev + out_off[i]. */
int pt_gfa_format(pt_ctx* ctx, const pt_gfa_writer* w, const int64_t* out_off_dev, uint8_t* out_dev, uint64_t* err_dev);

/* Counters for reports: kernel launches since create, records that took the
 * long-line path, tiles processed. */
int pt_stats(pt_ctx* ctx, uint64_t* kernel_launches, uint64_t* deferred_lines, uint64_t* tiles);

/* Why records were handed from the fast path to the exact per-record path since the last reset
 * (diagnostics): out[0..n) = long record / look-ahead, columns not 12 single tabs, integer syntax,
 * tag order or content, cs class, path column, step list full, walk (duplicate / unknown id, node
 * without bases, cs too short), record list full; the rest of out[0..n) is zero.  n <= 32.  Synchronises. */
int pt_debug_counters(pt_ctx* ctx, uint64_t* out, int n);

#ifdef __cplusplus
}
#endif
#endif /* PANTAS_AUG_H */
