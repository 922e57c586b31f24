/* dhr_b200.h -- C ABI of the B200-native GIP retrieval hot path.
 *
 * Drop-in boundary for castorini/dhr `retrieval/gip_retrieval.py` (reference @ e236f3d).
 * The reference has no FFI of its own (pure Python); these entry points are what a
 * ctypes / cffi binding of that file's hot functions binds to.  Each entry cites the
 * reference code it replaces.  Plain pointers and sizes only, no torch types; every
 * data pointer may be HOST or DEVICE memory (detected with cudaPointerGetAttributes).
 * All functions return DHR_OK (0) or a DHR_ERR_* code; dhr_strerror() maps codes to text.
 *
 * Scoring (gip_retrieval.py:110-125), S slices x G values per slice, C dense columns:
 *   score[p] = sum_{s<S} [q_idx[s]==p_idx[s]] * sum_{g<G} q_val[s*G+g]*p_val[s*G+g]
 *            + sum_{c<C} q_val[S*G+c]*p_val[S*G+c]
 * The reference is the G==1 case.  Top-k order: score descending, row ascending on ties.
 */
#ifndef DHR_B200_H
#define DHR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DHR_B200_VERSION 100

/* status codes */
#define DHR_OK               0
#define DHR_ERR_INVALID      1   /* bad argument (NULL, negative size, unknown dtype ...)            */
#define DHR_ERR_CUDA         2   /* a CUDA runtime call failed; dhr_last_cuda_error() has the text */
#define DHR_ERR_NOMEM        3   /* device or host allocation failed                                */
#define DHR_ERR_UNSUPPORTED  4   /* shape outside the supported envelope (group > 8, k > DHR_MAX_K) */
#define DHR_ERR_LOSSY        5   /* fp32 corpus value not representable in fp16 (index is fp16)     */
#define DHR_ERR_IDX_RANGE    6   /* corpus slice index negative or above the code range             */
#define DHR_ERR_STATE        7   /* call order violated (append after finalize, search before ...)  */
#define DHR_ERR_NO_DEVICE    8   /* no usable CUDA device                                           */

/* slice-index dtypes accepted on either side (encode.py:157,166 uint8; densify_corpus.py:31-34
 * int16 / int8; densify_query.py:73 int16; north_star uint16).  Equality is on the integer value. */
#define DHR_IDX_NONE 0
#define DHR_IDX_U8   1
#define DHR_IDX_I8   2
#define DHR_IDX_I16  3
#define DHR_IDX_U16  4
#define DHR_IDX_I32  5
#define DHR_IDX_I64  6

/* value dtypes: the on-disk index is fp16 (encode.py:156,165); the reference CPU path hands
 * fp32 copies of the same numbers to GIP_retrieval (gip_retrieval.py:275,313). */
#define DHR_VAL_F16 0
#define DHR_VAL_F32 1

/* dhr_search flags */
#define DHR_SEARCH_UNMASKED 1u   /* --IP first stage (gip_retrieval.py:139): plain inner product over all columns */

/* dhr_index_create flags */
#define DHR_INDEX_NARROW_CODES 1u /* store 16-bit slice indices as 8-bit codes (all values must be < 254) */
#define DHR_INDEX_LEX_POSTINGS 4u  /* experimental: keep the lexical part as per-(tile, slice) postings sorted by code (kernel K1p)
                                      instead of the tiled layout of the passage-per-thread kernel K1t; 8-bit codes only.
                                      Parity-tested, measured slower than K1t, hence opt-in */
#define DHR_INDEX_KEEP_ROWMAJOR 2u /* keep the row-major arrays resident after finalize (default: only the tiled copies stay;
                                      the row-major ones are rebuilt on the first call that needs them) */

#define DHR_MAX_K     12288      /* largest k handled by the fused selection (candidate capacity 16384) */
#define DHR_MAX_GROUP 8

typedef struct dhr_index dhr_index;

/* timing / accounting of the last dhr_search on an index (filled when profiling is enabled) */
typedef struct dhr_stats {
    int32_t  n_queries;
    int32_t  query_block;        /* QB: queries scored per corpus row visit by one scan CTA          */
    int32_t  query_groups;       /* groups of QB queries sharing one launch                          */
    int32_t  scan_variant;       /* 0 = direct 128-bit loads, 1 = TMA bulk (cp.async.bulk) staging   */
    int32_t  n_scan_launches;
    int32_t  n_select_launches;
    int32_t  n_prep_launches;
    int32_t  n_fallback_queries; /* queries re-run with the overflow-proof chunk schedule            */
    int32_t  n_kernel_launches;  /* every kernel launched by the call (prep, init, scan, select, ...)   */
    int32_t  rowmajor_rebuilds;  /* times the row-major arrays were rebuilt from the tiled copies by the call */
    double   scan_ms;            /* sum of CUDA-event durations of the scan launches                 */
    double   select_ms;          /* ... of the top-k selection launches                              */
    double   total_ms;           /* first launch to last launch of the search, CUDA events           */
    double   corpus_passes;      /* logical corpus passes made by the scan launches (sum rows*groups / N) */
    double   bytes_per_pass;     /* N * row_bytes of the HBM-resident layout                         */
    double   dense_flops;        /* 2 * Q * N * C issued to the tensor-core kernel K2 by the call        */
    double   lex_layout;         /* tile path: 0 = tiled lexical copy (K1t), 1 = postings (K1p)                           */
    double   alg_bytes;          /* algorithmic bytes of the scan launches: K1 rows*row_bytes per group of QB queries, K1t
                                    rows*(lexical value + code bytes) per tile of 64 queries, K2 rows*C_pad*2 per 128 queries */
} dhr_stats;

int         dhr_version(void);
const char* dhr_strerror(int status);
const char* dhr_last_cuda_error(void);
int         dhr_device_count(int* count);

/* ---- index lifetime -----------------------------------------------------------------------
 * Replaces the load + H2D of gip_retrieval.py:289-315 (pickle -> slice shard -> .cuda()).
 * The index keeps its own HBM-resident copy: fp16 lexical values, slice-index codes and the
 * fp16 dense block in three row-major arrays (see DESIGN.md "HBM layout").
 * row_offset is the global id of local row 0: a range shard built by the rule of
 * gip_retrieval.py:292-306 passes its first row here so results carry global rows. */
int dhr_index_create(dhr_index** out, int device, int64_t n_rows_capacity, int n_slices, int group,
                     int n_dense, int idx_dtype, int64_t row_offset, unsigned flags);
/* Append n rows.  vals: [n, n_slices*group + n_dense] (row stride in ELEMENTS), idx: [n, n_slices]
 * (ignored when n_slices == 0).  Host or device pointers. */
int dhr_index_append(dhr_index* h, int64_t n, int val_dtype, const void* vals, int64_t val_row_stride,
                     int idx_dtype, const void* idx, int64_t idx_row_stride);
/* Validates what was appended (fp16 representability, index range) and makes the index searchable. */
int dhr_index_finalize(dhr_index* h);
/* create + append + finalize in one call */
int dhr_index_open(dhr_index** out, int device, int64_t n_rows, int n_slices, int group, int n_dense,
                   int val_dtype, const void* vals, int64_t val_row_stride,
                   int idx_dtype, const void* idx, int64_t idx_row_stride, int64_t row_offset, unsigned flags);
int dhr_index_close(dhr_index* h);
int dhr_index_rows(const dhr_index* h, int64_t* n_rows);
int dhr_index_row_bytes(const dhr_index* h, int64_t* bytes);   /* HBM bytes per row of the resident layout */
int dhr_index_device_bytes(const dhr_index* h, int64_t* bytes); /* HBM bytes the index holds right now (all copies + workspaces) */

/* options: "scan_variant" (0|1), "query_block" (1|2|4|8), "query_groups" (1..64), "profile" (0|1), "tile_mode" (0|1),
 * "rowmajor" (0 = free the row-major arrays now and keep only the tiled copies; they are rebuilt on the first call that
 * needs them -- fp32 / lamda-scaled queries, --IP, rerank, overflow fallback; 1 = make them resident now);
 * tuning of the tile path (defaults are the measured best): "overlap" (0|1), "lanes" (1|2), "dense_variant" (0..3),
 * "dense_multicast" (0..2), "dense_prefetch" (0|1: TMA L2 prefetch ahead of the dense tile loads, default 0), "dense_lite"
 * (0|1: small-footprint dense kernel sharing the SMs with the lexical tile kernel, default 0), "lex_stages" (0 = auto, 2..8: ring
 * depth cap of the lexical tile kernel), "stream_priority" (0|1, before the first search) */
int dhr_index_set_option(dhr_index* h, const char* name, int64_t value);
int dhr_index_get_stats(const dhr_index* h, dhr_stats* out);

/* ---- search -------------------------------------------------------------------------------
 * Replaces GIP_retrieval exact branch (gip_retrieval.py:88-126,158-165), IP_retrieval (:60-85,
 * pass n_slices==0 index / q_idx NULL) and, with DHR_SEARCH_UNMASKED, the --IP first stage (:139).
 * q_vals [Q, W] fp16 or fp32 (the reference passes fp32), q_idx [Q, n_slices] any DHR_IDX_* dtype.
 * lamda multiplies the last n_dense query columns in fp32 exactly like :281-283 (pass 1 if the
 * caller already scaled).  Outputs [Q, k]: scores fp32, rows int64 (global = row_offset + local),
 * sorted by (score desc, row asc); when the index holds fewer than k rows the tail is
 * (-inf, -1) and out_counts[q] (optional) holds the number of valid entries.
 * Stream-ordered on `stream` (a cudaStream_t, NULL = default stream); the call returns after the
 * results have reached out_scores/out_rows. One search at a time per index. */
int dhr_search(dhr_index* h, int n_queries, int q_val_dtype, const void* q_vals, int64_t q_val_row_stride,
               int q_idx_dtype, const void* q_idx, int64_t q_idx_row_stride, float lamda, int k,
               unsigned flags, float* out_scores, int64_t* out_rows, int32_t* out_counts, void* stream);

/* ---- stream-ordered search for the sharded path ---------------------------------
 * Same search as dhr_search, but every launch is only ENQUEUED on `stream` and the call returns without waiting
 * (the one exception: fp32 or lamda != 1 queries need one flag read before the scan).  The result of query q, rank r is
 * the packed 64-bit key out_keys[q*k + r] = (order-preserving bits of the fp32 score << 32) | (0xFFFFFFFF - GLOBAL row);
 * 0 = padding.  Descending key order is (score desc, global row asc) on every shard, so merging shards (the step after
 * retrieval/merge.result.py:22-41 / the NCCL all-gather) is a plain 64-bit merge: dhr_merge_keys.  out_keys must be DEVICE
 * memory; row_offset + rows must be < 2^32 - 1.
 * Queries are processed in batches (dhr_search_batches); dhr_search_wait_batch makes another stream wait until batch b's
 * keys are final, which lets the caller exchange + merge batch b while batch b+1 is still being scanned.
 * dhr_search_complete synchronises, re-runs queries whose candidate buffer overflowed (adversarial row order; *n_rerun of
 * them, their keys are rewritten) and must be called before the next search on the index. */
int dhr_search_keys(dhr_index* h, int n_queries, int q_val_dtype, const void* q_vals, int64_t q_val_row_stride,
                    int q_idx_dtype, const void* q_idx, int64_t q_idx_row_stride, float lamda, int k,
                    unsigned flags, uint64_t* out_keys, void* stream);
int dhr_search_batches(const dhr_index* h, int* batch_size, int* n_batches);
int dhr_search_wait_batch(dhr_index* h, int batch, void* stream);
int dhr_search_complete(dhr_index* h, int* n_rerun, void* stream);
/* keys [P][part_stride] (each part holds [Q,k] packed keys, sorted descending per query) -> [Q,k] by key descending.
 * Writes (out_scores, out_rows) and / or out_keys (either may be NULL).  DEVICE pointers; stream-ordered, no host sync.
 * k <= 8192, any P. */
int dhr_merge_keys(int device, int n_parts, int n_queries, int k, const uint64_t* keys, int64_t part_stride,
                   float* out_scores, int64_t* out_rows, uint64_t* out_keys, void* stream);

/* Exact GIP on given candidate rows (rerank, gip_retrieval.py:142-150 and :205-215):
 * cand_rows [Q, M] LOCAL row ids (< 0 = skip).  Outputs as dhr_search (rows are global). */
int dhr_rerank(dhr_index* h, int n_queries, int q_val_dtype, const void* q_vals, int64_t q_val_row_stride,
               int q_idx_dtype, const void* q_idx, int64_t q_idx_row_stride, float lamda,
               const int64_t* cand_rows, int n_cand, int k, float* out_scores, int64_t* out_rows,
               int32_t* out_counts, void* stream);

/* Merge P per-shard top-k lists (replaces retrieval/merge.result.py:20-43): scores/rows [P, Q, k] -> [Q, k] by
 * (score desc, row asc); rows < 0 are padding.  Lists need not be sorted.  Runs on `device`; pointers may be host or
 * device; synchronous.  k <= 4096, any P (progressive merge, shared memory is O(k)). */
int dhr_topk_merge(int device, int n_parts, int n_queries, int k, const float* scores, const int64_t* rows,
                   float* out_scores, int64_t* out_rows, void* stream);

/* ---- next row: densify op on the device --------------------------------------
 * Replaces tevatron/DHR/utils.py:5-22 `densify` plus the fp16 / uint8 storage conversion of
 * tevatron/driver/encode.py:155-170,180-195: lexical_reps [batch, vocab] (fp32 or fp16, DEVICE memory) ->
 * drop the first remove_dims ids -> view(batch, R, dims) -> max over R.  Writes the fp16 values and the uint8
 * argmax (first maximum) straight into caller-provided DEVICE buffers with the given row strides (elements), e.g.
 * the value / index blocks of an index under construction. Stream-ordered, asynchronous. */
int dhr_densify(int device, int batch, int vocab, int dims, int remove_dims, int val_dtype, const void* reps,
                int64_t reps_row_stride, void* out_vals_f16, int64_t out_val_row_stride, uint8_t* out_idx,
                int64_t out_idx_row_stride, void* stream);

/* ---- next row: TREC run writer on the host -----------------------------------
 * Replaces the Python formatting loop of retrieval/gip_retrieval.py:329-342: for query q and rank r writes
 *   "{qid} Q0 {docid} {r+1} {score} {run_name}\n"
 * with {score} = Python's repr() of the fp32 score widened to double (what `.tolist()` + str.format produce).
 * rows [Q,k] index the docid table (LOCAL rows of the shard; < 0 = padding, skipped); counts [Q] may be NULL (= k).
 * Ids are either int64 arrays or strings packed in one buffer with [n + 1] byte offsets (exactly one form per table).
 * skip_equal != 0 drops lines whose docid equals the query id without renumbering the ranks (:340).
 * HOST pointers only; n_threads <= 0 uses all cores.  The file is created (append = 0) or appended to. */
int dhr_write_trec(const char* path, int append, int n_queries, int k, const int32_t* counts, const int64_t* rows,
                   const float* scores, const int64_t* qid_int, const char* qid_str, const int64_t* qid_off,
                   int64_t n_docids, const int64_t* docid_int, const char* docid_str, const int64_t* docid_off,
                   int skip_equal, const char* run_name, int n_threads, int64_t* lines_written);

/* Shard merge on the host, replacing retrieval/merge.result.py:20-43: reads the shards' TREC files in the given order,
 * groups lines by query id (order of first appearance), keeps the top `topk` per query by (score desc, position in the
 * concatenated shard lists asc) and writes out_path with ranks renumbered from 1 and scores printed as Python prints
 * float(text).  Lines must have the six space-separated fields gip_retrieval.py:341 writes. */
int dhr_merge_trec(int n_paths, const char* const* paths, const char* out_path, int topk, const char* run_name,
                   int n_threads, int64_t* lines_written);

#ifdef __cplusplus
}
#endif
#endif /* DHR_B200_H */
