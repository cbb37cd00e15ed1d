/*
 * neumann_b200.h — C ABI of the B200-native SIMILAR brute-force scan.
 *
 * The reference (Shadylukin/Neumann @ aae3d465) has NO FFI for this path: the seam is the
 * private trio "store.scan -> search_{sequential,parallel}[_with_metric] -> sort_by+truncate"
 * inside three public Rust methods.  Each entry point below names the reference code it
 * replaces (paths relative to the reference checkout).  INTEGRATION.md shows the Rust `-sys`
 * shim that binds them.
 *
 * Conventions
 *   - Every function returns an nm_status (0 = ok) unless stated otherwise; nothing throws or
 *     aborts across the ABI.  nm_last_error() returns a thread-local message for the last
 *     non-zero status on the calling thread.
 *   - The caller owns every host buffer it passes (in and out).  The library owns device
 *     memory, pinned staging and streams.  The row-index <-> key mapping stays with the
 *     caller, in mirror (row) order, like `HnswCacheEntry = (Arc<HNSWIndex>, Vec<String>)`
 *     (vector_engine/src/lib.rs:98).
 *   - nm_search* may be called concurrently from many host threads on one index
 *     (`VectorEngine: Send + Sync`, vector_engine/src/lib.rs:1127-1131; concurrent search +
 *     store is tested at :5615-5711).  Mutations take the index write lock and never tear an
 *     in-flight search.
 *   - Results per query: min(k, rows) hits, sorted by score descending; exact-score ties
 *     (-0.0 == +0.0) by ascending row; NaN scores last.  Scores are bit-identical to the
 *     reference arithmetic restated in oracle/nm_oracle.c.
 *   - There is no CPU fallback.  Without a CUDA device every compute entry point fails with
 *     NM_ERR_STORAGE.
 */
#ifndef NEUMANN_B200_H
#define NEUMANN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NM_ABI_VERSION 1

/* DistanceMetric (vector_engine/src/lib.rs:281-289).  Score conventions follow compute_score
 * (lib.rs:2231-2246): cosine = dot/(|q||x|) with zero guards, dot = raw, euclidean = 1/(1+d). */
typedef enum nm_metric { NM_COSINE = 0, NM_EUCLIDEAN = 1, NM_DOT_PRODUCT = 2 } nm_metric;

/* Status codes map 1:1 onto VectorError (vector_engine/src/lib.rs:102-149). */
typedef enum nm_status {
    NM_OK = 0,
    NM_ERR_EMPTY_VECTOR = 1,       /* VectorError::EmptyVector                                */
    NM_ERR_INVALID_TOP_K = 2,      /* VectorError::InvalidTopK                                */
    NM_ERR_DIMENSION_MISMATCH = 3, /* VectorError::DimensionMismatch                          */
    NM_ERR_STORAGE = 4,            /* VectorError::StorageError(msg): CUDA / NCCL / alloc     */
    NM_ERR_SEARCH_TIMEOUT = 5,     /* VectorError::SearchTimeout                              */
    NM_ERR_INVALID_ARGUMENT = 6,   /* null pointer, bad metric, bad device, row out of range  */
    NM_ERR_NOT_FOUND = 7,          /* VectorError::NotFound                                   */
    NM_ERR_CONFIGURATION = 8,      /* VectorError::ConfigurationError                         */
    NM_ERR_COLLECTION_EXISTS = 9,  /* VectorError::CollectionExists                           */
    NM_ERR_COLLECTION_NOT_FOUND = 10 /* VectorError::CollectionNotFound                       */
} nm_status;

typedef struct nm_index nm_index;

/* Largest k served by ONE scan pass; larger k is served exactly by ceil(k/1024) chained passes
 * (each a full scan admitting only hits below the previous pass's last one). */
#define NM_TOPK_FAST_MAX 1024u

/* ---- library ------------------------------------------------------------------------ */
int nm_abi_version(void);
const char *nm_last_error(void);
/* Number of visible CUDA devices (0 on a CPU-only host; never fails). */
int nm_device_count(void);

/* ---- device mirror of the `emb:` rows (replaces TensorStore::scan + get as the data
 *      source of the scan: tensor_store/src/lib.rs:948-999, slab_router.rs:217-305) -------- */

/* Create an empty mirror for rows of `dim` floats, sharded by contiguous row range over
 * `n_dev` devices of THIS process (devices == NULL, n_dev == 0 -> current device only). */
int nm_index_create(uint32_t dim, const int *devices, int n_dev, nm_index **out);
void nm_index_destroy(nm_index *idx);

/* Replace the mirror's contents with `n` row-major rows [n, dim] from host memory (pinned or
 * pageable).  Rows are split into contiguous ranges [g*n/G, (g+1)*n/G) over the devices.  */
int nm_index_load(nm_index *idx, const float *rows, uint64_t n);
/* Append `n` rows after the current last row (store_embedding of new keys, lib.rs:1840-1868).
 * The mirror grows IN PLACE: more physical memory is mapped behind the existing rows (CUDA
 * virtual memory management; the reference's own precedent is the chunked EmbeddingSlab,
 * tensor_store/src/embedding_slab.rs:27, 92-124), nothing is copied, so a mirror can grow up to
 * the free memory of the device.  On an in-process multi-device index the rows land on the last
 * shard (the end of the global row order); once a shard holds more than 1.5x its share the
 * rows are re-split into equal contiguous ranges — global row ids never change. */
int nm_index_append(nm_index *idx, const float *rows, uint64_t n);
/* Overwrite one row in place (store_embedding of an existing key). */
int nm_index_update(nm_index *idx, uint64_t row, const float *vec);
/* Delete row `row` by moving the LAST row into its slot (delete_embedding, lib.rs:1913-1925).
 * The caller applies the same swap to its key table.  *moved_from receives the index of the
 * row that was moved (== row when the last row itself was removed). */
int nm_index_swap_remove(nm_index *idx, uint64_t row, uint64_t *moved_from);
int nm_index_clear(nm_index *idx);
/* Copy row `row` back to host (diagnostics / tests). */
int nm_index_get_row(nm_index *idx, uint64_t row, float *out_vec);
/* Copy rows [first, first + n) back to host as a dense [n, dim] array (pinned or pageable).
 * Diagnostics: bench.py and the parity tests hand exactly the bytes the GPU scanned to the CPU
 * oracle this way. */
int nm_index_get_rows(nm_index *idx, uint64_t first, uint64_t n, float *out_rows);

uint64_t nm_index_rows(const nm_index *idx);
uint32_t nm_index_dim(const nm_index *idx);
int nm_index_device_count(const nm_index *idx);
/* Diagnostics: where shard `shard` (0 .. nm_index_device_count-1) stands. */
typedef struct nm_shard_info {
    int device;
    uint64_t rows;            /* rows held                                                 */
    uint64_t row_base;        /* global id of its first row                                 */
    uint64_t capacity_rows;   /* rows the mapped part of the mirror can hold                */
    uint64_t mapped_bytes;    /* physical memory behind the f32 mirror                      */
    uint64_t reserved_bytes;  /* virtual address range reserved for it                      */
    uint64_t chunks;          /* physical allocations mapped into the range                 */
    uint64_t remaps;          /* times the range was replaced by a larger one (no copy)     */
    uint64_t q8_rows;         /* rows of the int8 copy that are current (0 = no copy)       */
    int grows_in_place;       /* 1 = virtual memory management, 0 = realloc-and-copy fallback */
} nm_shard_info;
int nm_index_shard_info(nm_index *idx, int shard, nm_shard_info *out);

/* Fill the mirror ON DEVICE with the synthetic corpus of SURVEY 8d:
 *   x[r,c] = u24(splitmix64(splitmix64(seed) ^ ((row_offset + r)*dim + c))) * 2^-23 - 1
 * (bit-identical to oracle nmo_fill_synthetic).  Benchmark / test utility. */
int nm_index_fill_synthetic(nm_index *idx, uint64_t n, uint64_t seed, uint64_t row_offset);

/* ---- the scan (replaces search_sequential/parallel[_with_metric] + sort_by + truncate:
 *      vector_engine/src/lib.rs:2013-2034, 2070-2099, 1648-1686) ------------------------- */

/* queries: [nq, dim] host floats.  Outputs (host): out_rows [nq, k] GLOBAL row ids,
 * out_scores [nq, k], out_counts [nq] = hits written for that query (min(k, rows)); unused
 * slots are left untouched.  Validation mirrors the reference: nq == 0 or dim == 0 ->
 * NM_ERR_EMPTY_VECTOR, k == 0 -> NM_ERR_INVALID_TOP_K.  The zero-magnitude-query
 * short-circuit (lib.rs:1970-1974, 2066) stays in the host wrapper; here a zero cosine query
 * scores every row 0.0 exactly as cosine_similarity does (lib.rs:2261-2263).               */
int nm_search(nm_index *idx, const float *queries, uint32_t nq, uint32_t k, int metric,
              uint64_t *out_rows, float *out_scores, uint32_t *out_counts);

/* Pre-filtered scan (replaces search_with_pre_filter, vector_engine/src/lib.rs:3514-3557:
 * "filter first, then search the subset").  row_mask is a host bitmask over the mirror's rows,
 * bit (r % 64) of word r / 64 set = row r is eligible; ceil(rows / 64) words.  Ineligible rows
 * never rank, and 256-row blocks without any eligible row are not even read from HBM.
 * In-process multi-device indexes take their slices of the mask; on a collective index the mask
 * covers THIS process's rows (local row r = bit r). */
int nm_search_masked(nm_index *idx, const float *queries, uint32_t nq, uint32_t k, int metric,
                     const uint64_t *row_mask, uint64_t *out_rows, float *out_scores,
                     uint32_t *out_counts);

/* ---- metadata columns + device-side filter evaluation (finishes search_with_pre_filter,
 *      vector_engine/src/lib.rs:3514-3557, and search_filtered_in_collection :1698-1829: the
 *      FilterCondition tree is evaluated ON THE DEVICE into the row bitmask of the scan, with
 *      the type rules of evaluate_filter / compare_tensor_value_to_filter, lib.rs:3592-3684) ---- */
typedef enum nm_value_tag {
    NM_V_MISSING = 0, /* the row has no such field                                   */
    NM_V_NULL = 1,
    NM_V_BOOL = 2,    /* value 0 / 1                                                  */
    NM_V_INT = 3,     /* value = the int64                                            */
    NM_V_FLOAT = 4,   /* value = the bits of the double                               */
    NM_V_STRING = 5   /* value = code of the string in the CALLER's dictionary of that column */
} nm_value_tag;
/* Set the metadata of rows [first_row, first_row + n) in column `column` (one column per field;
 * ids are the caller's).  Rows never set read as NM_V_MISSING; appended rows start missing;
 * swap_remove moves the column entries with the row; load / clear / fill_synthetic drop all
 * columns.  Works on single-device, in-process multi-device and collective indexes (rows are
 * this process's rows). */
int nm_index_column_set(nm_index *idx, uint32_t column, uint64_t first_row, uint64_t n,
                        const uint8_t *tags, const uint64_t *values);
typedef enum nm_filter_kind {
    NM_F_TRUE = 0, NM_F_FALSE = 1,
    NM_F_AND = 2, NM_F_OR = 3,     /* pop two, push one                                        */
    NM_F_EXISTS = 4,               /* the row has the field                                     */
    NM_F_CMP = 5,                  /* field <cmp> literal; literal NULL / BOOL / INT / FLOAT    */
    NM_F_STR_TABLE = 6             /* the row holds a string whose code has its bit set in the
                                      table: the caller evaluated the string predicate (=, <,
                                      CONTAINS, STARTS_WITH, IN ...) once per DISTINCT string   */
} nm_filter_kind;
typedef enum nm_filter_cmp { NM_C_EQ = 0, NM_C_NE, NM_C_LT, NM_C_LE, NM_C_GT, NM_C_GE } nm_filter_cmp;
/* One op of a filter in POSTFIX order (<= 128 ops, stack depth <= 64).  A missing field or an
 * incomparable pair of types makes every comparison false, `!=` included (lib.rs:3642-3656). */
typedef struct nm_filter_op {
    uint8_t kind;         /* nm_filter_kind                                     */
    uint8_t cmp;          /* nm_filter_cmp (NM_F_CMP)                           */
    uint8_t lit_tag;      /* nm_value_tag of the literal (NM_F_CMP)             */
    uint8_t reserved;
    uint32_t column;      /* leaves                                             */
    uint64_t lit;         /* NM_F_CMP: int64 / double bits / bool               */
    uint32_t table_off;   /* NM_F_STR_TABLE: first u32 word of its bit table in `tables` */
    uint32_t table_bits;  /* NM_F_STR_TABLE: number of dictionary codes covered */
} nm_filter_op;
/* nm_search over the rows that pass the filter.  The mask is computed on the device (one small
 * kernel, HBM-bound on 9 bytes per row and referenced column), cached per filter until the next
 * mutation, and applied inside the scan: ineligible rows never rank and 256-row blocks without
 * an eligible row are not read.  Single-device, in-process multi-device and collective indexes
 * (every rank passes the same program; each shard evaluates its own rows). */
int nm_search_filtered(nm_index *idx, const float *queries, uint32_t nq, uint32_t k, int metric,
                       const nm_filter_op *program, uint32_t n_ops, const uint32_t *tables,
                       uint32_t n_table_words, uint64_t *out_rows, float *out_scores,
                       uint32_t *out_counts);
/* Diagnostics / count_matching: evaluate the filter only.  out_mask (may be NULL): bit (r % 64)
 * of word r / 64 = row r passes, ceil(rows / 64) words over this process's rows;
 * *out_eligible (may be NULL) = number of passing rows. */
int nm_index_filter_mask(nm_index *idx, const nm_filter_op *program, uint32_t n_ops,
                         const uint32_t *tables, uint32_t n_table_words, uint64_t *out_mask,
                         uint64_t *out_eligible);

/* Same scan with the query and the outputs resident in DEVICE memory of the index's first
 * device and the work enqueued on `stream` (a cudaStream_t).  With a caller stream the call
 * is fully ASYNCHRONOUS: it returns after enqueueing and the outputs are valid once the
 * stream reaches that point.  NULL = library stream, and the call then synchronises before
 * returning.  Single-device indexes only.  Mutations (load / append / update / swap_remove /
 * clear) wait for every asynchronous search issued before them, whatever the stream.  The
 * library keeps a scratch workspace per caller stream: call nm_index_release_stream before
 * destroying a stream that was used here.  Collective indexes: asynchronous searches are ordered
 * by ONE stream at a time; switching streams synchronises the previous one. */
int nm_search_device(nm_index *idx, const float *d_queries, uint32_t nq, uint32_t k,
                     int metric, uint64_t *d_out_rows, float *d_out_scores,
                     uint32_t *d_out_counts, void *stream);

/* Forget (and free the scratch of) a caller stream used with nm_search_device; synchronises it. */
int nm_index_release_stream(nm_index *idx, void *stream);
/* enable = 1: consecutive asynchronous single-query nm_search_device calls on one caller stream
 * OVERLAP (programmatic dependent launch): the next query's CTAs start streaming on the SMs the
 * current query has already left, while its last CTA still merges the per-SM candidates and
 * exchanges hits with the other shards; outputs and completion stay in call order.  The price
 * is a rule for the caller, which is why it is opt-in: the query of call i+1 is read while call
 * i is still running, so it must not be produced by a KERNEL enqueued on that stream after call
 * i-1 (copies — cudaMemcpyAsync — and anything enqueued earlier are fine).  Ignored (calls run
 * back to back) while profiling is on, for k > 1024, and for batches that share a corpus pass.
 * Default 0. */
int nm_index_set_pipelining(nm_index *idx, int enable);

/* ---- row-range sharding across processes (one process per GPU; replaces
 *      QueryRouter::execute_scatter_gather + ResultMerger::merge_top_k,
 *      query_router/src/lib.rs:1818-1900, distributed.rs:413-433) ------------------------- */
#define NM_COMM_ID_BYTES 128
/* Rank 0 creates the id and ships it to the other ranks over its own transport. */
int nm_comm_create_id(void *out_id /* NM_COMM_ID_BYTES */);
/* Attach this index (one shard = this process's rows) to an n_ranks communicator.
 * `row_base` = global index of this shard's first row; shards must be attached in ascending
 * row order by rank so that rank order == global row order (merge_top_k concatenates shards
 * in shard order).  After attach, nm_search / nm_search_device are COLLECTIVE: every rank
 * calls them with the same queries, k and metric, and every rank receives the merged result.
 * The exchange is ONE ncclAllGather of nq*k 16-byte candidates per rank. */
int nm_index_attach_comm(nm_index *idx, const void *id, int n_ranks, int rank,
                         uint64_t row_base);
/* Detaching (and nm_index_destroy of an attached index) tears the communicator down: like
 * ncclCommDestroy it must be called by EVERY rank; a rank that destroys its shard while the others
 * wait for it elsewhere deadlocks. */
int nm_index_detach_comm(nm_index *idx);

/* ---- instrumentation (counters the reference keeps in ShardAccessTracker,
 *      tensor_store/src/instrumentation.rs:1-40) ------------------------------------------ */
typedef struct nm_stats {
    uint64_t searches;        /* queries served                                   */
    uint64_t rows_scanned;    /* rows x queries                                   */
    uint64_t bytes_streamed;  /* algorithmic bytes: rows * dim * 4 per query pass */
    uint64_t scan_launches;   /* scan kernel launches                             */
    uint64_t merge_launches;  /* cross-shard merge kernel launches                */
    uint64_t h2d_bytes;       /* staging + query uploads                          */
    uint64_t d2h_bytes;       /* result downloads                                 */
    double last_scan_ms;      /* device time of the most recent nm_search scan(s); measured only
                                 while nm_index_set_profiling is on (0 otherwise) */
    double profiled_scan_ms;  /* sum of CUDA-event times around profiled scan launches    */
    uint64_t profiled_scans;  /* number of nm_search_device calls folded into the sum     */
    uint64_t prefilter_queries;   /* queries served through the int8 pre-filter            */
    uint64_t prefilter_fallbacks; /* of those, redone with the exact f32 scan              */
    uint64_t prefilter_kept;      /* rows whose score interval reached the running bound   */
    uint64_t coalesced_batches;   /* leader rounds of the single-query coalescer           */
    uint64_t coalesced_queries;   /* queries served by those rounds                        */
    uint64_t tc_queries;          /* queries served through the tensor-core batch pre-filter */
    uint64_t tc_fallbacks;        /* of those, redone by the exact path                     */
    uint64_t tc_survivors;        /* (row, query) pairs re-scored exactly                    */
    uint64_t filter_masks_built;  /* filter programs evaluated on the device (per shard)     */
    uint64_t filter_mask_hits;    /* filtered searches served from the per-filter mask cache */
} nm_stats;
int nm_index_stats(nm_index *idx, nm_stats *out);
/* When enabled, every asynchronous nm_search_device call brackets its scan launches (not the
 * all-gather / merge) with CUDA events on the caller's stream; nm_index_stats waits for them
 * and accumulates profiled_scan_ms / profiled_scans.  Used by bench.py for the roofline.
 * Synchronous searches (nm_search / _masked / _filtered) bracket their device work the same way
 * and report it as last_scan_ms.  Off by default: the two timing events of a call cost ~8 us,
 * a fifth of a query over a 10k-row corpus. */
int nm_index_set_profiling(nm_index *idx, int enable);
/* SURVEY 8f row 4 — quantised pre-filter with EXACT re-score (precedent:
 * ScalarQuantizedVector, tensor_store/src/hnsw.rs:308-356).  An int8 copy of the mirror (+1 byte
 * per element, +24 bytes per row) lets a pass read 4x fewer HBM bytes; every row's reference
 * score is bracketed in a rigorous interval and only rows whose interval reaches the k-th best
 * lower bound are re-scored with the exact f32 arithmetic, so results are bit-identical to the
 * f32 scan (tested).  Modes:
 *   2 (default, "auto")  BATCHES (nq >= 2, and coalesced rounds of concurrent single queries)
 *      use the tensor-core pre-filter below.  The copy is built by the first eligible batch —
 *      provided it leaves max(4 GiB, 1/16 of the device) of HBM free; otherwise batches stay on
 *      the exact kernels — and is kept up to date by every mutation from then on.  Single
 *      queries stay on the f32 scan (the headline path).
 *   1  the copy is built now and single queries use it too: dp4a scan + exact re-score
 *      (cosine and dot product, k <= 1024; anything else, a non-finite query or an
 *      overflowing candidate list falls back to the f32 scan).  Changes the bytes read per
 *      row, so it is benchmarked separately from the f32 roofline.
 *   0  no copy (frees it); every search runs on the exact f32 kernels. */
int nm_index_set_prefilter(nm_index *idx, int mode);
/* Batches (nq >= 2) on an index whose int8 copy exists (modes 1 and 2 above) go through the
 * tensor-core pre-filter
 * (tc_prefilter_kernels.cuh): ONE pass over the int8 copy serves up to 256 queries as an exact
 * s8 x s8 -> s32 GEMM on the tcgen05 tensor cores (accumulators in TMEM); every integer dot
 * product is turned into a rigorous score interval (cosine, dot product AND the reference's
 * scalar L2 chain), only entries whose interval reaches the query's running k-th best lower
 * bound are kept, and those are re-scored from the f32 mirror with the reference arithmetic.
 * Results are bit-identical to the exact batched kernels (tested); queries that are not
 * finite or whose candidate list overflows are redone by the exact path.  Shards of >= 65536
 * rows (single-device, in-process multi-device and collective indexes alike: a sharded index
 * computes each shard's hits this way and merges them as usual), k <= 1024.  enable = 0 keeps
 * batches on the exact kernels. */
int nm_index_set_tensor_core(nm_index *idx, int enable);
/* Diagnostics: the exact integer dot products the tensor-core pass computes,
 * out[q * rows + r] = sum_i int8(row r)[i] * int8(query q)[i], for the first min(nq, 256)
 * queries (host buffers).  Needs the int8 copy (pre-filter mode 1, or mode 2 after a first
 * eligible batch) and a single-device index. */
int nm_debug_tc_dots(nm_index *idx, const float *queries, uint32_t nq, int32_t *out);
/* Diagnostics: the int8 copy of row `row` and its scale (x ~ scale * int8). */
int nm_debug_q8_row(nm_index *idx, uint64_t row, int8_t *out_q8, float *out_scale);
/* Concurrent single-query nm_search calls (many host threads, the reference's serving pattern:
 * stress_tests/tests/mixed_workload_stress.rs:291-307) are coalesced: calls that arrive while
 * the GPU is busy ride the next corpus pass together, up to max_batch per pass (default 64;
 * 1 disables).  No timers are involved, so a lone call is never delayed; results are
 * bit-identical to isolated calls. */
int nm_index_set_coalescing(nm_index *idx, int max_batch);
/* Batches of >= 2 queries share one corpus pass (batch_kernels.cuh) by default; 0 forces one
 * scan per query.  Results are bit-identical either way (tested); this is a tuning knob. */
int nm_index_set_batching(nm_index *idx, int enable);

#ifdef __cplusplus
}
#endif
#endif /* NEUMANN_B200_H */
