/*
 * neumann_b200_engine.h — C view of the C++ host mirror (neumann_b200/csrc/vector_engine.hpp,
 * similar_router.hpp) so that non-C++ hosts and the Python test harness can drive the same
 * VectorEngine / SIMILAR surface the reference exposes in Rust.  The drop-in BOUNDARY for a
 * Rust host is include/neumann_b200.h; this header is the reference-facing API rebuilt on top
 * of it.  Function names follow vector_engine/src/lib.rs method names.
 */
#ifndef NEUMANN_B200_ENGINE_H
#define NEUMANN_B200_ENGINE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct nm_engine nm_engine;
typedef struct nm_results nm_results;

/* VectorEngineConfig (vector_engine/src/lib.rs:626-663).  0 / negative = "None". */
typedef struct nm_engine_config {
    uint64_t default_dimension;  /* 0 = None */
    float sparse_threshold;      /* default 0.5 */
    uint64_t parallel_threshold; /* default 5000 */
    int default_metric;          /* nm_metric */
    uint64_t max_dimension;      /* 0 = None */
    int64_t search_timeout_ms;   /* < 0 = None */
    int n_devices;               /* 0 = current device */
    int devices[8];
    int device_prefilter;        /* 1 = int8 copy of every mirror: dp4a / tensor-core pre-filters
                                    (bit-identical results; device-side addition, default 0) */
    uint64_t max_keys_per_scan;  /* 0 = None (lib.rs:638): bounds list_keys / clear */
} nm_engine_config;

void nm_engine_config_default(nm_engine_config *cfg);
/* VectorEngine::new / with_config (validates; NM_ERR_CONFIGURATION on a bad config). */
int nm_engine_create(const nm_engine_config *cfg /* NULL = defaults */, nm_engine **out);
void nm_engine_destroy(nm_engine *e);
const char *nm_engine_last_error(void);

int nm_engine_store_embedding(nm_engine *e, const char *key, const float *vec, size_t n);
int nm_engine_get_embedding(nm_engine *e, const char *key, float *out, size_t cap, size_t *len);
int nm_engine_delete_embedding(nm_engine *e, const char *key);
int nm_engine_exists(nm_engine *e, const char *key);
uint64_t nm_engine_count(nm_engine *e);

int nm_engine_search_similar(nm_engine *e, const float *query, size_t n, size_t top_k,
                             nm_results **out);
int nm_engine_search_similar_with_metric(nm_engine *e, const float *query, size_t n, size_t top_k,
                                         int metric, nm_results **out);
/* Batch form (no reference counterpart: there a batch is nq calls): queries [nq, n] row-major,
 * out[i] = what nm_engine_search_similar_with_metric returns for query i.  out must hold nq
 * pointers; free each with nm_results_free. */
int nm_engine_search_similar_batch(nm_engine *e, const float *queries, size_t nq, size_t n,
                                   size_t top_k, int metric, nm_results **out);
int nm_engine_compute_similarity(const float *a, size_t na, const float *b, size_t nb,
                                 float *out);

int nm_engine_create_collection(nm_engine *e, const char *name, uint64_t dimension /*0=None*/,
                                int metric);
int nm_engine_delete_collection(nm_engine *e, const char *name);
int nm_engine_collection_exists(nm_engine *e, const char *name);
int nm_engine_store_in_collection(nm_engine *e, const char *collection, const char *key,
                                  const float *vec, size_t n);
int nm_engine_delete_from_collection(nm_engine *e, const char *collection, const char *key);
uint64_t nm_engine_collection_count(nm_engine *e, const char *collection);
int nm_engine_search_in_collection(nm_engine *e, const char *collection, const float *query,
                                   size_t n, size_t top_k, nm_results **out);

/* Metadata + filtered search (vector_engine/src/lib.rs:3429-3557, 1698-1829).
 * `metadata` is a typed wire string: fields separated by 0x1f, each  name 0x1e type 0x1e value
 * with type one of i (int64) f (double) s (string) b ("0"/"1") n (null).
 * `where_expr` uses the SIMILAR ... WHERE grammar (field op literal, AND / OR, parentheses)
 * plus EXISTS(f), CONTAINS(f,'s'), STARTS_WITH(f,'s'), f IN (v, ...); "TRUE" matches all.
 * strategy: 0 auto, 1 pre-filter (device scan under a row bitmask), 2 post-filter. */
int nm_engine_store_embedding_with_metadata(nm_engine *e, const char *key, const float *vec,
                                            size_t n, const char *metadata);
int nm_engine_store_in_collection_with_metadata(nm_engine *e, const char *collection,
                                                const char *key, const float *vec, size_t n,
                                                const char *metadata);
int nm_engine_search_similar_filtered(nm_engine *e, const float *query, size_t n, size_t top_k,
                                      const char *where_expr, int strategy,
                                      size_t oversample_factor, nm_results **out);
int nm_engine_search_filtered_in_collection(nm_engine *e, const char *collection,
                                            const float *query, size_t n, size_t top_k,
                                            const char *where_expr, int strategy,
                                            size_t oversample_factor, nm_results **out);
int nm_engine_count_matching(nm_engine *e, const char *where_expr, uint64_t *out);
/* Tests / diagnostics (no device involved): the metadata columns the rows of dimension `dim` of
 * the default space are pushed to the device as, the postfix nm_filter_op program `where_expr`
 * compiles to, and the host's evaluate_filter verdict per row, as JSON text.  *out_len receives
 * the full length; `out` (may be NULL) receives at most out_cap - 1 bytes + NUL. */
int nm_engine_debug_filter_program(nm_engine *e, uint32_t dim, const char *where_expr, char *out,
                                   size_t out_cap, size_t *out_len);
/* Metadata maintenance (vector_engine/src/lib.rs:3346-3420); `metadata_wire` as for
 * nm_engine_store_embedding_with_metadata.  The device-side metadata columns follow at the next
 * filtered search. */
int nm_engine_update_metadata(nm_engine *e, const char *key, const char *metadata_wire);
int nm_engine_remove_metadata_field(nm_engine *e, const char *key, const char *field);
int nm_engine_has_metadata_field(nm_engine *e, const char *key, const char *field);
/* lib.rs:2340-2354, 2924-2940: *out = embeddings deleted (clear: at most max_keys_per_scan per
 * call).  keys_wire = keys separated by 0x1f. */
int nm_engine_clear(nm_engine *e, uint64_t *out);
int nm_engine_batch_delete_embeddings(nm_engine *e, const char *keys_wire, uint64_t *out);
/* lib.rs:2988-3058: search_similar / search_entities (entities != 0) with pagination.  limit < 0 =
 * no limit.  *total_count receives the number of ranked hits when count_total != 0 (else
 * UINT64_MAX); *has_more as the reference computes it. */
int nm_engine_search_paginated(nm_engine *e, int entities, const float *query, size_t n, size_t top_k,
                               size_t skip, int64_t limit, int count_total, nm_results **out,
                               uint64_t *total_count, int *has_more);
/* PointsService::query post-processing (neumann_server/src/service/points.rs:449-485);
 * score_threshold is ignored when has_threshold == 0. */
int nm_engine_query_points(nm_engine *e, const char *collection, const float *vector, size_t n,
                           size_t limit, size_t offset, int has_threshold, float score_threshold,
                           nm_results **out);

/* Unified entity mode (vector_engine/src/lib.rs:3060-3219): `_embedding` of entity keys. */
int nm_engine_set_entity_embedding(nm_engine *e, const char *entity_key, const float *vec, size_t n);
int nm_engine_remove_entity_embedding(nm_engine *e, const char *entity_key);
int nm_engine_entity_has_embedding(nm_engine *e, const char *entity_key);
int nm_engine_search_entities(nm_engine *e, const float *query, size_t n, size_t top_k,
                              nm_results **out);

/* QueryRouter::execute (legacy string path) / execute_parsed (AST path), SIMILAR + EMBED only.
 * *out is NULL for every result that is not QueryResult::Similar (Empty; the Value of EMBED GET
 * and the Count of EMBED DELETE / BATCH are available through the C++ API only). */
int nm_engine_execute(nm_engine *e, const char *command, nm_results **out);
int nm_engine_execute_parsed(nm_engine *e, const char *command, nm_results **out);

/* Device mirror introspection: rows held on host / on device for dimension `dim`. */
int nm_engine_mirror_rows(nm_engine *e, uint32_t dim, uint64_t *host_rows, uint64_t *device_rows);

size_t nm_results_len(const nm_results *r);
const char *nm_results_key(const nm_results *r, size_t i);
float nm_results_score(const nm_results *r, size_t i);
void nm_results_free(nm_results *r);

#ifdef __cplusplus
}
#endif
#endif
