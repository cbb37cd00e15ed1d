/*
 * nm_oracle.c — CPU restatement of Neumann's SIMILAR brute-force scan.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, the smoke check in
 * __graft_entry__.py and bench.py's cpu_baseline / --impl reference legs may load it.
 * The product path (neumann_b200/csrc) never links, loads or calls anything here.
 *
 * What it restates (paths relative to the reference checkout, commit aae3d465):
 *   - tensor_store/src/hnsw.rs:168-193   simd::dot_product      (wide::f32x8 lane tree)
 *   - tensor_store/src/hnsw.rs:198-222   simd::sum_of_squares
 *   - tensor_store/src/hnsw.rs:227-229   simd::magnitude
 *   - vector_engine/src/lib.rs:2249-2253 euclidean_distance     (scalar left fold)
 *   - vector_engine/src/lib.rs:2257-2266 cosine_similarity      (zero guards, no clamp)
 *   - vector_engine/src/lib.rs:2231-2246 compute_score          (1/(1+d) for Euclidean)
 *   - vector_engine/src/lib.rs:2013-2034 score all rows, stable sort desc, truncate k
 *   - query_router/src/distributed.rs:413-433 merge_top_k       (concat, stable sort, truncate)
 *
 * Third-party arithmetic that is NOT under /root/reference: `wide` 0.7.33 (Cargo.lock:4121).
 * Semantics relied on (crate documentation): f32x8 `*` is a lane-wise IEEE-754 multiply,
 * `+=` a lane-wise IEEE-754 add, no fusion, identical on every backend.  Rust std:
 * `Iterator::sum::<f32>()` is a left fold starting from 0.0 (MSRV 1.75 semantics, Cargo.toml:38;
 * compilers >= 1.83 start from -0.0, which can only change the sign of an all-(-0.0) sum),
 * `f32::sqrt` and `/` are correctly rounded, `slice::sort_by` is stable.
 *
 * Pinning status: the reference's own tests hold only tolerance KATs for this path
 * (tests/golden/reference_kats.json re-expresses every one; all pass against this file).
 * The reference cannot be compiled here (no rustc/cargo), so the BIT-LEVEL summation order
 * is pinned by the reference source only: "bit-level parity unpinned, KAT-pinned at 1e-6".
 *
 * Deliberate tightenings (the reference leaves these unspecified):
 *   - exact-score ties: ascending row index (reference order is HashSet-random,
 *     tensor_store/src/slab_router.rs:287-305; its own test accepts either order,
 *     vector_engine/src/lib.rs:6001-6019).  -0.0 and +0.0 tie, as under partial_cmp.
 *   - NaN scores rank last, ties among them by ascending row.
 *
 * Build: gcc -O3 -march=x86-64-v3 -ffp-contract=off -fno-fast-math -pthread (see Makefile).
 * -ffp-contract=off is REQUIRED: GCC's default may fuse l + a*b into an FMA.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#if defined(__FAST_MATH__)
#error "nm_oracle.c must not be built with -ffast-math"
#endif

enum { NMO_COSINE = 0, NMO_EUCLIDEAN = 1, NMO_DOT = 2 };

/* ---- hnsw.rs:168-193 ---------------------------------------------------------------- */
float nmo_dot_product(const float *a, const float *b, uint64_t n) {
    uint64_t chunks = n / 8, rem = n % 8;
    volatile float lane[8]; /* volatile: forbid re-association / vector re-shaping games */
    float l[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (uint64_t c = 0; c < chunks; ++c) {
        const float *pa = a + c * 8, *pb = b + c * 8;
        for (int j = 0; j < 8; ++j) {
            float p = pa[j] * pb[j]; /* va * vb  */
            l[j] = l[j] + p;         /* sum += .. */
        }
    }
    for (int j = 0; j < 8; ++j) lane[j] = l[j];
    float r = 0.0f; /* arr.iter().sum() : left fold from 0.0 */
    for (int j = 0; j < 8; ++j) r = r + lane[j];
    uint64_t start = chunks * 8;
    for (uint64_t i = 0; i < rem; ++i) {
        float p = a[start + i] * b[start + i];
        r = r + p;
    }
    return r;
}

/* ---- hnsw.rs:198-222 ---------------------------------------------------------------- */
float nmo_sum_of_squares(const float *v, uint64_t n) {
    uint64_t chunks = n / 8, rem = n % 8;
    volatile float lane[8];
    float l[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (uint64_t c = 0; c < chunks; ++c) {
        const float *pv = v + c * 8;
        for (int j = 0; j < 8; ++j) {
            float p = pv[j] * pv[j];
            l[j] = l[j] + p;
        }
    }
    for (int j = 0; j < 8; ++j) lane[j] = l[j];
    float r = 0.0f;
    for (int j = 0; j < 8; ++j) r = r + lane[j];
    uint64_t start = chunks * 8;
    for (uint64_t i = 0; i < rem; ++i) {
        float p = v[start + i] * v[start + i];
        r = r + p;
    }
    return r;
}

/* ---- hnsw.rs:227-229 ---------------------------------------------------------------- */
float nmo_magnitude(const float *v, uint64_t n) { return sqrtf(nmo_sum_of_squares(v, n)); }

/* ---- lib.rs:2249-2253 : scalar sequential fold, then sqrt ----------------------------- */
float nmo_euclidean_distance(const float *a, const float *b, uint64_t n) {
    float s = 0.0f;
    for (uint64_t i = 0; i < n; ++i) {
        float d = a[i] - b[i];
        float p = d * d;
        s = s + p;
    }
    return sqrtf(s);
}

/* ---- lib.rs:2257-2266 ----------------------------------------------------------------- */
float nmo_cosine_similarity(const float *a, const float *b, uint64_t n, float a_magnitude) {
    float dot = nmo_dot_product(a, b, n);
    float b_magnitude = nmo_magnitude(b, n);
    if (a_magnitude == 0.0f || b_magnitude == 0.0f) return 0.0f;
    float den = a_magnitude * b_magnitude;
    return dot / den;
}

/* ---- lib.rs:2231-2246 ----------------------------------------------------------------- */
float nmo_compute_score(const float *query, const float *stored, uint64_t n, float qmag,
                        int metric) {
    switch (metric) {
    case NMO_COSINE:
        return nmo_cosine_similarity(query, stored, n, qmag);
    case NMO_DOT:
        return nmo_dot_product(query, stored, n);
    default: {
        float dist = nmo_euclidean_distance(query, stored, n);
        float den = 1.0f + dist;
        return 1.0f / den;
    }
    }
}

/* lib.rs:2277-2295 compute_similarity (public helper): 0.0 when |a| == 0 */
float nmo_compute_similarity(const float *a, const float *b, uint64_t n) {
    float am = nmo_magnitude(a, n);
    if (am == 0.0f) return 0.0f;
    return nmo_cosine_similarity(a, b, n, am);
}

/* ---- ordering ------------------------------------------------------------------------ */
/* Total order used everywhere: higher score first; -0.0 == +0.0; NaN last; then lower row. */
static inline uint32_t orderable(float s) {
    uint32_t u;
    memcpy(&u, &s, 4);
    if ((u & 0x7fffffffu) > 0x7f800000u) return 0u; /* NaN ranks below -inf */
    if (u == 0x80000000u) u = 0u;                   /* -0.0 ties with +0.0  */
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

typedef struct {
    float score;
    uint64_t row;
} nmo_hit;

static int hit_cmp(const void *pa, const void *pb) {
    const nmo_hit *a = (const nmo_hit *)pa, *b = (const nmo_hit *)pb;
    uint32_t ka = orderable(a->score), kb = orderable(b->score);
    if (ka != kb) return ka > kb ? -1 : 1;
    if (a->row != b->row) return a->row < b->row ? -1 : 1;
    return 0;
}

/* Score every row (lib.rs:2115-2138 search_sequential[_with_metric], rows in mirror order). */
void nmo_score_rows(const float *rows, uint64_t n, uint32_t dim, const float *query, int metric,
                    float *out_scores) {
    float qmag = nmo_magnitude(query, dim);
    for (uint64_t r = 0; r < n; ++r)
        out_scores[r] = nmo_compute_score(query, rows + r * (uint64_t)dim, dim, qmag, metric);
}

/* Faithful search: all N scores -> stable sort desc -> truncate (lib.rs:2013-2034).
 * Returns the number of results written (min(k, n)).  The reference's validation and
 * zero-query short-circuits live in the host wrapper, not here.            */
uint64_t nmo_search(const float *rows, uint64_t n, uint32_t dim, const float *query, uint64_t k,
                    int metric, uint64_t *out_rows, float *out_scores) {
    if (n == 0 || k == 0) return 0;
    nmo_hit *h = (nmo_hit *)malloc(sizeof(nmo_hit) * n);
    if (!h) return 0;
    float qmag = nmo_magnitude(query, dim);
    for (uint64_t r = 0; r < n; ++r) {
        h[r].score = nmo_compute_score(query, rows + r * (uint64_t)dim, dim, qmag, metric);
        h[r].row = r;
    }
    /* qsort is not stable, but hit_cmp is a total order whose tie-break is the original
     * position, so the result equals a stable sort by score. */
    qsort(h, n, sizeof(nmo_hit), hit_cmp);
    uint64_t m = k < n ? k : n;
    for (uint64_t i = 0; i < m; ++i) {
        out_rows[i] = h[i].row;
        out_scores[i] = h[i].score;
    }
    free(h);
    return m;
}

/* ---- multi-threaded form (rayon par_iter analogue, lib.rs:2142-2167) ------------------- */
/* Row-range split over T host threads; each keeps a bounded sorted top-k (cheaper than the
 * reference's full sort — generous to the reference), then the per-thread lists are merged
 * with the same total order.  Result is identical to nmo_search.                         */
typedef struct {
    const float *rows;
    uint64_t lo, hi;
    uint32_t dim;
    const float *query;
    float qmag;
    int metric;
    uint64_t k;
    nmo_hit *top; /* k slots */
    uint64_t cnt;
} nmo_job;

static void topk_insert(nmo_hit *top, uint64_t *cnt, uint64_t k, nmo_hit h) {
    if (*cnt == k) {
        if (hit_cmp(&h, &top[k - 1]) >= 0) return;
    } else {
        (*cnt)++;
    }
    uint64_t i = *cnt - 1;
    while (i > 0 && hit_cmp(&h, &top[i - 1]) < 0) {
        top[i] = top[i - 1];
        --i;
    }
    top[i] = h;
}

static void *job_main(void *arg) {
    nmo_job *j = (nmo_job *)arg;
    for (uint64_t r = j->lo; r < j->hi; ++r) {
        nmo_hit h;
        h.score = nmo_compute_score(j->query, j->rows + r * (uint64_t)j->dim, j->dim, j->qmag,
                                    j->metric);
        h.row = r;
        topk_insert(j->top, &j->cnt, j->k, h);
    }
    return NULL;
}

uint64_t nmo_search_mt(const float *rows, uint64_t n, uint32_t dim, const float *query,
                       uint64_t k, int metric, int threads, uint64_t *out_rows,
                       float *out_scores) {
    if (n == 0 || k == 0) return 0;
    if (threads < 1) threads = 1;
    if ((uint64_t)threads > n) threads = (int)n;
    if (k > n) k = n;
    nmo_job *jobs = (nmo_job *)calloc((size_t)threads, sizeof(nmo_job));
    pthread_t *tid = (pthread_t *)calloc((size_t)threads, sizeof(pthread_t));
    float qmag = nmo_magnitude(query, dim);
    for (int t = 0; t < threads; ++t) {
        jobs[t].rows = rows;
        jobs[t].lo = n * (uint64_t)t / (uint64_t)threads;
        jobs[t].hi = n * (uint64_t)(t + 1) / (uint64_t)threads;
        jobs[t].dim = dim;
        jobs[t].query = query;
        jobs[t].qmag = qmag;
        jobs[t].metric = metric;
        jobs[t].k = k;
        jobs[t].top = (nmo_hit *)malloc(sizeof(nmo_hit) * k);
        jobs[t].cnt = 0;
    }
    for (int t = 1; t < threads; ++t) pthread_create(&tid[t], NULL, job_main, &jobs[t]);
    job_main(&jobs[0]);
    for (int t = 1; t < threads; ++t) pthread_join(tid[t], NULL);
    uint64_t total = 0;
    for (int t = 0; t < threads; ++t) total += jobs[t].cnt;
    nmo_hit *all = (nmo_hit *)malloc(sizeof(nmo_hit) * (total ? total : 1));
    uint64_t p = 0;
    for (int t = 0; t < threads; ++t) {
        memcpy(all + p, jobs[t].top, sizeof(nmo_hit) * jobs[t].cnt);
        p += jobs[t].cnt;
        free(jobs[t].top);
    }
    qsort(all, total, sizeof(nmo_hit), hit_cmp);
    uint64_t m = k < total ? k : total;
    for (uint64_t i = 0; i < m; ++i) {
        out_rows[i] = all[i].row;
        out_scores[i] = all[i].score;
    }
    free(all);
    free(jobs);
    free(tid);
    return m;
}

/* ---- distributed.rs:413-433 merge_top_k ------------------------------------------------ */
/* Shard lists are concatenated in shard order, stably sorted by score desc, truncated.
 * `rows` carry GLOBAL row ids here; with contiguous ascending row ranges per shard the
 * stable order equals (score desc, global row asc).  counts[s] entries per shard, laid out
 * back to back.                                                                            */
uint64_t nmo_merge_top_k(const uint64_t *rows, const float *scores, const uint64_t *counts,
                         uint64_t n_shards, uint64_t k, uint64_t *out_rows, float *out_scores) {
    uint64_t total = 0;
    for (uint64_t s = 0; s < n_shards; ++s) total += counts[s];
    if (total == 0 || k == 0) return 0;
    nmo_hit *all = (nmo_hit *)malloc(sizeof(nmo_hit) * total);
    /* position in the concatenation is the stable-sort tie-break */
    uint64_t *pos_row = (uint64_t *)malloc(sizeof(uint64_t) * total);
    for (uint64_t i = 0; i < total; ++i) {
        all[i].score = scores[i];
        all[i].row = i;
        pos_row[i] = rows[i];
    }
    qsort(all, total, sizeof(nmo_hit), hit_cmp);
    uint64_t m = k < total ? k : total;
    for (uint64_t i = 0; i < m; ++i) {
        out_rows[i] = pos_row[all[i].row];
        out_scores[i] = all[i].score;
    }
    free(all);
    free(pos_row);
    return m;
}

/* ---- synthetic corpus generator -------------------------------------------- */
/* x[r,c] = u24(splitmix64(splitmix64(seed) ^ (r*dim + c))) * 2^-23 - 1, exact in f32, so
 * host and device produce identical bits by construction.  (The seed is hashed once first
 * so that two nearby seeds never alias shifted copies of the same stream.)                */
static inline uint64_t splitmix64(uint64_t x) {
    uint64_t z = x + 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

float nmo_synth_value(uint64_t seed, uint64_t flat_index) {
    uint32_t u24 = (uint32_t)(splitmix64(splitmix64(seed) ^ flat_index) >> 40);
    return (float)u24 * 0x1p-23f - 1.0f;
}

void nmo_fill_synthetic(float *rows, uint64_t n, uint32_t dim, uint64_t seed,
                        uint64_t row_offset) {
    for (uint64_t r = 0; r < n; ++r)
        for (uint32_t c = 0; c < dim; ++c)
            rows[r * (uint64_t)dim + c] = nmo_synth_value(seed, (row_offset + r) * dim + c);
}

typedef struct {
    float *rows;
    uint64_t lo, hi, row_offset, seed;
    uint32_t dim;
} fill_job;

static void *fill_main(void *arg) {
    fill_job *j = (fill_job *)arg;
    nmo_fill_synthetic(j->rows + j->lo * (uint64_t)j->dim, j->hi - j->lo, j->dim, j->seed,
                       j->row_offset + j->lo);
    return NULL;
}

void nmo_fill_synthetic_mt(float *rows, uint64_t n, uint32_t dim, uint64_t seed,
                           uint64_t row_offset, int threads) {
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;
    fill_job jobs[256];
    pthread_t tid[256];
    for (int t = 0; t < threads; ++t) {
        jobs[t].rows = rows;
        jobs[t].lo = n * (uint64_t)t / (uint64_t)threads;
        jobs[t].hi = n * (uint64_t)(t + 1) / (uint64_t)threads;
        jobs[t].row_offset = row_offset;
        jobs[t].seed = seed;
        jobs[t].dim = dim;
        pthread_create(&tid[t], NULL, fill_main, &jobs[t]);
    }
    for (int t = 0; t < threads; ++t) pthread_join(tid[t], NULL);
}
