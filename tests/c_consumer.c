/* c_consumer.c — a plain C11 client of include/neumann_b200.h: proves the drop-in boundary
 * without Python/ctypes in the way (the Rust `-sys` shim of INTEGRATION.md binds exactly these
 * symbols).  Built and run by tests/test_c_consumer.py:
 *   gcc -std=c11 -Wall -Wextra -Werror -pedantic tests/c_consumer.c -Iinclude -Lneumann_b200 -lneumann_b200
 * Exit codes: 0 = everything checked on a GPU; 10 = no CUDA device, and the library said so the
 * documented way (NM_ERR_STORAGE + message, nothing else touched); anything else = failure. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "neumann_b200.h"
#include "neumann_b200_engine.h"

#define CHECK(cond)                                                                  \
    do {                                                                             \
        if (!(cond)) {                                                               \
            fprintf(stderr, "c_consumer: %s:%d: %s failed (last error: %s)\n", __FILE__, __LINE__, \
                    #cond, nm_last_error());                                         \
            return 1;                                                                \
        }                                                                            \
    } while (0)

int main(void) {
    CHECK(nm_abi_version() == NM_ABI_VERSION);
    CHECK(sizeof(nm_filter_op) == 24);
    nm_index *idx = NULL;
    if (nm_device_count() == 0) {
        int rc = nm_index_create(4, NULL, 0, &idx);
        CHECK(rc == NM_ERR_STORAGE && idx == NULL);
        CHECK(strstr(nm_last_error(), "no CUDA device") != NULL);
        CHECK(nm_index_create(0, NULL, 0, &idx) == NM_ERR_EMPTY_VECTOR);
        CHECK(nm_index_rows(NULL) == 0 && nm_index_dim(NULL) == 0);
        nm_index_destroy(NULL);
        puts("c_consumer: no CUDA device, failed loudly as documented");
        return 10;
    }
    /* 6 rows x 4 dims; the reference's own KAT shapes (vector_engine/src/lib.rs:4183-4239) */
    enum { N = 6, D = 4, K = 3 };
    const float rows[N * D] = {
        1.f, 0.f, 0.f, 0.f, /* 0: identical to the query      cos  1          */
        0.f, 1.f, 0.f, 0.f, /* 1: orthogonal                   cos  0          */
        -1.f, 0.f, 0.f, 0.f, /* 2: opposite                     cos -1          */
        1.f, 1.f, 0.f, 0.f, /* 3: 45 degrees                   cos  sqrt(2)/2  */
        0.f, 0.f, 0.f, 0.f, /* 4: zero row                     cos  0 exactly  */
        2.f, 0.f, 0.f, 0.f, /* 5: same direction, longer       cos  1 (tie with 0 -> lower row first) */
    };
    const float q[D] = {1.f, 0.f, 0.f, 0.f};
    CHECK(nm_index_create(D, NULL, 0, &idx) == NM_OK && idx != NULL);
    CHECK(nm_index_load(idx, rows, 4) == NM_OK);
    CHECK(nm_index_append(idx, rows + 4 * D, 2) == NM_OK);
    CHECK(nm_index_rows(idx) == N && nm_index_dim(idx) == D && nm_index_device_count(idx) == 1);
    uint64_t out_rows[K];
    float out_scores[K];
    uint32_t count = 0;
    CHECK(nm_search(idx, q, 1, K, NM_COSINE, out_rows, out_scores, &count) == NM_OK);
    CHECK(count == K && out_rows[0] == 0 && out_rows[1] == 5 && out_rows[2] == 3);
    CHECK(out_scores[0] == 1.0f && out_scores[1] == 1.0f);
    CHECK(fabsf(out_scores[2] - 0.70710678f) < 1e-6f);
    /* euclidean: score = 1 / (1 + d) (lib.rs:2231-2246) */
    CHECK(nm_search(idx, q, 1, 2, NM_EUCLIDEAN, out_rows, out_scores, &count) == NM_OK);
    CHECK(count == 2 && out_rows[0] == 0 && out_scores[0] == 1.0f && out_scores[1] == 0.5f);
    /* validation mirrors the reference */
    CHECK(nm_search(idx, q, 1, 0, NM_COSINE, out_rows, out_scores, &count) == NM_ERR_INVALID_TOP_K);
    CHECK(nm_search(idx, q, 0, K, NM_COSINE, out_rows, out_scores, &count) == NM_ERR_EMPTY_VECTOR);
    CHECK(nm_search(idx, q, 1, K, 7, out_rows, out_scores, &count) == NM_ERR_INVALID_ARGUMENT);
    /* pre-filtered search: host bitmask, then a device-evaluated filter over a metadata column */
    const uint64_t mask = (1u << 2) | (1u << 3) | (1u << 4);
    CHECK(nm_search_masked(idx, q, 1, K, NM_COSINE, &mask, out_rows, out_scores, &count) == NM_OK);
    CHECK(count == 3 && out_rows[0] == 3 && out_rows[1] == 4 && out_rows[2] == 2);
    const uint8_t tags[N] = {NM_V_INT, NM_V_INT, NM_V_INT, NM_V_FLOAT, NM_V_MISSING, NM_V_INT};
    const double two_and_a_half = 2.5;
    uint64_t vals[N] = {10, 20, 30, 0, 0, 40};
    memcpy(&vals[3], &two_and_a_half, 8);
    CHECK(nm_index_column_set(idx, 1, 0, N, tags, vals) == NM_OK);
    nm_filter_op prog[1];
    memset(prog, 0, sizeof(prog));
    prog[0].kind = NM_F_CMP;
    prog[0].cmp = NM_C_LT;
    prog[0].lit_tag = NM_V_INT;
    prog[0].column = 1;
    prog[0].lit = 25; /* rows 0, 1 (ints) and 3 (2.5 as float); row 4 has no such field */
    uint64_t eligible = 0, fmask = 0;
    CHECK(nm_index_filter_mask(idx, prog, 1, NULL, 0, &fmask, &eligible) == NM_OK);
    CHECK(eligible == 3 && fmask == ((1u << 0) | (1u << 1) | (1u << 3)));
    CHECK(nm_search_filtered(idx, q, 1, K, NM_COSINE, prog, 1, NULL, 0, out_rows, out_scores, &count) == NM_OK);
    CHECK(count == 3 && out_rows[0] == 0 && out_rows[1] == 3 && out_rows[2] == 1);
    /* mutations */
    uint64_t moved = 99;
    CHECK(nm_index_swap_remove(idx, 0, &moved) == NM_OK && moved == 5 && nm_index_rows(idx) == 5);
    CHECK(nm_search(idx, q, 1, 1, NM_COSINE, out_rows, out_scores, &count) == NM_OK);
    CHECK(count == 1 && out_rows[0] == 0 && out_scores[0] == 1.0f); /* row 5 moved into slot 0 */
    float back[D];
    CHECK(nm_index_get_row(idx, 0, back) == NM_OK && back[0] == 2.f);
    nm_shard_info info;
    CHECK(nm_index_shard_info(idx, 0, &info) == NM_OK && info.rows == 5 && info.row_base == 0);
    nm_stats st;
    CHECK(nm_index_stats(idx, &st) == NM_OK && st.searches >= 5);
    CHECK(nm_index_clear(idx) == NM_OK && nm_index_rows(idx) == 0);
    nm_index_destroy(idx);
    puts("c_consumer: ok");
    return 0;
}
