/* mb_e2e_latency.c — per-call latency of nm_search with host buffers from plain C (no Python /
 * ctypes in the way): what a Rust host binding the C ABI would see.
 *   gcc -O2 scripts/mb_e2e_latency.c -Iinclude -Lneumann_b200 -lneumann_b200 -Wl,-rpath,$PWD/neumann_b200 -o /tmp/mb_e2e && /tmp/mb_e2e */
#define _POSIX_C_SOURCE 200809L
#include <stdio.h>
#include <stdlib.h>
#include <time.h>

#include "neumann_b200.h"

static double now_us(void) {
    struct timespec t;
    clock_gettime(CLOCK_MONOTONIC, &t);
    return t.tv_sec * 1e6 + t.tv_nsec * 1e-3;
}

static int cmp_d(const void *a, const void *b) {
    double x = *(const double *)a, y = *(const double *)b;
    return x < y ? -1 : x > y;
}

static int run(uint64_t n, uint32_t d, uint32_t k, int reps) {
    nm_index *idx = NULL;
    if (nm_index_create(d, NULL, 0, &idx)) return 1;
    if (nm_index_fill_synthetic(idx, n, 0x5EED0001ull, 0)) return 1;
    float *q = malloc(sizeof(float) * d * 16);
    for (uint32_t i = 0; i < d * 16; ++i) q[i] = (float)((i * 2654435761u) >> 8) / 8388608.0f - 1.0f;
    uint64_t rows[1024];
    float scores[1024];
    uint32_t count = 0;
    double *ts = malloc(sizeof(double) * reps);
    for (int i = 0; i < 20; ++i)
        if (nm_search(idx, q + (i % 16) * d, 1, k, NM_COSINE, rows, scores, &count)) return 1;
    for (int i = 0; i < reps; ++i) {
        double t0 = now_us();
        if (nm_search(idx, q + (i % 16) * d, 1, k, NM_COSINE, rows, scores, &count)) return 1;
        ts[i] = now_us() - t0;
    }
    double sum = 0;
    for (int i = 0; i < reps; ++i) sum += ts[i];
    qsort(ts, reps, sizeof(double), cmp_d);
    printf("C ABI %llux%u top-%u: mean %8.2f us  p50 %8.2f  p10 %8.2f  p90 %8.2f\n", (unsigned long long)n, d, k,
           sum / reps, ts[reps / 2], ts[reps / 10], ts[reps * 9 / 10]);
    nm_index_destroy(idx);
    free(q);
    free(ts);
    return 0;
}

int main(void) {
    if (run(10000, 128, 5, 3000)) { fprintf(stderr, "%s\n", nm_last_error()); return 1; }
    if (run(1000000, 768, 10, 500)) { fprintf(stderr, "%s\n", nm_last_error()); return 1; }
    return 0;
}
