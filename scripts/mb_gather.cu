// Microbenchmark: exact re-score access pattern.  R random rows of D floats out of N rows,
// one thread per row walking its row in 128-byte blocks (as tc_score_row does), in random
// order vs sorted by row.  Answers: is the scattered-row gather bound by latency, by DRAM, or
// by address translation (then sorting by row helps)?
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o scripts/mb_gather scripts/mb_gather.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>
#include <stdint.h>

__global__ void gather_kernel(const float *rows, const uint32_t *idx, uint32_t n, uint32_t dim,
                              float *out) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float4 *x = reinterpret_cast<const float4 *>(rows + (size_t)idx[i] * dim);
        float s = 0.f;
        for (uint32_t b = 0; b < dim / 4; b += 8) {
            float4 v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = __ldg(x + b + j);
#pragma unroll
            for (int j = 0; j < 8; ++j) s += v[j].x + v[j].y + v[j].z + v[j].w;
        }
        out[i] = s;
    }
}

int main(int argc, char **argv) {
    size_t N = argc > 1 ? atoll(argv[1]) : 10000000;
    uint32_t D = argc > 2 ? atoi(argv[2]) : 1536;
    uint32_t R = argc > 3 ? atoi(argv[3]) : 458000;
    float *rows;
    cudaMalloc(&rows, N * D * 4);
    cudaMemset(rows, 0, N * D * 4);
    std::vector<uint32_t> h(R);
    uint64_t s = 88172645463325252ull;
    for (auto &v : h) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; v = (uint32_t)(s % N); }
    uint32_t *d_idx; float *out;
    cudaMalloc(&d_idx, R * 4); cudaMalloc(&out, R * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int mode = 0; mode < 3; ++mode) {
        std::vector<uint32_t> v = h;
        if (mode == 1) std::sort(v.begin(), v.end());
        if (mode == 2) {  // sorted within chunks of 64K entries (per-query lists nearly sorted)
            for (size_t o = 0; o < v.size(); o += 1792) std::sort(v.begin() + o, v.begin() + std::min(v.size(), o + 1792));
        }
        cudaMemcpy(d_idx, v.data(), R * 4, cudaMemcpyHostToDevice);
        for (int threads : {128, 256}) for (int blocks : {148 * 2, 148 * 8}) {
            float best = 1e9;
            for (int it = 0; it < 4; ++it) {
                cudaEventRecord(e0);
                gather_kernel<<<blocks, threads>>>(rows, d_idx, R, D, out);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1); best = std::min(best, ms);
            }
            printf("mode %d (%s) blocks %d x %d: %.3f ms  %.2f TB/s\n", mode,
                   mode == 0 ? "random" : mode == 1 ? "sorted" : "per-list sorted", blocks, threads, best,
                   (double)R * D * 4 / best / 1e9);
        }
    }
    return 0;
}
