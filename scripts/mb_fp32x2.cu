// Microbenchmark: scalar vs packed (f32x2) non-fused FP32 throughput on sm_100a.
// Each thread runs NCH independent dependent-chains of (sub, mul, add) like the L2 scan.
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

template <int NCH>
__global__ void k_scalar(float *out, const float *in, int iters) {
    float acc[NCH], x[NCH];
#pragma unroll
    for (int j = 0; j < NCH; ++j) { acc[j] = 0.f; x[j] = in[threadIdx.x + j]; }
    float q = in[0];
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < NCH; ++j) {
            float d = __fsub_rn(q, x[j]);
            acc[j] = __fadd_rn(acc[j], __fmul_rn(d, d));
        }
        q = __fadd_rn(q, 1e-7f);
    }
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < NCH; ++j) s += acc[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__device__ __forceinline__ uint64_t pk(float a, float b) {
    return ((uint64_t)__float_as_uint(b) << 32) | __float_as_uint(a);
}
template <int NCH>  // NCH packed chains = 2*NCH scalar chains
__global__ void k_packed(float *out, const float *in, int iters) {
    uint64_t acc[NCH], x[NCH];
#pragma unroll
    for (int j = 0; j < NCH; ++j) { acc[j] = 0ull; x[j] = pk(in[threadIdx.x + j], in[threadIdx.x + j + 1]); }
    uint64_t q = pk(in[0], in[1]);
    uint64_t eps = pk(1e-7f, 1e-7f);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < NCH; ++j) {
            uint64_t d, m;
            asm volatile("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(q), "l"(x[j]));
            asm volatile("mul.rn.f32x2 %0, %1, %1;" : "=l"(m) : "l"(d));
            asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(acc[j]) : "l"(m));
        }
        asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(q) : "l"(eps));
    }
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < NCH; ++j) s += __uint_as_float((uint32_t)acc[j]) + __uint_as_float((uint32_t)(acc[j] >> 32));
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
    float *in, *out;
    cudaMalloc(&in, 4096 * 4); cudaMalloc(&out, 148 * 8 * 1024 * 4);
    cudaMemset(in, 0, 4096 * 4);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    int iters = 20000;
    for (int rep = 0; rep < 2; ++rep) {
        for (int threads : {256, 512, 1024}) {
            int grid = 148 * (1024 / threads);
            float ms;
            cudaEventRecord(a); k_scalar<16><<<grid, threads>>>(out, in, iters); cudaEventRecord(b); cudaEventSynchronize(b);
            cudaEventElapsedTime(&ms, a, b);
            double elems = (double)grid * threads * iters * 16;   // scalar (sub,mul,add) triples
            if (rep) printf("scalar  thr=%4d: %.3f ms  %.2f Gtriples/s  (%.2f T lane-ops/s)\n", threads, ms, elems / ms / 1e6, 3 * elems / ms / 1e9);
            cudaEventRecord(a); k_packed<8><<<grid, threads>>>(out, in, iters); cudaEventRecord(b); cudaEventSynchronize(b);
            cudaEventElapsedTime(&ms, a, b);
            elems = (double)grid * threads * iters * 16;          // 8 packed chains = 16 scalar triples
            if (rep) printf("packed  thr=%4d: %.3f ms  %.2f Gtriples/s  (%.2f T lane-ops/s)\n", threads, ms, elems / ms / 1e6, 3 * elems / ms / 1e9);
        }
    }
    printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
