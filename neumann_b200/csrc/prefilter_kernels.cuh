// prefilter_kernels.cuh — SURVEY 8f row 4: quantised pre-filter with EXACT re-score.
//
// Idea (precedent in the reference: ScalarQuantizedVector, tensor_store/src/hnsw.rs:308-356):
// keep an int8 copy of the corpus (1 byte per element + 16 bytes per row), scan THAT with
// dp4a (4x fewer HBM bytes than the f32 scan), and turn every approximate dot product into a
// rigorous interval [lo, hi] that contains the reference-arithmetic f32 dot product of the
// row.  Because every later step of compute_score (divide by |q||x|, ...) is a monotone,
// correctly rounded f32 operation, the interval maps to a score interval [lb, ub] with the
// reference's own operations.  With tau = the k-th largest lb, every row of the exact top-k
// satisfies ub >= tau; those few candidates are re-scored with the exact lane-tree
// arithmetic from the f32 mirror and selected exactly as the f32 scan does.  The result is
// bit-identical to nm_search without the pre-filter (tests/test_gpu_prefilter.py); if the
// candidate list overflows, or anything is non-finite, the caller falls back to the f32 scan.
//
// Error model (all data finite, quantisation scales normal f32 numbers; rows that violate
// either are flagged and always re-scored, a query that does falls back to the f32 scan).  Row r: x_i = s_r (xt_i + d_i), xt_i in [-127,127] integer,
// |d_i| <= 0.5001 (rint + the rounding of the f32 division x_i / s_r).  Query likewise with
// s_q, qt_i, |e_i| <= 0.5001.  With I = sum qt_i xt_i (exact in int32) and D* the real dot:
//   |D* - s_q s_r I| <= s_q s_r B_r,   B_r = 0.5001 (Q1 + X1_r) + 0.2502 dim,
//   Q1 = sum |qt_i|, X1_r = sum |xt_i|.
// The reference's f32 lane tree D^ obeys |D^ - D*| <= g S with S = sum |q_i x_i| <=
// s_q s_r (127.51 X1_r + B_r) and g = 2 (dim + 16) 2^-24 (any summation order; generous).
// lo/hi are evaluated in double and rounded outward to f32.
#pragma once
#include "scan_kernels.cuh"

namespace nm {

#ifdef __CUDACC__

// exact reference arithmetic on one row straight from global memory (re-score path)
__device__ __forceinline__ float exact_score_row(const float *__restrict__ q,
                                                 const float *__restrict__ x, uint32_t dim,
                                                 float qmag, int metric) {
    if (metric == kEuclidean) {
        float s = 0.0f;
        for (uint32_t i = 0; i < dim; ++i) {
            float d = __fsub_rn(q[i], x[i]);
            s = __fadd_rn(s, __fmul_rn(d, d));
        }
        return __fdiv_rn(1.0f, __fadd_rn(1.0f, __fsqrt_rn(s)));
    }
    float dl[8] = {0, 0, 0, 0, 0, 0, 0, 0}, sl[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const uint32_t chunks = dim / 8u;
    for (uint32_t c = 0; c < chunks; ++c) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float xv = x[c * 8u + j];
            dl[j] = __fadd_rn(dl[j], __fmul_rn(q[c * 8u + j], xv));
            if (metric == kCosine) sl[j] = __fadd_rn(sl[j], __fmul_rn(xv, xv));
        }
    }
    float dot = fold_lanes(dl), ssq = fold_lanes(sl);
    for (uint32_t i = chunks * 8u; i < dim; ++i) {
        float xv = x[i];
        dot = __fadd_rn(dot, __fmul_rn(q[i], xv));
        if (metric == kCosine) ssq = __fadd_rn(ssq, __fmul_rn(xv, xv));
    }
    if (metric == kDot) return dot;
    float rmag = __fsqrt_rn(ssq);
    return (qmag == 0.0f || rmag == 0.0f) ? 0.0f : __fdiv_rn(dot, __fmul_rn(qmag, rmag));
}

// ---------------------------------------------------------------------------------------
// quantise rows [first, first + n): one warp per row
// ---------------------------------------------------------------------------------------
// norms[r] = (>= ||xt||_2, >= ||x / s_r - xt||_2): the Cauchy-Schwarz inputs of the batch
// pre-filter's error bound (tc_prefilter_kernels.cuh); +inf for rows the analysis skips.
__global__ void quantize_rows_kernel(const float *__restrict__ rows, uint32_t pitch, uint32_t dim,
                                     uint64_t first, uint64_t n, int8_t *q8, uint32_t pitch8,
                                     RowMeta *meta, float2 *norms, uint32_t *nonfinite_flag) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint64_t warp = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
    const uint64_t n_warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t r = first + warp; r < first + n; r += n_warps) {
        const float *x = rows + r * pitch;
        float mx = 0.0f;
        bool bad = false;
        for (uint32_t i = lane; i < dim; i += 32u) {
            float v = x[i];
            bad |= !(fabsf(v) <= 3.4028234e38f);  // inf or NaN
            mx = fmaxf(mx, fabsf(v));
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            bad |= __shfl_xor_sync(0xffffffffu, (int)bad, o) != 0;
        }
        float scale = bad ? 0.0f : __fdiv_rn(mx, 127.0f);
        // a denormal scale no longer satisfies |x/scale - rint| <= 0.5001: treat the row as
        // "unknown" (always a candidate, exact re-score decides)
        if (mx > 0.0f && scale < 1.17549435e-38f) {
            bad = true;
            scale = 0.0f;
        }
        uint32_t x1 = 0, x2 = 0;  // sum |xt|, sum xt^2 (exact: 127^2 dim < 2^32)
        float dd = 0.0f;          // sum (x / s - xt)^2, every step rounded up
        int8_t *out = q8 + r * pitch8;
        for (uint32_t w = lane; w * 4u < pitch8; w += 32u) {
            uint32_t packed = 0;
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const uint32_t i = w * 4u + b;
                int v = 0;
                if (i < dim && scale > 0.0f) {
                    const float q = __fdiv_rn(x[i], scale);
                    v = __float2int_rn(q);
                    v = max(-127, min(127, v));
                    const float d = __fsub_rn(q, (float)v);  // exact (Sterbenz / small integers)
                    dd = __fmaf_ru(d, d, dd);
                }
                x1 += (uint32_t)abs(v);
                x2 += (uint32_t)(v * v);
                packed |= ((uint32_t)(uint8_t)(int8_t)v) << (8 * b);
            }
            reinterpret_cast<uint32_t *>(out)[w] = packed;
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            x1 += __shfl_xor_sync(0xffffffffu, x1, o);
            x2 += __shfl_xor_sync(0xffffffffu, x2, o);
            dd = __fadd_ru(dd, __shfl_xor_sync(0xffffffffu, dd, o));
        }
        // reference-arithmetic magnitude: lanes 0..7 are the f32x8 lanes
        float acc = 0.0f;
        const uint32_t chunks = dim / 8u;
        if (lane < 8u)
            for (uint32_t c = 0; c < chunks; ++c) {
                float v = x[c * 8u + lane];
                acc = __fadd_rn(acc, __fmul_rn(v, v));
            }
        float ssq = 0.0f;
#pragma unroll
        for (int j = 0; j < 8; ++j) ssq = __fadd_rn(ssq, __shfl_sync(0xffffffffu, acc, j));
        if (lane == 0) {
            for (uint32_t i = chunks * 8u; i < dim; ++i) ssq = __fadd_rn(ssq, __fmul_rn(x[i], x[i]));
            RowMeta m;
            m.scale = scale;
            m.x1 = x1;
            m.rmag = __fsqrt_rn(ssq);
            m.flags = bad ? 1u : 0u;
            meta[r] = m;
            if (norms) {
                // the f32 quotient x / s is off by <= 127.01 * 2^-24 per element
                const float dn = __fadd_ru(__fsqrt_ru(dd), __fmul_ru(7.7e-6f, __fsqrt_ru((float)dim)));
                norms[r] = bad ? make_float2(INFINITY, INFINITY)
                               : make_float2(__fsqrt_ru(__uint2float_ru(x2)), dn);
            }
            if (bad) atomicOr(nonfinite_flag, 1u);
        }
    }
}

#ifdef __CUDACC__
// score of the reference's Euclidean metric from the f32 chain sum (lib.rs:2244, 2249-2253)
__device__ __forceinline__ float tc_l2_score(float s) {
    return __fdiv_rn(1.0f, __fadd_rn(1.0f, __fsqrt_rn(s)));
}
constexpr double kTcU = 5.9604644775390625e-08;  // 2^-24

// bounds on the real |x|^2 of a row from its reference-arithmetic magnitude: rmag =
// fl(sqrt(lane tree)), every term non-negative, so the tree is within (dim/8 + 16) u relative
// and the sqrt / re-squaring add ~2u; 1e-37 covers products that underflowed.
__device__ __forceinline__ void tc_row_sq_bounds(const RowMeta &m, uint32_t dim, double &a_lo,
                                                 double &a_hi) {
    const double a = (double)m.rmag * (double)m.rmag;
    const double rel = ((double)(dim / 8u) + 24.0) * kTcU * 1.01;
    a_lo = a * (1.0 - rel) - 1e-37;
    a_hi = a * (1.0 + rel) + 1e-37;
    if (a_lo < 0.0) a_lo = 0.0;
}

#endif  // __CUDACC__

// ---------------------------------------------------------------------------------------
// pre-filter scan
// ---------------------------------------------------------------------------------------
struct PrefilterParams {
    const float *query;       // [dim] f32
    const RowMeta *meta;      // [rows]
    uint64_t *cand;           // [grid, k] per-CTA best LOWER-bound keys
    uint32_t *ctl;            // [0] ticket [1] row-block cursor [2] kept count [3] status
                              // [4] tau_ord (k-th best lower bound, score part) [5] exact count
    KeptEntry *kept;          // [kKeptCap]
    uint32_t n_rows;
    uint32_t dim;
    uint32_t k;
    uint32_t n_stages;
    uint32_t q_words;         // int8 query words in smem (multiple of 32)
    int metric;               // kCosine, kDot or kEuclidean
};

// status bits
constexpr uint32_t kPfOverflow = 1u;   // kept list overflowed
constexpr uint32_t kPfNonFinite = 2u;  // query not finite

__device__ __forceinline__ void tma_load_2d_nohint(void *dst, const CUtensorMap *tmap, int32_t x,
                                                   int32_t y, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%2, %3}], [%4];" ::"r"(smem_u32(dst)),
        "l"(tmap), "r"(x), "r"(y), "r"(smem_u32(bar))
        : "memory");
}

constexpr uint32_t kKeptStage = 1024;  // kept entries staged per CTA between flushes

// All consumer threads, uniform call: move the staged kept entries to the global list.
__device__ __forceinline__ void flush_kept(const KeptEntry *kept_s, uint32_t *kept_cnt_s,
                                           KeptEntry *kept_g, uint32_t *ctl, uint32_t t) {
    const uint32_t n = kept_cnt_s[0];
    if (t == 0) kept_cnt_s[1] = atomicAdd(ctl + 2, n);
    consumer_sync();
    const uint32_t base = kept_cnt_s[1];
    for (uint32_t i = t; i < n; i += kRowsPerBlock) {
        if (base + i < kKeptCap) kept_g[base + i] = kept_s[i];
    }
    if (t == 0) {
        if (base + n > kKeptCap) atomicOr(ctl + 3, kPfOverflow);
        kept_cnt_s[0] = 0u;
    }
    consumer_sync();
}

__global__ void __launch_bounds__(kScanThreads, 1)
prefilter_scan_kernel(const __grid_constant__ CUtensorMap tmap8, const PrefilterParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t *stages = smem;
    uint64_t *cand_buf = reinterpret_cast<uint64_t *>(stages + p.n_stages * kStageBytes);
    uint32_t *q8_s = reinterpret_cast<uint32_t *>(cand_buf + kCandCap);
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(q8_s + p.q_words);
    uint64_t *empty_bar = full_bar + kMaxStages;
    uint64_t *thr_s = empty_bar + kMaxStages;
    uint32_t *cnt_s = reinterpret_cast<uint32_t *>(thr_s + 1);
    uint32_t *ticket_s = cnt_s + 1;
    uint32_t *rb_ring = cnt_s + 2;                         // kMaxStages
    float *red_f = reinterpret_cast<float *>(rb_ring + kMaxStages);  // 8 floats
    uint32_t *red_u = reinterpret_cast<uint32_t *>(red_f + 8);       // 8 uints
    float *qs_s = reinterpret_cast<float *>(red_u + 8);              // [0] s_q [1] qmag [2] Q1
    uint32_t *kept_cnt_s = reinterpret_cast<uint32_t *>(qs_s + 4);   // [0] count [1] flush base
    KeptEntry *kept_s = reinterpret_cast<KeptEntry *>(kept_cnt_s + 2);  // kKeptStage entries
    double *red_d = reinterpret_cast<double *>(kept_s + kKeptStage);    // 8 doubles

    const uint32_t tid = threadIdx.x;
    const uint32_t warp = tid >> 5;
    const uint32_t n_stages = p.n_stages;
    const uint32_t n_rb = (p.n_rows + kRowsPerBlock - 1) / kRowsPerBlock;
    const uint32_t n_kc = (p.dim + 127u) / 128u;  // 128 int8 per row per stage

    if (tid == 0) {
        for (uint32_t s = 0; s < n_stages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], kConsumerWarps);
        }
        *thr_s = 0ull;
        *cnt_s = 0u;
        kept_cnt_s[0] = 0u;
        fence_mbar_init();
    }
    __syncthreads();

    if (warp == kConsumerWarps) {
        if (tid == kRowsPerBlock) {
            uint32_t stage = 0, phase = 0;
            uint32_t rb = blockIdx.x;
            for (;;) {
                if (rb >= n_rb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1u);
                    rb_ring[stage] = 0xffffffffu;
                    mbar_arrive(&full_bar[stage]);
                    break;
                }
                const uint32_t next = atomicAdd(p.ctl + 1, 1u) + gridDim.x;
                for (uint32_t kc = 0; kc < n_kc; ++kc) {
                    mbar_wait(&empty_bar[stage], phase ^ 1u);
                    if (kc == 0) rb_ring[stage] = rb;
                    mbar_arrive_expect_tx(&full_bar[stage], kStageBytes);
                    tma_load_2d_nohint(stages + stage * kStageBytes, &tmap8, (int32_t)(kc * 128u),
                                       (int32_t)(rb * kRowsPerBlock), &full_bar[stage]);
                    if (++stage == n_stages) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
                rb = next;
            }
        }
        return;
    }

    // ---- consumers ----
    const uint32_t t = tid, lane = t & 31u;
    // query statistics: max |q|, finiteness
    float mx = 0.0f;
    bool bad = false;
    double qsq = 0.0;  // real |q|^2 (Euclidean: d^2 = |x|^2 + |q|^2 - 2 q.x)
    for (uint32_t i = t; i < p.dim; i += kRowsPerBlock) {
        float v = __ldg(p.query + i);
        bad |= !(fabsf(v) <= 3.4028234e38f);
        mx = fmaxf(mx, fabsf(v));
        qsq += (double)v * (double)v;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        bad |= __shfl_xor_sync(0xffffffffu, (int)bad, o) != 0;
        qsq += __shfl_xor_sync(0xffffffffu, qsq, o);
    }
    if (lane == 0) {
        red_f[warp] = mx;
        red_u[warp] = bad ? 1u : 0u;
        red_d[warp] = qsq;
    }
    consumer_sync();
    mx = 0.0f;
    bad = false;
    qsq = 0.0;
#pragma unroll
    for (int w = 0; w < kConsumerWarps; ++w) {
        mx = fmaxf(mx, red_f[w]);
        bad |= red_u[w] != 0u;
        qsq += red_d[w];
    }
    double c_lo = qsq * (1.0 - 1e-12) - 1e-40, c_hi = qsq * (1.0 + 1e-12) + 1e-40;
    if (c_lo < 0.0) c_lo = 0.0;
    const float s_q = __fdiv_rn(mx, 127.0f);
    if (mx > 0.0f && s_q < 1.17549435e-38f) bad = true;  // denormal scale: let the f32 scan decide
    consumer_sync();
    // quantise the query into packed int8 words (zero padded), Q1 = sum |qt|
    uint32_t q1 = 0;
    for (uint32_t w = t; w < p.q_words; w += kRowsPerBlock) {
        uint32_t packed = 0;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const uint32_t i = w * 4u + b;
            int v = 0;
            if (i < p.dim && s_q > 0.0f && !bad) {
                v = __float2int_rn(__fdiv_rn(__ldg(p.query + i), s_q));
                v = max(-127, min(127, v));
            }
            q1 += (uint32_t)abs(v);
            packed |= ((uint32_t)(uint8_t)(int8_t)v) << (8 * b);
        }
        q8_s[w] = packed;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) q1 += __shfl_xor_sync(0xffffffffu, q1, o);
    if (lane == 0) red_u[warp] = q1;
    // |q| with the reference lane tree (warp 0)
    if (warp == 0) {
        float acc = 0.0f;
        const uint32_t chunks = p.dim / 8u;
        if (lane < 8u)
            for (uint32_t c = 0; c < chunks; ++c) {
                float v = __ldg(p.query + c * 8u + lane);
                acc = __fadd_rn(acc, __fmul_rn(v, v));
            }
        float r = 0.0f;
#pragma unroll
        for (int j = 0; j < 8; ++j) r = __fadd_rn(r, __shfl_sync(0xffffffffu, acc, j));
        if (lane == 0) {
            for (uint32_t i = chunks * 8u; i < p.dim; ++i) {
                float v = __ldg(p.query + i);
                r = __fadd_rn(r, __fmul_rn(v, v));
            }
            qs_s[1] = __fsqrt_rn(r);
        }
    }
    consumer_sync();
    q1 = 0;
#pragma unroll
    for (int w = 0; w < kConsumerWarps; ++w) q1 += red_u[w];
    const float qmag = qs_s[1];
    if (bad && t == 0) atomicOr(p.ctl + 3, kPfNonFinite);

    TopKState st;
    st.buf = cand_buf;
    st.cnt_smem = cnt_s;
    st.thr_smem = thr_s;
    st.count = 0;
    st.k = p.k;
    st.cap = 512u;
    while (st.cap < p.k + (uint32_t)kRowsPerBlock) st.cap <<= 1;

    const double g = 2.0 * ((double)p.dim + 16.0) * 5.9604644775390625e-08;  // 2 (dim+16) 2^-24
    const uint32_t swz = t & 7u;
    uint32_t stage = 0, phase = 0;
    for (;;) {
        mbar_wait(&full_bar[stage], phase);
        const uint32_t rb = rb_ring[stage];
        if (rb == 0xffffffffu) break;
        const uint32_t row = rb * kRowsPerBlock + t;
        const uint32_t g_thr = *reinterpret_cast<volatile uint32_t *>(p.ctl + 4);
        RowMeta m;
        m.scale = 0.0f;
        m.x1 = 0;
        m.rmag = 0.0f;
        m.flags = 0;
        if (row < p.n_rows) {
            const float4 raw = __ldg(reinterpret_cast<const float4 *>(p.meta + row));
            m.scale = raw.x;
            m.x1 = __float_as_uint(raw.y);
            m.rmag = raw.z;
            m.flags = __float_as_uint(raw.w);
        }
        int acc = 0;
        for (uint32_t kc = 0; kc < n_kc; ++kc) {
            mbar_wait(&full_bar[stage], phase);
            const uint8_t *srow = stages + stage * kStageBytes + t * 128u;
            const uint4 *qv = reinterpret_cast<const uint4 *>(q8_s) + kc * 8u;
#pragma unroll
            for (uint32_t u = 0; u < 8; ++u) {
                const uint4 x = *reinterpret_cast<const uint4 *>(srow + ((u ^ swz) << 4));
                const uint4 q = qv[u];
                acc = __dp4a((int)x.x, (int)q.x, acc);
                acc = __dp4a((int)x.y, (int)q.y, acc);
                acc = __dp4a((int)x.z, (int)q.z, acc);
                acc = __dp4a((int)x.w, (int)q.w, acc);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty_bar[stage]);
            if (++stage == n_stages) {
                stage = 0;
                phase ^= 1u;
            }
        }
        // interval for the reference dot product, then for the score
        uint64_t lb_key = 0ull;
        uint32_t ub_ord = 0u;
        if (row < p.n_rows) {
            const double ss = (double)s_q * (double)m.scale;
            const double B = 0.5001 * ((double)q1 + (double)m.x1) + 0.2502 * (double)p.dim;
            const double S = 127.51 * (double)m.x1 + B;
            const double E = (ss * (B + g * S)) * 1.000001 + 1e-37;
            const double Dt = ss * (double)acc;
            float lo = __double2float_rd(Dt - E), hi = __double2float_ru(Dt + E);
            // S bounds every partial sum of the reference's f32 arithmetic: below 1e37 nothing
            // can overflow there, so the interval analysis applies; otherwise "unknown"
            bool wild = !(ss * S < 1e37) || !(fabs(Dt) + E < 1e37);
            float lb, ub;
            if (p.metric == kEuclidean) {
                // the same interval as the tensor-core batch pre-filter (tc_interval): the real dot
                // lies within ss*B of ss*I, the real d^2 = |x|^2 + |q|^2 - 2 dot, and the f32 chain
                // of non-negative terms is within (dim + 4) u relative of the real d^2; the score
                // 1 / (1 + sqrt(.)) is monotone and correctly rounded
                const double Ed = ss * B * 1.000001 + 1e-37;
                double a_lo, a_hi;
                tc_row_sq_bounds(m, p.dim, a_lo, a_hi);
                const double gc = ((double)p.dim + 4.0) * kTcU * 1.01;
                double dlo = a_lo + c_lo - 2.0 * (Dt + Ed);
                double dhi = a_hi + c_hi - 2.0 * (Dt - Ed);
                dlo = dlo * (1.0 - gc) * (1.0 - 1e-12) - 1e-36;
                dhi = dhi * (1.0 + gc) * (1.0 + 1e-12) + 1e-36;
                wild = !(dhi < 1e37) || !(a_hi < 1e37) || !(c_hi < 1e37);
                ub = tc_l2_score((dlo > 0.0) ? __double2float_rd(dlo) : 0.0f);
                lb = tc_l2_score((dhi > 0.0) ? __double2float_ru(dhi) : 0.0f);
            } else if (p.metric == kCosine) {
                if (qmag == 0.0f || m.rmag == 0.0f) {
                    lb = ub = 0.0f;
                } else {
                    const float den = __fmul_rn(qmag, m.rmag);
                    lb = __fdiv_rn(lo, den);
                    ub = __fdiv_rn(hi, den);
                    if (!(den > 0.0f) || !(den < 3.0e38f)) {  // |q||x| under/overflowed: unknown
                        lb = -INFINITY;
                        ub = INFINITY;
                    }
                }
            } else {
                lb = lo;
                ub = hi;
            }
            if (wild || (m.flags & 1u)) {
                lb = -INFINITY;
                ub = INFINITY;
            }
            lb_key = make_key(__float_as_uint(lb), row);
            ub_ord = score_to_ord(__float_as_uint(ub));
        }
        // keep every row whose upper bound reaches the running k-th best lower bound: this
        // CTA's own, or the best any CTA has published so far (each is <= the final tau)
        const uint32_t thr_ord = max((uint32_t)(*st.thr_smem >> 32), g_thr);
        const bool keep = (row < p.n_rows) && ub_ord >= thr_ord;
        // kept rows are staged in shared memory and flushed in bulk: one global atomic per
        // ~1000 entries instead of one per warp (the single counter would serialise in L2)
        const uint32_t ballot = __ballot_sync(0xffffffffu, keep);
        if (ballot) {
            const uint32_t leader = __ffs(ballot) - 1;
            uint32_t base = 0;
            if (lane == leader) base = atomicAdd(kept_cnt_s, (uint32_t)__popc(ballot));
            base = __shfl_sync(0xffffffffu, base, leader);
            if (keep) {
                KeptEntry e;
                e.row = row;
                e.ub_ord = ub_ord;
                kept_s[base + __popc(ballot & ((1u << lane) - 1u))] = e;
            }
        }
        topk_offer(st, lb_key, t);
        // refresh the threshold eagerly while it is cheap (small buffer -> short sort): a
        // stale threshold lets thousands of hopeless rows into the kept list
        if (st.count >= max(2u * p.k, 64u) && st.count > p.k) {
            topk_prune(st, t);
            if (t == 0) atomicMax(p.ctl + 4, (uint32_t)(*st.thr_smem >> 32));
        }
        // (topk_offer's barrier made every append of this row block visible)
        if (kept_cnt_s[0] > kKeptStage - (uint32_t)kRowsPerBlock)
            flush_kept(kept_s, kept_cnt_s, p.kept, p.ctl, t);
    }
    consumer_sync();
    if (kept_cnt_s[0]) flush_kept(kept_s, kept_cnt_s, p.kept, p.ctl, t);

    // ---- k-th best lower bound over all CTAs ----
    topk_prune(st, t);
    uint64_t *my_cand = p.cand + (uint64_t)blockIdx.x * p.k;
    for (uint32_t i = t; i < p.k; i += kRowsPerBlock) my_cand[i] = (i < st.count) ? st.buf[i] : 0ull;
    __threadfence();
    consumer_sync();
    if (t == 0) *ticket_s = atomicAdd(p.ctl, 1u);
    consumer_sync();
    if (*ticket_s != gridDim.x - 1) return;
    __threadfence();
    MergeScratch ms;
    ms.hist = reinterpret_cast<uint32_t *>(stages);
    ms.sc = ms.hist + 256;
    if (t == 0) *st.cnt_smem = 0u;
    st.count = 0;
    st.cap = kCandCap;
    consumer_sync();
    merge_published(st, t, p.cand, gridDim.x * p.k, p.k, ms);
    if (t == 0) {
        // fewer than k rows in total: no threshold, everything kept is a candidate
        p.ctl[4] = (st.count >= p.k) ? (uint32_t)(st.buf[p.k - 1] >> 32) : 0u;
        p.ctl[0] = 0u;
        p.ctl[1] = 0u;
    }
}

// ---------------------------------------------------------------------------------------
// exact re-score of the surviving candidates + final selection
// ---------------------------------------------------------------------------------------
struct RescoreParams {
    const float *query;
    const float *rows;      // f32 mirror
    uint32_t pitch;
    uint32_t dim;
    const KeptEntry *kept;
    uint32_t *ctl;          // as PrefilterParams::ctl; [6] ticket of this kernel
    uint64_t *exact_keys;   // [kKeptCap] exact keys of the survivors
    uint64_t *out_rows;     // [k]
    float *out_scores;      // [k]
    uint32_t *out_count;
    uint64_t row_base;
    uint32_t k;
    int metric;
};

constexpr uint32_t kRescoreChunk = 512;   // floats of a row staged per warp at a time

__global__ void __launch_bounds__(kRowsPerBlock)
prefilter_rescore_kernel(const RescoreParams p) {
    __shared__ __align__(16) uint64_t buf[kCandCap];          // final selection buffer
    __shared__ __align__(16) float stage_x[8][kRescoreChunk]; // one row chunk per warp
    __shared__ uint32_t surv[2048];                           // surviving rows of this round
    __shared__ uint64_t thr_s;
    __shared__ uint32_t cnt_s, ticket_s, n_surv;
    __shared__ uint32_t hist[256 + 16];
    __shared__ float qmag_s;
    const uint32_t t = threadIdx.x, lane = t & 31u, warp = t >> 5;
    const uint32_t kept = min(p.ctl[2], kKeptCap);
    const uint32_t tau = p.ctl[4];
    if (t == 0) {
        thr_s = 0ull;
        cnt_s = 0u;
    }
    // |q| (reference lane tree), warp 0
    if (t < 32u) {
        float acc = 0.0f;
        const uint32_t chunks = p.dim / 8u;
        if (lane < 8u)
            for (uint32_t c = 0; c < chunks; ++c) {
                float v = __ldg(p.query + c * 8u + lane);
                acc = __fadd_rn(acc, __fmul_rn(v, v));
            }
        float r = 0.0f;
#pragma unroll
        for (int j = 0; j < 8; ++j) r = __fadd_rn(r, __shfl_sync(0xffffffffu, acc, j));
        if (lane == 0) {
            for (uint32_t i = chunks * 8u; i < p.dim; ++i) {
                float v = __ldg(p.query + i);
                r = __fadd_rn(r, __fmul_rn(v, v));
            }
            qmag_s = __fsqrt_rn(r);
        }
    }
    __syncthreads();
    const float qmag = qmag_s;
    // this CTA's contiguous slice of the kept list, in rounds of 2048 entries
    const uint32_t per_cta = (kept + gridDim.x - 1) / gridDim.x;
    const uint32_t lo = blockIdx.x * per_cta, hi = min(kept, lo + per_cta);
    for (uint32_t base = lo; base < hi; base += 2048u) {
        if (t == 0) n_surv = 0u;
        __syncthreads();
        for (uint32_t i = base + t; i < min(hi, base + 2048u); i += kRowsPerBlock) {
            const KeptEntry e = p.kept[i];
            if (e.ub_ord >= tau) surv[atomicAdd(&n_surv, 1u)] = e.row;
        }
        __syncthreads();
        const uint32_t ns = n_surv;
        // one warp per surviving row: stage the row chunk by chunk (coalesced), lanes 0..7 run
        // the f32x8 lanes over it in the reference order, lane 0 folds and finishes the score
        for (uint32_t si = warp; si < ns; si += 8u) {
            const uint32_t row = surv[si];
            const float *x = p.rows + (size_t)row * p.pitch;
            const uint32_t full = (p.dim / 8u) * 8u;  // elements covered by whole f32x8 groups
            float dacc = 0.0f, sacc = 0.0f, eacc = 0.0f;
            for (uint32_t c0 = 0; c0 < p.dim; c0 += kRescoreChunk) {
                const uint32_t len = min(kRescoreChunk, p.dim - c0);
                __syncwarp();
                for (uint32_t i = lane; i < len; i += 32u) stage_x[warp][i] = x[c0 + i];
                __syncwarp();
                if (p.metric == kEuclidean) {
                    if (lane == 0)
                        for (uint32_t i = 0; i < len; ++i) {
                            float df = __fsub_rn(__ldg(p.query + c0 + i), stage_x[warp][i]);
                            eacc = __fadd_rn(eacc, __fmul_rn(df, df));
                        }
                } else if (lane < 8u) {
                    // kRescoreChunk is a multiple of 8, so lane j keeps owning i = j (mod 8)
                    for (uint32_t i = lane; i < len && c0 + i < full; i += 8u) {
                        const float xv = stage_x[warp][i];
                        dacc = __fadd_rn(dacc, __fmul_rn(__ldg(p.query + c0 + i), xv));
                        sacc = __fadd_rn(sacc, __fmul_rn(xv, xv));
                    }
                }
            }
            float score;
            if (p.metric == kEuclidean) {
                score = __fdiv_rn(1.0f, __fadd_rn(1.0f, __fsqrt_rn(eacc)));
            } else {
                float dot = 0.0f, ssq = 0.0f;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    dot = __fadd_rn(dot, __shfl_sync(0xffffffffu, dacc, j));
                    ssq = __fadd_rn(ssq, __shfl_sync(0xffffffffu, sacc, j));
                }
                if (lane == 0) {
                    for (uint32_t i = full; i < p.dim; ++i) {  // scalar tail, straight from global
                        const float xv = x[i];
                        dot = __fadd_rn(dot, __fmul_rn(__ldg(p.query + i), xv));
                        ssq = __fadd_rn(ssq, __fmul_rn(xv, xv));
                    }
                }
                if (p.metric == kDot) {
                    score = dot;
                } else {
                    const float rmag = __fsqrt_rn(ssq);
                    score = (qmag == 0.0f || rmag == 0.0f) ? 0.0f
                                                           : __fdiv_rn(dot, __fmul_rn(qmag, rmag));
                }
            }
            if (lane == 0) {
                const uint32_t pos = atomicAdd(p.ctl + 5, 1u);
                p.exact_keys[pos] = make_key(__float_as_uint(score), row);
            }
        }
        __syncthreads();
    }
    __threadfence();
    __syncthreads();
    if (t == 0) ticket_s = atomicAdd(p.ctl + 6, 1u);
    __syncthreads();
    if (ticket_s != gridDim.x - 1) return;
    __threadfence();
    TopKState st;
    st.buf = buf;
    st.cnt_smem = &cnt_s;
    st.thr_smem = &thr_s;
    st.count = 0;
    st.k = p.k;
    st.cap = kCandCap;
    MergeScratch ms;
    ms.hist = hist;
    ms.sc = hist + 256;
    const uint32_t n_exact = *reinterpret_cast<volatile uint32_t *>(p.ctl + 5);
    merge_published(st, t, p.exact_keys, n_exact, p.k, ms);
    TopKOutputs o;
    o.out_keys = nullptr;
    o.out_hits = nullptr;
    o.out_rows = p.out_rows;
    o.out_scores = p.out_scores;
    o.out_count = p.out_count;
    o.row_base = p.row_base;
    o.accumulate_count = 0;
    write_outputs(st, t, p.k, o);
    if (t == 0) {
        p.ctl[7] = p.ctl[2];  // kept entries, for the statistics
        p.ctl[2] = 0u;
        p.ctl[5] = 0u;
        p.ctl[6] = 0u;
    }
}

#endif  // __CUDACC__
}  // namespace nm
