// batch_kernels.cuh — batched-query SIMILAR scan (BASELINE config 4: "batch 256 queries").
//
// The reference has no batch API: a batch is nq independent
// search_similar_with_metric calls, so every (query, row) score must still be the bit-exact
// result of compute_score (vector_engine/src/lib.rs:2231-2246).  What a GPU can share is the
// corpus pass: one trip of the rows through shared memory serves QB queries.
//
//   prepare_batch_kernel   transposes QB queries into chunk-major [dim/32][32][QB] order (so a
//                          k-chunk of all QB queries is ONE contiguous bulk copy) and computes
//                          |q| per query with the f32x8 lane tree.
//   score_batch_kernel     same TMA stage ring as scan_topk_kernel plus a 1-D bulk copy of
//                          the query chunk per stage; each consumer thread owns one corpus row
//                          and QB accumulator sets in registers, updated with the same
//                          non-fusable __fsub_rn/__fmul_rn/__fadd_rn as the 1-query kernel.
//                          (Packed f32x2 PTX was tried and dropped: ptxas contracts
//                          mul.rn.f32x2 + add.rn.f32x2 into FFMA2, which breaks bit parity, and
//                          FADD2/FFMA2 issue at half rate on sm_100a, so there is no FP32
//                          throughput to gain — measured with scripts/mb_fp32x2.cu.)
//                          Writes scores[q][row].
//   select_batch_kernel    per query: stream the score row, keys -> threshold + candidate
//                          buffer -> per-CTA top-k -> last CTA per query merges (radix select).
//
// With QB queries per pass the scan is FP32-issue bound, not HBM bound (3 lane-ops per
// element per query for L2, 2 for dot/cosine): see DESIGN.md 4.4 for the roofline.
#pragma once
#include "scan_kernels.cuh"

namespace nm {

#ifdef __CUDACC__

// 1-D bulk copy global -> shared, completion counted in bytes on `bar`.
__device__ __forceinline__ void bulk_load_1d(void *dst, const void *src, uint32_t bytes,
                                             uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
            "r"(smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// ---------------------------------------------------------------------------------------
// prepare: queries [nq_pass, dim] (row-major) -> qt [n_kc][32][QB] (+ zero padding), qmag[QB]
// ---------------------------------------------------------------------------------------
__global__ void prepare_batch_kernel(const float *queries, uint32_t nq_pass, uint32_t dim,
                                     uint32_t qb, uint32_t n_kc, float *qt, float *qmag,
                                     const uint32_t *gate) {
    if (!gate_open(gate, nq_pass)) return;  // conditional pass: no query of it was flagged
    const uint32_t total = n_kc * 32u * qb;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += gridDim.x * blockDim.x) {
        const uint32_t q = i % qb;
        const uint32_t e = (i / qb) % 32u;
        const uint32_t kc = i / (qb * 32u);
        const uint32_t col = kc * 32u + e;
        qt[i] = (q < nq_pass && col < dim) ? queries[(size_t)q * dim + col] : 0.0f;
    }
    // |q| with the lane tree (hnsw.rs:198-229): one warp per query, lanes 0..7 = f32x8 lanes
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t n_warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t q = warp; q < qb; q += n_warps) {
        float r = 0.0f;
        if (q < nq_pass) {
            const float *v = queries + (size_t)q * dim;
            const uint32_t chunks = dim / 8u;
            float acc = 0.0f;
            if (lane < 8u)
                for (uint32_t c = 0; c < chunks; ++c) {
                    float x = v[c * 8u + lane];
                    acc = __fadd_rn(acc, __fmul_rn(x, x));
                }
#pragma unroll
            for (int j = 0; j < 8; ++j) r = __fadd_rn(r, __shfl_sync(0xffffffffu, acc, j));
            for (uint32_t i = chunks * 8u; i < dim; ++i) r = __fadd_rn(r, __fmul_rn(v[i], v[i]));
            r = __fsqrt_rn(r);
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) (void)__shfl_sync(0xffffffffu, 0.0f, j);
        }
        if (lane == 0) qmag[q] = r;
    }
}

// ---------------------------------------------------------------------------------------
// score: all (query, row) scores of one pass
// ---------------------------------------------------------------------------------------
struct BatchScoreParams {
    const float *qt;         // [n_kc][32][QB]
    const float *qmag;       // [QB]
    float *scores;           // [QB][score_stride]
    uint32_t *cursor;        // dynamic row-block cursor (reset to 0 by the last CTA to finish)
    uint32_t *done;          // finish ticket (reset likewise)
    uint64_t score_stride;   // floats between consecutive queries' score rows
    uint32_t n_rows;
    uint32_t dim;
    uint32_t n_stages;
    uint32_t evict_first;
    const uint32_t *gate;    // conditional pass: runs only if any of gate[0..gate_n) is set
    uint32_t gate_n;
};

// Stage stride: corpus box + query chunk, rounded up to 1 KiB because the next stage's corpus
// box must start 1024-byte aligned for SWIZZLE_128B (QB = 4 has a 512-byte query chunk).
template <int QB>
__host__ __device__ constexpr int batch_stage_bytes() {
    return kStageBytes + ((32 * QB * 4 + 1023) / 1024) * 1024;
}
// Bytes that actually land in a stage (what the mbarrier expects).
template <int QB>
__host__ __device__ constexpr int batch_stage_tx_bytes() {
    return kStageBytes + 32 * QB * 4;
}

template <int METRIC, int QB>
__global__ void __launch_bounds__(kScanThreads, 1)
score_batch_kernel(const __grid_constant__ CUtensorMap tmap, const BatchScoreParams p) {
    static_assert(QB % 4 == 0, "QB must be a multiple of 4");
    constexpr int kStage = batch_stage_bytes<QB>();
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    // stage s: [corpus box 32 KiB][query chunk 32*QB floats]; stage size is a multiple of 1024
    uint8_t *stages = smem;
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(stages + p.n_stages * kStage);
    uint64_t *empty_bar = full_bar + kMaxStages;
    uint32_t *rb_ring = reinterpret_cast<uint32_t *>(empty_bar + kMaxStages);
    float *qmag_s = reinterpret_cast<float *>(rb_ring + kMaxStages);

    const uint32_t tid = threadIdx.x;
    const uint32_t warp = tid >> 5;
    if (!gate_open(p.gate, p.gate_n)) return;
    const uint32_t n_stages = p.n_stages;
    const uint32_t n_rb = (p.n_rows + kRowsPerBlock - 1) / kRowsPerBlock;
    const uint32_t n_kc = (p.dim + kChunkFloats - 1) / kChunkFloats;

    if (tid == 0) {
        for (uint32_t s = 0; s < n_stages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], kConsumerWarps);
        }
        fence_mbar_init();
    }
    if (tid < QB) qmag_s[tid] = p.qmag[tid];
    __syncthreads();

    if (warp == kConsumerWarps) {
        if (tid == kRowsPerBlock) {
            const uint64_t policy = p.evict_first ? policy_evict_first() : policy_evict_normal();
            uint32_t stage = 0, phase = 0;
            uint32_t rb = blockIdx.x;
            for (;;) {
                if (rb >= n_rb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1u);
                    rb_ring[stage] = 0xffffffffu;
                    mbar_arrive(&full_bar[stage]);
                    break;
                }
                const uint32_t next = atomicAdd(p.cursor, 1u) + gridDim.x;
                for (uint32_t kc = 0; kc < n_kc; ++kc) {
                    mbar_wait(&empty_bar[stage], phase ^ 1u);
                    if (kc == 0) rb_ring[stage] = rb;
                    uint8_t *sb = stages + stage * kStage;
                    mbar_arrive_expect_tx(&full_bar[stage], batch_stage_tx_bytes<QB>());
                    tma_load_2d(sb, &tmap, (int32_t)(kc * kChunkFloats),
                                (int32_t)(rb * kRowsPerBlock), &full_bar[stage], policy);
                    bulk_load_1d(sb + kStageBytes, p.qt + (size_t)kc * 32u * QB, 32u * QB * 4u,
                                 &full_bar[stage]);
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

    const uint32_t t = tid;
    const uint32_t lane = t & 31u;
    const uint32_t swz = t & 7u;
    uint32_t stage = 0, phase = 0;

    // accumulators
    //   Euclidean: acc[q]                         (one sequential chain per query)
    //   dot/cos  : acc[lane j][q] (8 f32x8 lanes per query), ss[j] shared by all queries
    constexpr int kAcc = (METRIC == kEuclidean) ? QB : 8 * QB;
    float acc[kAcc];
    float ss[8];

    for (;;) {
        mbar_wait(&full_bar[stage], phase);
        const uint32_t rb = rb_ring[stage];
        if (rb == 0xffffffffu) break;
#pragma unroll
        for (int i = 0; i < kAcc; ++i) acc[i] = 0.0f;
#pragma unroll
        for (int j = 0; j < 8; ++j) ss[j] = 0.0f;

        for (uint32_t kc = 0; kc < n_kc; ++kc) {
            mbar_wait(&full_bar[stage], phase);
            const uint8_t *srow = stages + stage * kStage + t * 128u;
            const float *qs = reinterpret_cast<const float *>(stages + stage * kStage + kStageBytes);
            // two float4 units (8 consecutive elements = one f32x8 group) per iteration
            if (METRIC == kEuclidean) {
#pragma unroll 1
                for (uint32_t u = 0; u < 8; u += 2) {
                    const float4 v0 = *reinterpret_cast<const float4 *>(srow + ((u ^ swz) << 4));
                    const float4 v1 = *reinterpret_cast<const float4 *>(srow + (((u + 1) ^ swz) << 4));
                    const float xv[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
                    const float *qe = qs + (u * 4u) * QB;
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const float x = xv[e];
                        const float4 *qp = reinterpret_cast<const float4 *>(qe + e * QB);
#pragma unroll
                        for (int g = 0; g < QB / 4; ++g) {
                            const float4 qq = qp[g];  // queries 4g..4g+3 at this element (broadcast)
                            const float d0 = __fsub_rn(qq.x, x), d1 = __fsub_rn(qq.y, x);
                            const float d2 = __fsub_rn(qq.z, x), d3 = __fsub_rn(qq.w, x);
                            acc[4 * g + 0] = __fadd_rn(acc[4 * g + 0], __fmul_rn(d0, d0));
                            acc[4 * g + 1] = __fadd_rn(acc[4 * g + 1], __fmul_rn(d1, d1));
                            acc[4 * g + 2] = __fadd_rn(acc[4 * g + 2], __fmul_rn(d2, d2));
                            acc[4 * g + 3] = __fadd_rn(acc[4 * g + 3], __fmul_rn(d3, d3));
                        }
                    }
                }
            } else {
                // dot / cosine: the query values of element e+1 are fetched while element e is
                // being accumulated (explicit double buffer: with 2 warps per scheduler the
                // 29-cycle LDS latency is otherwise exposed, ncu short_scoreboard stalls)
                constexpr int G = QB / 4;
                float4 qn[G];
                {
                    const float4 *qp = reinterpret_cast<const float4 *>(qs);
#pragma unroll
                    for (int g = 0; g < G; ++g) qn[g] = qp[g];
                }
#pragma unroll 1
                for (uint32_t u = 0; u < 8; u += 2) {
                    const float4 v0 = *reinterpret_cast<const float4 *>(srow + ((u ^ swz) << 4));
                    const float4 v1 = *reinterpret_cast<const float4 *>(srow + (((u + 1) ^ swz) << 4));
                    const float xv[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
                    const float *qe = qs + (u * 4u) * QB;
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const float x = xv[e];
                        if (METRIC == kCosine) ss[e] = __fadd_rn(ss[e], __fmul_rn(x, x));
                        float4 qc[G];
#pragma unroll
                        for (int g = 0; g < G; ++g) qc[g] = qn[g];
                        // next element of this stage (the last one re-reads element 31: harmless)
                        const uint32_t nxt = min(u * 4u + (uint32_t)e + 1u, 31u);
                        const float4 *qp = reinterpret_cast<const float4 *>(qs + nxt * QB);
                        (void)qe;
#pragma unroll
                        for (int g = 0; g < G; ++g) qn[g] = qp[g];
#pragma unroll
                        for (int g = 0; g < G; ++g) {
                            float *a = acc + e * QB + 4 * g;
                            a[0] = __fadd_rn(a[0], __fmul_rn(qc[g].x, x));
                            a[1] = __fadd_rn(a[1], __fmul_rn(qc[g].y, x));
                            a[2] = __fadd_rn(a[2], __fmul_rn(qc[g].z, x));
                            a[3] = __fadd_rn(a[3], __fmul_rn(qc[g].w, x));
                        }
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty_bar[stage]);
            if (++stage == n_stages) {
                stage = 0;
                phase ^= 1u;
            }
        }

        // ---- finalise QB scores of this row ----
        const uint32_t row = rb * kRowsPerBlock + t;
        float rmag = 0.0f;
        if (METRIC == kCosine) rmag = __fsqrt_rn(fold_lanes(ss));
        if (row < p.n_rows) {
#pragma unroll
            for (int q = 0; q < QB; ++q) {
                float sc;
                if (METRIC == kEuclidean) {
                    sc = __fdiv_rn(1.0f, __fadd_rn(1.0f, __fsqrt_rn(acc[q])));
                } else {
                    float d = 0.0f;
#pragma unroll
                    for (int l = 0; l < 8; ++l) d = __fadd_rn(d, acc[l * QB + q]);
                    if (METRIC == kCosine) {
                        const float m = qmag_s[q];
                        sc = (m == 0.0f || rmag == 0.0f) ? 0.0f : __fdiv_rn(d, __fmul_rn(m, rmag));
                    } else {
                        sc = d;
                    }
                }
                p.scores[(size_t)q * p.score_stride + row] = sc;
            }
        }
    }
    // last CTA to finish resets the cursors for the next launch
    consumer_sync();
    if (t == 0) {
        __threadfence();
        if (atomicAdd(p.done, 1u) == gridDim.x - 1) {
            *p.cursor = 0u;
            *p.done = 0u;
        }
    }
}

// ---------------------------------------------------------------------------------------
// select: top-k per query over the score matrix
// ---------------------------------------------------------------------------------------
struct BatchSelectParams {
    const float *scores;     // [nq_pass][score_stride]
    uint64_t score_stride;
    uint64_t *cand;          // [nq_pass][ctas_per_query][k]
    uint32_t *tickets;       // [nq_pass], left at 0
    ShardHit *out_hits;      // [nq_pass][out_stride] (may be null)
    uint64_t *out_rows;      // [nq_pass][out_stride] (may be null)
    float *out_scores;       // [nq_pass][out_stride] (may be null)
    uint32_t *out_counts;    // [nq_pass] (may be null)
    uint64_t row_base;
    uint32_t out_stride;     // slots per query in the outputs (the caller's k)
    uint32_t n_rows;
    uint32_t k;              // <= kMaxFastK
    const uint32_t *gate;    // conditional pass (see BatchScoreParams)
    uint32_t gate_n;
};

__global__ void __launch_bounds__(kRowsPerBlock)
select_batch_kernel(const BatchSelectParams p) {
    __shared__ __align__(16) uint64_t buf[kCandCap];
    __shared__ uint64_t thr_s;
    __shared__ uint32_t cnt_s, ticket_s;
    __shared__ uint32_t hist[256 + 16];
    const uint32_t t = threadIdx.x;
    const uint32_t q = blockIdx.y;
    const uint32_t n_cta = gridDim.x;
    if (!gate_open(p.gate, p.gate_n)) return;
    if (t == 0) {
        thr_s = 0ull;
        cnt_s = 0u;
    }
    __syncthreads();
    TopKState st;
    st.buf = buf;
    st.cnt_smem = &cnt_s;
    st.thr_smem = &thr_s;
    st.count = 0;
    st.k = p.k;
    st.cap = 512u;
    while (st.cap < p.k + (uint32_t)kRowsPerBlock) st.cap <<= 1;

    const float *srow = p.scores + (size_t)q * p.score_stride;
    const uint32_t n_rb = (p.n_rows + kRowsPerBlock - 1) / kRowsPerBlock;
    for (uint32_t rb = blockIdx.x; rb < n_rb; rb += n_cta) {
        const uint32_t row = rb * kRowsPerBlock + t;
        uint64_t key = 0ull;
        if (row < p.n_rows) key = make_key(__float_as_uint(__ldcs(srow + row)), row);
        topk_offer(st, key, t);
    }
    topk_prune(st, t);
    uint64_t *lists = p.cand + (size_t)q * n_cta * p.k;
    uint64_t *mine = lists + (size_t)blockIdx.x * p.k;
    for (uint32_t i = t; i < p.k; i += kRowsPerBlock) mine[i] = (i < st.count) ? st.buf[i] : 0ull;
    __threadfence();
    consumer_sync();
    if (t == 0) ticket_s = atomicAdd(p.tickets + q, 1u);
    consumer_sync();
    if (ticket_s != n_cta - 1) return;
    __threadfence();
    MergeScratch ms;
    ms.hist = hist;
    ms.sc = hist + 256;
    if (t == 0) cnt_s = 0u;
    st.count = 0;
    st.cap = kCandCap;
    consumer_sync();
    merge_published(st, t, lists, n_cta * p.k, p.k, ms);
    TopKOutputs o;
    o.out_keys = nullptr;
    o.out_hits = p.out_hits ? p.out_hits + (size_t)q * p.out_stride : nullptr;
    o.out_rows = p.out_rows ? p.out_rows + (size_t)q * p.out_stride : nullptr;
    o.out_scores = p.out_scores ? p.out_scores + (size_t)q * p.out_stride : nullptr;
    o.out_count = p.out_counts ? p.out_counts + q : nullptr;
    o.row_base = p.row_base;
    o.accumulate_count = 0;
    write_outputs(st, t, p.k, o);
    if (t == 0) p.tickets[q] = 0u;
}

#endif  // __CUDACC__
}  // namespace nm
