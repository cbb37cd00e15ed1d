// tc_prefilter_kernels.cuh — tensor-core pre-filter for BATCHES of queries (BASELINE config 4:
// "10M x 1536 f32 L2 TOP 100, batch 256 queries"), with EXACT re-score.
//
// The reference has no batch API: a batch is nq independent
// search_similar_with_metric calls (vector_engine/src/lib.rs:2049-2101), so every returned
// (row, score) must still be the bit-exact result of compute_score (lib.rs:2231-2246).  The
// exact batched kernels (batch_kernels.cuh) are FP32-issue bound: 3 non-fusable lane-ops per
// (element, query).  This path gets the same answer from the int8 mirror copy
// (prefilter_kernels.cuh) and the 5th-generation tensor cores:
//
//   tc_prepare_queries_kernel   per query: int8 quantisation (own scale), sum|qt|, |q|^2
//                               bounds, reference-arithmetic |q|.
//   tc_gemm_filter_kernel       persistent GEMM  I[row, q] = sum_i xt[row,i] * qt[q,i]  with
//                               tcgen05.mma kind::i8 (s8 x s8 -> s32, EXACT): A = 128 corpus
//                               rows x 128 B and B = up to 256 queries x 128 B per stage, both
//                               K-major SWIZZLE_128B boxes staged by TMA; accumulators
//                               [128 lanes x 256 columns] s32 in TMEM, double buffered (512
//                               columns) so the epilogue of tile t overlaps the MMAs of t+1.
//                               Launched in CTA pairs (cta_group::2: each CTA stages its 128
//                               corpus rows and half of the queries, one M = 256 UMMA reads
//                               both) or, NM_TC_PAIR=0, one CTA per tile.  Warp roles: 0 = TMA
//                               producer, 1 = MMA issuer (one thread), 2..17 = epilogue (four
//                               warps per TMEM lane quarter, splitting the query columns).
//                               The epilogue never writes the score matrix.  Each exact
//                               integer dot product gives a rigorous interval [lb, ub] for
//                               the reference score (same error model as the 1-query
//                               pre-filter, extended to the scalar L2 chain); an entry is kept
//                               only if ub can still reach the query's running threshold tau
//                               (a score that k rows seen so far reach).  A coarse test per
//                               16-query chunk and a 6-instruction f32 test with provable
//                               slack screen every (row, query); the few that pass are parked
//                               in shared memory and, once the accumulator has been released,
//                               evaluated rigorously in double and appended to the query's
//                               kept list.
//   tc_refine_kernel            per query: radix-select the k-th best lb of the kept list,
//                               re-score the k entries with the best lower bounds exactly and
//                               take the k-th best exact score as tau, compact the list to
//                               ub >= tau, derive the screen coefficients and the row range of
//                               the next phase (control block in device memory).
//   tc_sort_* / tc_score_sorted_kernel / tc_select_kernel
//                               survivors of all queries bucket-sorted by row and re-scored
//                               from the f32 mirror in corpus order with the reference
//                               arithmetic (a scattered gather is bound by address
//                               translation: 1.2 vs 4.3 TB/s), then selected per query exactly
//                               like the f32 scan (same keys, same tie rule).
//                               (tc_rescore_kernel: the per-query form, for dim % 4 != 0.)
//   tc_pack_hits_kernel         sharded indexes: a shard's result as ShardHit[nq, k].
//
// The corpus is walked in phases of geometrically growing row ranges (first 2048 rows: keep
// everything; then x4..x15 each, sized on the device from the pass rate just observed) with a
// refine step between phases, so tau tightens while only ~k..16k entries per query and phase
// are kept.  Every row of the exact top-k has ub >= score >= tau, so the result is
// bit-identical to the f32 scan; a query whose list overflows or that is not finite is flagged
// and redone by the exact path.  tests/test_tc_model_cpu.py and tests/test_tc_screen_cpu.py
// restate the interval and the screen in numpy; keep their constants in step with this file.
#pragma once
#include "prefilter_kernels.cuh"

namespace nm {

constexpr uint32_t kTcM = 128;                             // corpus rows per tile == TMEM lanes
constexpr uint32_t kTcKBytes = 128;                        // int8 elements per k-block
constexpr uint32_t kTcMaxQ = 256;                          // queries per pass == UMMA N max
// One TMA ring; a stage = the corpus box + the query box of one k-block.  CTAS = 1: the CTA
// stages all 256 queries (16 + 32 KiB, 4 stages).  CTAS = 2 (cta_group::2): the two CTAs of a
// cluster each stage their own 128 corpus rows and HALF of the queries (16 + 16 KiB, 6 stages)
// and one UMMA of M = 256 reads both halves, so the query bytes each SM pulls from L2 halve
// and more k-blocks are in flight: the loop is bound by bytes in flight / memory latency.
// The loop is bound by SHARED-MEMORY bandwidth, not by HBM or the tensor pipe: every operand
// byte is written once by TMA and read once by the UMMA (s8 x s8 at N = 256 reads 96 B/clk,
// restaging the queries for every corpus tile writes another 64-96 B/clk, the SM moves
// 128 B/clk).  So whenever this CTA's share of the int8 queries fits beside >= 5 corpus stages
// it is staged ONCE per launch ("resident": 256 queries x dim <= 896 on a pair, any dim for
// small batches) and the ring carries only corpus tiles (16 KiB stages, up to 10).
constexpr uint32_t kTcStages1 = 4;
constexpr uint32_t kTcStages2 = 6;
constexpr uint32_t kTcMaxStages = 10;
constexpr uint32_t kTcRingBytes = 192u * 1024u;
constexpr uint32_t kTcABytes = kTcM * kTcKBytes;           // 16 KiB
constexpr uint32_t kTcBBytesMax = kTcMaxQ * kTcKBytes;     // 32 KiB
constexpr uint32_t kTcEpilogueWarps = 16;                  // 4 per TMEM lane quarter
constexpr uint32_t kTcColParts = kTcEpilogueWarps / 4;     // they split the 16-query chunks
constexpr uint32_t kTcRoleWarps = 2;                       // TMA producer, MMA issuer
constexpr uint32_t kTcThreads = 32 * (kTcRoleWarps + kTcEpilogueWarps);
constexpr uint32_t kTcKeptCap = 32768;                     // kept entries per query
constexpr uint32_t kTcPhase0Rows = 2048;                   // first phase: keep everything (>= k)
constexpr uint32_t kTcTmemCols = 512;                      // 2 accumulators x 256 columns
constexpr uint32_t kTcPendCap = 64;                        // parked entries per epilogue warp
constexpr uint32_t kTcPendDrain = 24;                      // drained once this many are parked

constexpr uint32_t kTcFlagUnusable = 1u;   // query not finite / zero / denormal scale
constexpr uint32_t kTcFlagOverflow = 2u;   // kept list overflowed

struct TcKept {
    uint32_t row;
    uint32_t lb_ord;
    uint32_t ub_ord;
};

struct alignas(16) TcQueryMeta {
    double c_lo, c_hi;   // bounds on the real |q|^2
    float s_q;           // quantisation scale
    float qmag;          // reference lane-tree |q|
    uint32_t q1;         // sum |qt_i|
    uint32_t flags;      // kTcFlag*
    float qnorm;         // >= ||qt||_2            (Cauchy-Schwarz form of the error bound)
    float enorm;         // >= ||q / s_q - qt||_2
    uint32_t tau_ord;    // k-th best lower bound so far (0 = none yet)
    uint32_t pad;
};

// Phase control of one pass, device resident: the refine kernel sizes the next row range from
// the pass rate it just observed, so the host enqueues a fixed number of (gemm, refine) pairs
// without reading anything back; pairs past the end of the corpus exit at once.
struct TcCtl {
    uint32_t row_begin;   // current phase [row_begin, row_end)
    uint32_t row_end;
    uint32_t next_rows;   // atomicMin over queries: rows the next phase may cover
    uint32_t ticket;
    uint32_t phases;      // phases completed (statistics)
    uint32_t pad[3];
};

// the GEMM kernel stages the whole records in shared memory for the rigorous evaluation
using TcQm = TcQueryMeta;

inline size_t tc_gemm_smem_bytes() {
    return 1024 + (size_t)kTcRingBytes + (size_t)kTcMaxQ * 16 + 256 +
           (size_t)kTcEpilogueWarps * (kTcPendCap * 12 + 4) + (size_t)kTcMaxQ * 48 + 256;
}

#ifdef __CUDACC__

// ---------------------------------------------------------------------------------------
// PTX helpers: tcgen05 / TMEM
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void tc_fence_before() {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// mbarrier wait with a watchdog: a protocol bug traps (launch failure) instead of hanging the
// GPU.  SLEEP_NS > 0 backs off between polls so that a waiting role does not eat the issue
// slots of the warps that share its scheduler.
template <uint32_t SLEEP_NS>
__device__ __forceinline__ void mbar_wait_wd(uint64_t *bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    long long t0 = 0;
    for (uint32_t spins = 1;; ++spins) {
        uint32_t done;
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}\n"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
        if (done) return;
        if (SLEEP_NS) __nanosleep(SLEEP_NS);
        if ((spins & 1023u) == 0u) {
            if (t0 == 0) t0 = clock64();
            else if (clock64() - t0 > 6000000000ll) __trap();
        }
    }
}
// ---- thread-block cluster helpers (cta_group::2) ----
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of the same shared-memory offset in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_shared(uint32_t saddr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    // (default .release at CTA scope, as CUTLASS's ClusterBarrier::arrive(cta_id): the TMEM reads
    // are ordered by tcgen05.fence::before_thread_sync; a cluster-scope release costs a
    // MEMBAR.GPU per arrive)
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load of a cta_group::2 pair: data lands in THIS CTA's shared memory, the bytes are
// counted on the barrier at `bar_cluster_addr` (the leader CTA's)
__device__ __forceinline__ void tma_load_2d_pair(void *dst, const CUtensorMap *tmap, int32_t x,
                                                 int32_t y, uint32_t bar_cluster_addr,
                                                 uint64_t policy) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
        ".L2::cache_hint [%0], [%1, {%2, %3}], [%4], %5;" ::"r"(smem_u32(dst)),
        "l"(tmap), "r"(x), "r"(y), "r"(bar_cluster_addr), "l"(policy)
        : "memory");
}
__device__ __forceinline__ void tc_mma_i8_pair(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc,
                                               uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// commit of the pair's MMAs: arrives on the barrier at this offset in BOTH CTAs
__device__ __forceinline__ void tc_commit_pair(uint64_t *bar) {
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 "
        "[%0], %1;" ::"r"(smem_u32(bar)),
        "h"((uint16_t)3)
        : "memory");
}

// K-major SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor layout):
// [0,14) start address >> 4, [16,30) leading byte offset >> 4 (1: unused for swizzled
// K-major), [32,46) stride byte offset >> 4 (1024 B = 8 rows x 128 B), [46,48) version = 1,
// [61,64) layout type = 2 (SWIZZLE_128B).
__device__ __forceinline__ uint64_t tc_smem_desc(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3fffu) | (1ull << 16) | (64ull << 32) | (1ull << 46) |
           (2ull << 61);
}
// instruction descriptor (cute::UMMA::InstrDescriptor): c_format S32 = 2 @4, a/b format
// INT8 = 1 @7/@10, K-major A and B, N >> 3 @17, M >> 4 @24
__device__ __forceinline__ uint32_t tc_idesc_i8(uint32_t m, uint32_t n) {
    return (2u << 4) | (1u << 7) | (1u << 10) | ((n >> 3) << 17) | ((m >> 4) << 24);
}
__device__ __forceinline__ void tc_mma_i8(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc,
                                          uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tc_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                     smem_u32(bar))
                 : "memory");
}
// TMEM -> registers: 16 consecutive columns of this thread's lane.  The load is asynchronous;
// tc_ld_wait ties the registers to the wait so no consumer can be scheduled before it.
__device__ __forceinline__ void tc_ld16_issue(uint32_t taddr, int (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
          "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
          "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ int tc_ld1(uint32_t taddr) {
    int v;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v) : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" : "+r"(v) : : "memory");
    return v;
}
__device__ __forceinline__ void tc_ld_wait(int (&v)[16]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]),
                   "+r"(v[6]), "+r"(v[7]), "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]),
                   "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15])
                 :
                 : "memory");
}

// ---------------------------------------------------------------------------------------
// error model
// ---------------------------------------------------------------------------------------
// (tc_l2_score, kTcU and tc_row_sq_bounds live in prefilter_kernels.cuh: the single-query
// pre-filter uses the same Euclidean interval)

// Rigorous interval of the reference score of (row, query) from the exact integer dot I.
// wild == the analysis does not apply (non-finite data, possible overflow): always a candidate.
// nr = (>= ||xt||_2, >= ||x / s_r - xt||_2) of the row (quantize_rows_kernel).  Each of the three
// cross terms of  q.x = s_q s_r (I + qt.d + xt.e + e.d)  is bounded by the smaller of its
// Hoelder (L1 x Linf) and Cauchy-Schwarz (L2 x L2) bounds; BL is the L1-only bound the epilogue
// screen is built on (B <= BL, so everything kept here also passes the screen).
template <class QM>
__device__ __forceinline__ void tc_interval(int metric, int I, const RowMeta &m, const float2 nr,
                                            const QM &qm, uint32_t dim, uint32_t &lb_ord,
                                            uint32_t &ub_ord) {
    bool wild = (m.flags & 1u) != 0u || (qm.flags & kTcFlagUnusable) != 0u;
    const double g = 2.0 * ((double)dim + 16.0) * kTcU;
    const double ss = (double)qm.s_q * (double)m.scale;
    const double BL = 0.5001 * ((double)qm.q1 + (double)m.x1) + 0.2502 * (double)dim;
    const double B = fmin(0.5001 * (double)qm.q1, (double)qm.qnorm * (double)nr.y) +
                     fmin(0.5001 * (double)m.x1, (double)nr.x * (double)qm.enorm) +
                     fmin(0.2502 * (double)dim, (double)qm.enorm * (double)nr.y);
    const double Dt = ss * (double)I;
    float lb, ub;
    if (metric == kEuclidean) {
        // real dot within ss*B of ss*I; real d^2 = |x|^2 + |q|^2 - 2 dot; the f32 chain of
        // non-negative terms is within (dim + 4) u relative of the real d^2
        const double E = ss * B * 1.000001 + 1e-37;
        double a_lo, a_hi;
        tc_row_sq_bounds(m, dim, a_lo, a_hi);
        const double gc = ((double)dim + 4.0) * kTcU * 1.01;
        double lo = a_lo + qm.c_lo - 2.0 * (Dt + E);
        double hi = a_hi + qm.c_hi - 2.0 * (Dt - E);
        lo = lo * (1.0 - gc) * (1.0 - 1e-12) - 1e-36;
        hi = hi * (1.0 + gc) * (1.0 + 1e-12) + 1e-36;
        if (!(hi < 1e37) || !(a_hi < 1e37) || !(qm.c_hi < 1e37)) wild = true;
        const float slo = (lo > 0.0) ? __double2float_rd(lo) : 0.0f;
        const float shi = (hi > 0.0) ? __double2float_ru(hi) : 0.0f;
        ub = tc_l2_score(slo);
        lb = tc_l2_score(shi);
    } else {
        const double S = 127.51 * (double)m.x1 + BL;
        const double E = (ss * (B + g * S)) * 1.000001 + 1e-37;
        const float lo = __double2float_rd(Dt - E), hi = __double2float_ru(Dt + E);
        if (!(ss * S < 1e37) || !(fabs(Dt) + E < 1e37)) wild = true;
        if (metric == kCosine) {
            if (qm.qmag == 0.0f || m.rmag == 0.0f) {
                lb = ub = 0.0f;
            } else {
                const float den = __fmul_rn(qm.qmag, m.rmag);
                lb = __fdiv_rn(lo, den);
                ub = __fdiv_rn(hi, den);
                if (!(den > 1.17549435e-38f) || !(den < 3.0e38f)) wild = true;
            }
        } else {
            lb = lo;
            ub = hi;
        }
    }
    lb_ord = wild ? 0u : score_to_ord(__float_as_uint(lb));
    ub_ord = wild ? 0xffffffffu : score_to_ord(__float_as_uint(ub));
}

// Screen coefficients.  The epilogue keeps (row, q) for the rigorous test unless
//     float(I) + Br  <  alpha_r * w_q + beta_r * u_q + v_q          (all f32, fma)
// Derivation per metric (ss = s_q s_r, B = cB (Q1 + X1) + cD dim):
//   L2     ub >= tau  <=>  chain lower bound <= T (T = largest f32 s with score(s) >= tau)
//          <=>  I + B >= (A_lo + C_lo - T') / (2 ss):  alpha = A_lo/(2 s_r), beta = 1/(2 s_r),
//          w = 1/s_q, u = (C_lo - T')/s_q
//   dot    ub >= tau  <=>  ss (I + B') >= tau-:        alpha = 0, beta = 1/s_r, u = tau-/s_q
//   cosine ub >= tau  <=>  ss (I + B') >= tau- |q||x|: alpha = 0, beta = rmag/s_r,
//          u = tau- qmag / s_q
// v = -(query part of B), Br = row part of B.  Every per-query term is moved by 2^-16 of its
// magnitude (and Br by 2^-20 + the I2F slack) in the direction that keeps MORE, which covers
// the f32 roundings of the test (each <= 2^-24 of the sum of the term magnitudes) and the
// 1.000001 factors of the rigorous formulas.  NaN (inf - inf, 0 * inf) compares false: kept.
constexpr double kTcEpsQ = 1.52587890625e-05;      // 2^-16
constexpr double kTcEpsR = 9.5367431640625e-07;    // 2^-20
constexpr double kTcCB = 0.5001 * 1.000002;
constexpr double kTcCD = 0.2502 * 1.000002;

__device__ __forceinline__ float4 tc_pass_all(uint32_t tau_ord) {
    return make_float4(0.0f, 0.0f, -INFINITY, __uint_as_float(tau_ord));
}
// unusable queries are redone by the exact path: keep nothing for them
__device__ __forceinline__ float4 tc_pass_none() {
    return make_float4(0.0f, 0.0f, INFINITY, __uint_as_float(0xffffffffu));
}

__device__ __forceinline__ float4 tc_make_coef(int metric, uint32_t tau_ord, const TcQueryMeta &qm,
                                               uint32_t dim) {
    if ((qm.flags & kTcFlagUnusable) || !(qm.s_q > 0.0f)) return tc_pass_none();
    if (tau_ord == 0u) return tc_pass_all(tau_ord);
    const float tau = __uint_as_float(ord_to_score_bits(tau_ord));
    const double g = 2.0 * ((double)dim + 16.0) * kTcU;
    const double sq = (double)qm.s_q;
    double w = 0.0, u, v;
    if (metric == kEuclidean) {
        if (!(tau > 0.0f) || !(tc_l2_score(0.0f) >= tau)) return tc_pass_all(tau_ord);
        // T = largest non-negative f32 s with score(s) >= tau (score is monotone non-increasing)
        uint32_t lo_b = 0u, hi_b = 0x7f800000u;  // score(+inf) = 0 < tau
        while (hi_b - lo_b > 1u) {
            const uint32_t mid = lo_b + ((hi_b - lo_b) >> 1);
            if (tc_l2_score(__uint_as_float(mid)) >= tau) lo_b = mid;
            else hi_b = mid;
        }
        // keep while the (real) chain lower bound is below the next float above T
        const double Tn = (double)__uint_as_float(lo_b + 1u);
        if (!(Tn < 1e37)) return tc_pass_all(tau_ord);
        const double gc = ((double)dim + 4.0) * kTcU * 1.01;
        const double Tp = (Tn + 2e-36) / ((1.0 - gc) * (1.0 - 2e-12));
        w = 1.0 / sq;
        u = (qm.c_lo - Tp) / sq;
        v = -(kTcCB * (double)qm.q1 + kTcCD * (double)dim);
    } else {
        // hi = RU(Dt + E) >= tau needs Dt + E > pred(tau)
        double tp = (double)tau;
        if (metric == kCosine) {
            // the quotient rounds to >= tau only if hi/den >= tau - ulp; tiny |tau| (incl. 0,
            // where a negative quotient may round to -0.0 == +0.0) gets an absolute margin
            if (fabs(tp) < 1e-30) tp = fmin(tp, 0.0) - 1e-30;
            tp *= (double)qm.qmag;
        } else {
            tp = (double)__uint_as_float(ord_to_score_bits(tau_ord - 1u));  // pred(tau)
            if (!(fabs(tp) < 1e38)) return tc_pass_all(tau_ord);
            tp -= 1e-37;
        }
        u = tp / sq;
        v = -((1.0 + g) * 1.000002 * (0.5001 * (double)qm.q1 + 0.2502 * (double)dim));
    }
    w *= (1.0 - kTcEpsQ);
    u -= kTcEpsQ * fabs(u);
    v -= kTcEpsQ * fabs(v);
    return make_float4(__double2float_rd(w), __double2float_rd(u), __double2float_rd(v),
                       __uint_as_float(tau_ord));
}

// per-row screen coefficients (alpha, beta, Br), f32 with directed rounding (runs once per row
// and tile in every epilogue thread).  rc: per-launch constants from tc_row_consts.
struct TcRowConsts {
    float one_minus_rel;   // 1 - rel of tc_row_sq_bounds, rounded down
    float cb;              // row part of the interval half-width per unit of x1, rounded up
    float addc;            // I2F / rounding slack, rounded up
};
__device__ __forceinline__ TcRowConsts tc_row_consts(int metric, uint32_t dim) {
    const double g = 2.0 * ((double)dim + 16.0) * kTcU;
    const double rel = ((double)(dim / 8u) + 24.0) * kTcU * 1.01;
    const double cb = (metric == kEuclidean) ? kTcCB : 1.000002 * (0.5001 * (1.0 + g) + 127.51 * g);
    TcRowConsts rc;
    rc.one_minus_rel = __double2float_rd(1.0 - rel);
    rc.cb = __double2float_ru(cb * (1.0 + kTcEpsR));
    rc.addc = __double2float_ru(kTcEpsR * (16129.0 * (double)dim + 1.0));
    return rc;
}
__device__ __forceinline__ void tc_row_coef(int metric, const RowMeta &m, const TcRowConsts &rc,
                                            float &alpha, float &beta, float &br) {
    // tiny scales: the absolute slack terms (<= 4e-36 / (s_q s_r)) would no longer be covered
    if ((m.flags & 1u) || !(m.scale >= 1e-15f) || !(m.rmag < 3.0e38f)) {
        alpha = 0.0f;
        beta = 0.0f;
        br = INFINITY;  // always re-evaluated rigorously
        return;
    }
    if (metric == kEuclidean) {
        // alpha <= A_lo / (2 s_r) with A_lo as in tc_row_sq_bounds: every step rounds down
        float a_lo = __fmul_rd(__fmul_rd(m.rmag, m.rmag), rc.one_minus_rel);
        a_lo = fmaxf(__fsub_rd(a_lo, 1e-37f), 0.0f);
        const float two_s = __fmul_rn(2.0f, m.scale);  // exact
        alpha = __fmul_rd(__fmul_rd(a_lo, __frcp_rd(two_s)), 0.99999904f);  // (1 - 2^-20)
        beta = __frcp_rn(two_s);
    } else {
        alpha = 0.0f;
        beta = (metric == kCosine) ? __fdiv_rn(m.rmag, m.scale) : __frcp_rn(m.scale);
    }
    br = __fmaf_ru(rc.cb, __uint2float_ru(m.x1), rc.addc);
}

// ---------------------------------------------------------------------------------------
// prepare: one CTA (256 threads) per query slot; slots >= nq are zero rows of the B operand
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
tc_prepare_queries_kernel(const float *__restrict__ queries, uint32_t nq, uint32_t dim,
                          uint32_t pitch8, int8_t *q8, TcQueryMeta *qmeta, float4 *coef,
                          uint32_t *kept_n, uint32_t *kept_prev, TcCtl *ctl, uint32_t n_rows,
                          uint32_t masked) {
    __shared__ float red_f[8];
    __shared__ uint32_t red_u[8];
    __shared__ double red_d[8];
    __shared__ double red_e[8];
    __shared__ float qmag_s;
    const uint32_t q = blockIdx.x, t = threadIdx.x, lane = t & 31u, warp = t >> 5;
    uint32_t *out = reinterpret_cast<uint32_t *>(q8 + (size_t)q * pitch8);
    if (q == 0 && t == 0) {
        TcCtl c;
        c.row_begin = 0u;
        c.row_end = min(n_rows, kTcPhase0Rows);
        c.next_rows = 0xffffffffu;
        c.ticket = 0u;
        c.phases = 0u;
        c.pad[0] = c.pad[1] = c.pad[2] = 0u;
        *ctl = c;
    }
    if (q >= nq) {
        for (uint32_t w = t; w * 4u < pitch8; w += 256u) out[w] = 0u;
        return;
    }
    const float *v = queries + (size_t)q * dim;
    float mx = 0.0f;
    bool bad = false;
    double c = 0.0;
    for (uint32_t i = t; i < dim; i += 256u) {
        const float x = __ldg(v + i);
        bad |= !(fabsf(x) <= 3.4028234e38f);
        mx = fmaxf(mx, fabsf(x));
        c += (double)x * (double)x;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        bad |= __shfl_xor_sync(0xffffffffu, (int)bad, o) != 0;
        c += __shfl_xor_sync(0xffffffffu, c, o);
    }
    if (lane == 0) {
        red_f[warp] = mx;
        red_u[warp] = bad ? 1u : 0u;
        red_d[warp] = c;
    }
    __syncthreads();
    mx = 0.0f;
    bad = false;
    c = 0.0;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
        mx = fmaxf(mx, red_f[w]);
        bad |= red_u[w] != 0u;
        c += red_d[w];
    }
    const float s_q = bad ? 0.0f : __fdiv_rn(mx, 127.0f);
    if (!(s_q >= 1e-15f)) bad = true;  // zero / tiny scale: the exact path decides
    __syncthreads();
    uint32_t q1 = 0;
    double qq = 0.0, ee = 0.0;  // sum qt^2, sum (q / s_q - qt)^2
    for (uint32_t w = t; w * 4u < pitch8; w += 256u) {
        uint32_t packed = 0;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const uint32_t i = w * 4u + b;
            int x = 0;
            if (i < dim && !bad) {
                const float r = __fdiv_rn(__ldg(v + i), s_q);
                x = __float2int_rn(r);
                x = max(-127, min(127, x));
                const double e = (double)r - (double)x;
                qq += (double)(x * x);
                ee += e * e;
            }
            q1 += (uint32_t)abs(x);
            packed |= ((uint32_t)(uint8_t)(int8_t)x) << (8 * b);
        }
        out[w] = packed;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        qq += __shfl_xor_sync(0xffffffffu, qq, o);
        ee += __shfl_xor_sync(0xffffffffu, ee, o);
    }
    __syncthreads();  // red_d is reused
    if (lane == 0) {
        red_d[warp] = qq;
        red_e[warp] = ee;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) q1 += __shfl_xor_sync(0xffffffffu, q1, o);
    if (lane == 0) red_u[warp] = q1;
    // |q| with the reference lane tree (hnsw.rs:198-229), warp 0
    if (warp == 0) {
        float acc = 0.0f;
        const uint32_t chunks = dim / 8u;
        if (lane < 8u)
            for (uint32_t cix = 0; cix < chunks; ++cix) {
                const float x = __ldg(v + cix * 8u + lane);
                acc = __fadd_rn(acc, __fmul_rn(x, x));
            }
        float r = 0.0f;
#pragma unroll
        for (int j = 0; j < 8; ++j) r = __fadd_rn(r, __shfl_sync(0xffffffffu, acc, j));
        if (lane == 0) {
            for (uint32_t i = chunks * 8u; i < dim; ++i) {
                const float x = __ldg(v + i);
                r = __fadd_rn(r, __fmul_rn(x, x));
            }
            qmag_s = __fsqrt_rn(r);
        }
    }
    __syncthreads();
    if (t == 0) {
        q1 = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) q1 += red_u[w];
        TcQueryMeta m;
        m.c_lo = c * (1.0 - 1e-12) - 1e-40;
        if (m.c_lo < 0.0) m.c_lo = 0.0;
        m.c_hi = c * (1.0 + 1e-12) + 1e-40;
        m.s_q = bad ? 0.0f : s_q;
        m.qmag = qmag_s;
        m.q1 = q1;
        m.flags = bad ? kTcFlagUnusable : 0u;
        m.tau_ord = 0u;
        m.pad = 0u;
        // ||qt||_2 exactly (integers), ||e||_2 with the f32 rounding of q / s_q (|r| <= 127.01,
        // so each e_i is off by at most 7.6e-6) folded in; both rounded up
        double qq2 = 0.0, ee2 = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) {
            qq2 += red_d[w];
            ee2 += red_e[w];
        }
        m.qnorm = __double2float_ru(sqrt(qq2) * (1.0 + 1e-9));
        m.enorm = __double2float_ru((sqrt(ee2) + 7.7e-6 * sqrt((double)dim)) * (1.0 + 1e-9));
        qmeta[q] = m;
        coef[q] = bad ? tc_pass_none() : tc_pass_all(0u);
        // phase 0 fills slots [0, rows) directly; with a row mask it appends like every other phase
        kept_n[q] = (bad || masked) ? 0u : min(n_rows, kTcPhase0Rows);
        kept_prev[q] = 0u;
    }
}

// ---------------------------------------------------------------------------------------
// GEMM + filter
// ---------------------------------------------------------------------------------------
struct TcGemmParams {
    const RowMeta *meta;        // [rows]
    const float2 *norms;        // [rows] (||xt||_2, ||x / s_r - xt||_2), rounded up
    const TcQueryMeta *qmeta;   // [nq]
    const float4 *coef;         // [nq] screen coefficients of this phase
    TcKept *kept;               // [nq][kTcKeptCap]
    uint32_t *kept_n;           // [nq]
    int *dump;                  // debug: [nq][dump_stride] integer dot products (may be null)
    uint64_t dump_stride;
    const TcCtl *ctl;           // row range of this phase
    uint32_t n_rows;
    uint32_t dim;
    uint32_t nq;
    uint32_t n_pad;             // nq rounded up to 16 (UMMA N)
    uint32_t evict_first;
    uint32_t screen;            // 0: skip the f32 screen (every entry evaluated rigorously; tests)
    uint32_t shift;             // the screen works on I >> shift (|I >> shift| < 2^22)
    int metric;
    const uint32_t *row_mask;   // optional pre-filter: bit r set = row r takes part (may be null)
};

// An entry that passed the screen, parked until the accumulator has been handed back to the MMA
// issuer: the rigorous evaluation (double precision, a list reservation in global memory) must
// not sit between the TMEM reads of a tile and the release of its accumulator.
struct TcPend {         // the row's constants are re-read (L2) when the entry is evaluated
    int I;
    uint32_t q;
    uint32_t row;
};
// rigorous re-evaluation + append of ONE (row, query) entry that passed the screen.  Called
// divergently: every lane walks its own hits, so the latencies of the list reservations of a
// chunk overlap instead of queueing up column by column.
__device__ __noinline__ void tc_keep_entry(const TcGemmParams &p, const TcQm *qm_s, uint32_t q,
                                           int I, uint32_t row, const RowMeta &m, uint32_t tau_ord,
                                           bool phase0) {
    const TcQm &qm = qm_s[q];
    if (qm.flags & kTcFlagUnusable) return;
    const float2 nr = __ldg(p.norms + row);
    uint32_t lb_ord, ub_ord;
    tc_interval(p.metric, I, m, nr, qm, p.dim, lb_ord, ub_ord);
    if (ub_ord < tau_ord) return;
    // the first phase keeps every row of [0, kTcPhase0Rows): slot == row, the list length was
    // set by the prepare kernel, no reservation needed
    const uint32_t pos = phase0 ? row : atomicAdd(p.kept_n + q, 1u);
    if (pos < kTcKeptCap) {
        TcKept e;
        e.row = row;
        e.lb_ord = lb_ord;
        e.ub_ord = ub_ord;
        p.kept[(size_t)q * kTcKeptCap + pos] = e;
    }
}

// CTAS = 1: one CTA per 128-row tile.  CTAS = 2: launched in clusters of two; the pair owns a
// 256-row super tile (CTA r: rows r*128..), CTA 0 (the leader) issues the MMAs for both.
template <int CTAS>
__global__ void __launch_bounds__(kTcThreads, 1)
tc_gemm_filter_kernel(const __grid_constant__ CUtensorMap tmap_a,
                      const __grid_constant__ CUtensorMap tmap_q,
                      const __grid_constant__ TcGemmParams p) {
    constexpr uint32_t kStagesStream = CTAS == 2 ? kTcStages2 : kTcStages1;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t *ring = smem;                                          // stage: [A 16 KiB][B ...]
    float4 *coef_s = reinterpret_cast<float4 *>(ring + kTcRingBytes);
    float4 *cmin_s = coef_s + kTcMaxQ;             // [16] per 16-query chunk: min w, min u, min v
    TcPend *pend_s = reinterpret_cast<TcPend *>(cmin_s + kTcMaxQ / 16u);  // [warps][kTcPendCap]
    uint32_t *pend_cnt_s = reinterpret_cast<uint32_t *>(pend_s + kTcEpilogueWarps * kTcPendCap);
    TcQm *qm_s = reinterpret_cast<TcQm *>(pend_cnt_s + kTcEpilogueWarps);  // [256]
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(qm_s + kTcMaxQ);
    uint64_t *empty_bar = full_bar + kTcMaxStages;
    uint64_t *tfull_bar = empty_bar + kTcMaxStages;  // [2] accumulator ready
    uint64_t *tempty_bar = tfull_bar + 2;            // [2] accumulator drained
    uint64_t *bfull_bar = tempty_bar + 2;            // resident queries staged
    uint32_t *tmem_base_s = reinterpret_cast<uint32_t *>(bfull_bar + 1);

    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31u;
    const uint32_t row_begin = p.ctl->row_begin, row_end = min(p.ctl->row_end, p.n_rows);
    if (row_begin >= row_end) return;  // past the end of the corpus: nothing to do (uniform)
    const uint32_t cr = CTAS == 2 ? cluster_ctarank() : 0u;        // rank in the pair
    const uint32_t unit = blockIdx.x / CTAS, n_units = gridDim.x / CTAS;  // pair / CTA index
    constexpr uint32_t kTileRows = kTcM * CTAS;
    const uint32_t n_tiles = (row_end - row_begin + kTileRows - 1) / kTileRows;
    const uint32_t n_kb = (p.dim + kTcKBytes - 1) / kTcKBytes;
    const uint32_t n_b = p.n_pad / CTAS;           // query rows this CTA stages per k-block
    const uint32_t b_bytes = (n_b * kTcKBytes + 1023u) & ~1023u;
    const bool resident = n_kb * b_bytes + 5u * kTcABytes <= kTcRingBytes;
    const uint32_t kStages =
        resident ? min(kTcMaxStages, (kTcRingBytes - n_kb * b_bytes) / kTcABytes) : kStagesStream;
    const uint32_t kStageBytes = resident ? kTcABytes : kTcRingBytes / kStagesStream;
    uint8_t *ring_a = ring + (resident ? n_kb * b_bytes : 0u);  // resident queries come first

    if (tid == 0) {
        for (uint32_t s = 0; s < kTcMaxStages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        mbar_init(bfull_bar, 1);
        for (uint32_t a = 0; a < 2; ++a) {
            mbar_init(&tfull_bar[a], 1);
            // 1 CTA: one arrive per epilogue warp.  Pair: one per CTA on the leader's barrier (a
            // remote arrive drags a MEMBAR.GPU along, so the warps first meet on a named barrier)
            mbar_init(&tempty_bar[a], CTAS == 2 ? 2u : kTcEpilogueWarps);
        }
        fence_mbar_init();
    }
    if (warp == 1) {
        if (CTAS == 2) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                             smem_u32(tmem_base_s)),
                         "r"(kTcTmemCols)
                         : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                             smem_u32(tmem_base_s)),
                         "r"(kTcTmemCols)
                         : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    {
        // the screen compares in units of 2^shift (exact power-of-two scaling) and with the
        // int -> float conversion constant 1.5 * 2^23 folded into the query's additive term
        const float sc = __uint_as_float((127u - p.shift) << 23);
        for (uint32_t i = tid; i < kTcMaxQ; i += kTcThreads) {
            float4 c = make_float4(0.0f, 0.0f, INFINITY, 0.0f);  // padding: never passes
            if (i < p.nq) {
                c = p.coef[i];
                c.x *= sc;
                c.y *= sc;
                c.z = __double2float_rd((double)c.z * (double)sc + 12582912.0);
            }
            coef_s[i] = c;
        }
        if (tid < kTcEpilogueWarps) pend_cnt_s[tid] = 0u;
        for (uint32_t i = tid; i < p.nq * 3u; i += kTcThreads)  // 48-byte records
            reinterpret_cast<uint4 *>(qm_s)[i] = reinterpret_cast<const uint4 *>(p.qmeta)[i];
    }
    __syncthreads();
    if (tid < kTcMaxQ / 16u) {
        // coarse screen of a whole 16-query chunk: alpha, beta >= 0, so the smallest threshold of
        // the chunk is at least alpha min(w) + beta min(u) + min(v)
        float4 mn = make_float4(INFINITY, INFINITY, INFINITY, 0.0f);
        for (uint32_t j = 0; j < 16u && tid * 16u + j < p.nq; ++j) {
            const float4 c = coef_s[tid * 16u + j];
            mn.x = fminf(mn.x, c.x);
            mn.y = fminf(mn.y, c.y);
            mn.z = fminf(mn.z, c.z);
        }
        cmin_s[tid] = mn;
    }
    tc_fence_before();
    if (CTAS == 2) cluster_sync_all();  // barriers initialised in both CTAs before any remote use
    else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_base_s;
    // the leader's barriers as seen from this CTA (identity for the leader / 1-CTA)
    const uint32_t lead_full0 = CTAS == 2 ? mapa_shared(smem_u32(&full_bar[0]), 0u) : 0u;
    const uint32_t lead_tempty0 = CTAS == 2 ? mapa_shared(smem_u32(&tempty_bar[0]), 0u) : 0u;

    if (warp == 0) {
        // ===== TMA producer: this CTA's corpus rows and its share of the queries =====
        if (lane == 0) {
            const uint64_t pol_a = p.evict_first ? policy_evict_first() : policy_evict_normal();
            const uint64_t pol_q = policy_evict_normal();
            const uint32_t tx = (kTcABytes + (resident ? 0u : n_b * kTcKBytes)) * CTAS;  // pair: both
            const uint32_t lead_bfull = CTAS == 2 ? mapa_shared(smem_u32(bfull_bar), 0u) : 0u;
            if (resident) {  // this CTA's share of the queries, once
                if (cr == 0u) mbar_arrive_expect_tx(bfull_bar, n_kb * n_b * kTcKBytes * CTAS);
                for (uint32_t kb = 0; kb < n_kb; ++kb) {
                    if (CTAS == 2)
                        tma_load_2d_pair(ring + kb * b_bytes, &tmap_q, (int32_t)(kb * kTcKBytes),
                                         (int32_t)(cr * n_b), lead_bfull, pol_q);
                    else
                        tma_load_2d(ring + kb * b_bytes, &tmap_q, (int32_t)(kb * kTcKBytes), 0,
                                    bfull_bar, pol_q);
                }
            }
            uint32_t stage = 0, phase = 0;
            for (uint32_t t = unit; t < n_tiles; t += n_units) {
                const int32_t row0 = (int32_t)(row_begin + t * kTileRows + cr * kTcM);
                for (uint32_t kb = 0; kb < n_kb; ++kb) {
                    mbar_wait_wd<20>(&empty_bar[stage], phase ^ 1u);
                    uint8_t *sa = ring_a + stage * kStageBytes;
                    if (CTAS == 2) {
                        if (cr == 0u) mbar_arrive_expect_tx(&full_bar[stage], tx);
                        const uint32_t fb = lead_full0 + stage * 8u;
                        tma_load_2d_pair(sa, &tmap_a, (int32_t)(kb * kTcKBytes), row0, fb, pol_a);
                        if (!resident)
                            tma_load_2d_pair(sa + kTcABytes, &tmap_q, (int32_t)(kb * kTcKBytes),
                                             (int32_t)(cr * n_b), fb, pol_q);
                    } else {
                        mbar_arrive_expect_tx(&full_bar[stage], tx);
                        tma_load_2d(sa, &tmap_a, (int32_t)(kb * kTcKBytes), row0, &full_bar[stage],
                                    pol_a);
                        if (!resident)
                            tma_load_2d(sa + kTcABytes, &tmap_q, (int32_t)(kb * kTcKBytes), 0,
                                        &full_bar[stage], pol_q);
                    }
                    if (++stage == kStages) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===== MMA issuer: one thread (of the leader CTA) =====
        if (lane == 0 && cr == 0u) {
            const uint32_t idesc = tc_idesc_i8(kTcM * CTAS, p.n_pad);
            if (resident) mbar_wait_wd<0>(bfull_bar, 0u);
            uint32_t stage = 0, phase = 0, it = 0;
            for (uint32_t t = unit; t < n_tiles; t += n_units, ++it) {
                const uint32_t acc = it & 1u, aph = (it >> 1) & 1u;
                mbar_wait_wd<32>(&tempty_bar[acc], aph ^ 1u);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * kTcMaxQ;
                for (uint32_t kb = 0; kb < n_kb; ++kb) {
                    mbar_wait_wd<0>(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(ring_a + stage * kStageBytes);
                    const uint64_t adesc = tc_smem_desc(sa);
                    const uint64_t bdesc =
                        tc_smem_desc(resident ? smem_u32(ring + kb * b_bytes) : sa + kTcABytes);
#pragma unroll
                    for (uint32_t ks = 0; ks < kTcKBytes / 32u; ++ks) {  // UMMA K = 32 int8 = 32 B
                        if (CTAS == 2)
                            tc_mma_i8_pair(d_tmem, adesc + 2ull * ks, bdesc + 2ull * ks, idesc,
                                           (kb | ks) != 0u ? 1u : 0u);
                        else
                            tc_mma_i8(d_tmem, adesc + 2ull * ks, bdesc + 2ull * ks, idesc,
                                      (kb | ks) != 0u ? 1u : 0u);
                    }
                    // frees the stage (in both CTAs) once these MMAs retire
                    if (CTAS == 2) tc_commit_pair(&empty_bar[stage]);
                    else tc_commit(&empty_bar[stage]);
                    if (++stage == kStages) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
                if (CTAS == 2) tc_commit_pair(&tfull_bar[acc]);
                else tc_commit(&tfull_bar[acc]);
            }
        }
        __syncwarp();
    } else {
        // ===== epilogue: warp w owns TMEM lanes 32 (w % 4) .. +31; the two warps of a lane
        //       quarter split the query columns =====
        const uint32_t qd = warp & 3u;
        const uint32_t part = (warp - kTcRoleWarps) >> 2;  // chunks part, part + kTcColParts, ...
        const uint32_t n_chunks = (p.nq + 15u) / 16u;
        const float sc = __uint_as_float((127u - p.shift) << 23);
        const int sh = (int)p.shift;
        const bool phase0 = row_begin == 0u && p.row_mask == nullptr;
        const TcRowConsts rcst = tc_row_consts(p.metric, p.dim);
        TcPend *pq = pend_s + (warp - kTcRoleWarps) * kTcPendCap;
        uint32_t *pcnt = pend_cnt_s + (warp - kTcRoleWarps);
        auto load_meta = [&](uint32_t r, bool ok) {
            float4 raw = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            if (ok) raw = __ldg(reinterpret_cast<const float4 *>(p.meta + r));
            return raw;
        };
        auto drain = [&]() {  // warp-uniform: one parked entry per lane and round
            __syncwarp();
            const uint32_t n_pend = min(*pcnt, kTcPendCap);
            for (uint32_t e = lane; e < n_pend; e += 32u) {
                const TcPend pe = pq[e];
                const float4 raw = __ldg(reinterpret_cast<const float4 *>(p.meta + pe.row));
                RowMeta rm;
                rm.scale = raw.x;
                rm.x1 = __float_as_uint(raw.y);
                rm.rmag = raw.z;
                rm.flags = __float_as_uint(raw.w);
                tc_keep_entry(p, qm_s, pe.q, pe.I, pe.row, rm, __float_as_uint(coef_s[pe.q].w), phase0);
            }
            __syncwarp();
            if (lane == 0) *pcnt = 0u;
            __syncwarp();
        };
        uint32_t it = 0;
        float4 raw_next = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        if (unit < n_tiles) {
            const uint32_t r0 = row_begin + unit * kTileRows + cr * kTcM + qd * 32u + lane;
            raw_next = load_meta(r0, r0 < row_end);
        }
        for (uint32_t t = unit; t < n_tiles; t += n_units, ++it) {
            const uint32_t acc = it & 1u, aph = (it >> 1) & 1u;
            const uint32_t row = row_begin + t * kTileRows + cr * kTcM + qd * 32u + lane;
            bool valid = row < row_end;
            if (valid && p.row_mask) valid = ((__ldg(p.row_mask + (row >> 5)) >> (row & 31u)) & 1u) != 0u;
            const float4 raw = raw_next;
            {   // next tile's row constants: in flight while this tile is screened
                const uint32_t tn = t + n_units;
                const uint32_t rn = row_begin + tn * kTileRows + cr * kTcM + qd * 32u + lane;
                raw_next = load_meta(rn, tn < n_tiles && rn < row_end);
            }
            RowMeta m;
            m.scale = raw.x;
            m.x1 = __float_as_uint(raw.y);
            m.rmag = raw.z;
            m.flags = __float_as_uint(raw.w);
            float alpha, beta, br;
            tc_row_coef(p.metric, m, rcst, alpha, beta, br);
            if (!p.screen) br = INFINITY;
            // lhs = float((I >> shift) + ceil(br / 2^shift) + 5) via the 1.5 * 2^23 bit trick (the
            // 5 covers the floor of the shift and the roundings of the two fma below, whose
            // results are near 1.26e7 where one ulp is 1); rows that must always be kept get
            // a huge lhs
            const float brs = __fmul_ru(br, sc);
            const int add_r = (brs < 2000000.0f) ? 0x4B400000 + (__float2int_ru(brs) + 5) : 0x7f000000;
            if (lane == 0) mbar_wait_wd<64>(&tfull_bar[acc], aph);
            __syncwarp();
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((qd * 32u) << 16) + acc * kTcMaxQ;
            auto process = [&](int (&v)[16], uint32_t c0) {
                // coarse: the largest dot product of the chunk against its smallest threshold
                int mx = max(max(max(v[0], v[1]), max(v[2], v[3])), max(max(v[4], v[5]), max(v[6], v[7])));
                mx = max(mx, max(max(max(v[8], v[9]), max(v[10], v[11])),
                                 max(max(v[12], v[13]), max(v[14], v[15]))));
                const float4 cm = cmin_s[c0 >> 4];
                const float lhs_c = __int_as_float((mx >> sh) + add_r);
                const float rhs_c = fmaf(alpha, cm.x, fmaf(beta, cm.y, cm.z));
                if (!__any_sync(0xffffffffu, valid && !(lhs_c < rhs_c))) return;
                // fine: every column against its own threshold
                uint32_t mask = 0u;
#pragma unroll
                for (uint32_t j = 0; j < 16u; ++j) {
                    const float4 cq = coef_s[c0 + j];
                    const float lhs = __int_as_float((v[j] >> sh) + add_r);
                    const float rhs = fmaf(alpha, cq.x, fmaf(beta, cq.y, cq.z));
                    if (!(lhs < rhs)) mask |= 1u << j;
                }
                {
                    const uint32_t left = p.nq - c0;  // c0 < nq
                    mask &= left >= 16u ? 0xffffu : ((1u << left) - 1u);
                    if (!valid) mask = 0u;
                }
                // Park the hits.  The reservation is warp-cooperative (one entry per lane and
                // round, slots handed out by ballot) and a full queue is drained by the WHOLE warp
                // right here: in the early phases, where a few per cent of all entries pass, the
                // rigorous evaluation then still runs 32 lanes wide instead of lane by lane.
                uint32_t any = __ballot_sync(0xffffffffu, mask != 0u);
                while (any) {
                    const uint32_t n_new = __popc(any);
                    uint32_t cnt = *pcnt;
                    if (cnt + n_new > kTcPendCap) {
                        drain();
                        cnt = 0u;
                    }
                    if (mask) {
                        const uint32_t j = __ffs(mask) - 1u;
                        mask &= mask - 1u;
                        int I = 0;
#pragma unroll
                        for (uint32_t jj = 0; jj < 16u; ++jj)
                            if (jj == j) I = v[jj];
                        TcPend e;
                        e.I = I;
                        e.q = c0 + j;
                        e.row = row;
                        pq[cnt + __popc(any & ((1u << lane) - 1u))] = e;
                    }
                    __syncwarp();
                    if (lane == 0) *pcnt = cnt + n_new;
                    __syncwarp();
                    any = __ballot_sync(0xffffffffu, mask != 0u);
                }
            };
            if (p.dump) {  // diagnostics only (nm_debug_tc_dots)
                for (uint32_t c = 0; c < p.nq; ++c) {
                    const int I = tc_ld1(taddr + c);
                    if (valid && ((c >> 4) % kTcColParts) == part)
                        p.dump[(size_t)c * p.dump_stride + row] = I;
                }
            }
            if (part < n_chunks) {
                int va[16], vb[16];
                tc_ld16_issue(taddr + part * 16u, va);
                tc_ld_wait(va);
                for (uint32_t ch = part; ch < n_chunks; ch += 2u * kTcColParts) {
                    const uint32_t chb = ch + kTcColParts, chn = ch + 2u * kTcColParts;
                    const bool has_b = chb < n_chunks;
                    if (has_b) tc_ld16_issue(taddr + chb * 16u, vb);
                    process(va, ch * 16u);
                    if (has_b) {
                        tc_ld_wait(vb);
                        if (chn < n_chunks) tc_ld16_issue(taddr + chn * 16u, va);
                        process(vb, chb * 16u);
                        if (chn < n_chunks) tc_ld_wait(va);
                    }
                }
            }
            tc_fence_before();
            if (CTAS == 2) {
                asm volatile("bar.sync 2, %0;" ::"n"(32 * kTcEpilogueWarps) : "memory");
                if (warp == kTcRoleWarps && lane == 0) mbar_arrive_cluster(lead_tempty0 + acc * 8u);
            } else {
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty_bar[acc]);
            }
            // the accumulator is back with the MMA issuer.  Parked entries are evaluated one per
            // lane once enough of them have gathered (the latency of a batch -- query constants,
            // double precision, a list reservation in global memory -- is the same for 1 or 32)
            if (*pcnt >= kTcPendDrain) drain();
        }
        drain();
    }
    tc_fence_before();
    if (CTAS == 2) cluster_sync_all();  // no CTA may leave while its peer still signals it
    else __syncthreads();
    tc_fence_after();
    if (warp == 1) {
        if (CTAS == 2)
            asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                         "r"(kTcTmemCols)
                         : "memory");
        else
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                         "r"(kTcTmemCols)
                         : "memory");
    }
}

// ---------------------------------------------------------------------------------------
// refine: per query, tighten tau and compact the kept list
// ---------------------------------------------------------------------------------------
// tau must be a score that at least k rows reach.  The k-th best LOWER bound is one, but it
// sits a whole interval width below the true k-th best score, which lets ~e^(lambda width)
// times too many entries through the screen.  So the k entries with the best lower bounds are
// re-scored exactly here (reference arithmetic, from the f32 mirror): the k-th best of their
// exact scores is also reached by k rows and is, up to the quantisation noise, the true k-th
// best score of the rows seen so far.  Re-scored entries keep lb = ub = exact score.
struct TcRefineParams {
    TcKept *kept;
    uint32_t *kept_n;
    uint32_t *kept_prev;   // [nq] list length after the previous refine
    TcQueryMeta *qmeta;
    float4 *coef;
    TcCtl *ctl;
    const float *queries;  // [nq][dim]
    const float *rows;     // f32 mirror
    uint32_t pitch;
    uint32_t n_rows;
    uint32_t dim;
    uint32_t k;
    int metric;
    uint32_t grow0;      // row-range growth after the first phase (its pass rate says nothing)
    uint32_t allow_mul;  // the pass rate is assumed to drop at least this much from phase to phase
};

constexpr uint32_t kTcTopCap = 2048;  // >= kMaxFastK; entries re-scored per refine at most

// k-th largest of val(0..n) (n >= k >= 1), MSB-first radix select; all 256 threads call.
// Thread t owns digit bin t; the bin holding the k-th largest is found with a parallel suffix
// sum (warp shuffles + 8 warp totals), not a serial walk over the 256 bins.
template <class F>
__device__ __forceinline__ uint32_t tc_radix_kth(F val, uint32_t n, uint32_t k, uint32_t *hist,
                                                 uint32_t *sel_s, uint32_t t) {
    __shared__ uint32_t wtot[8];
    const uint32_t lane = t & 31u, warp = t >> 5;
    if (t == 0) {
        sel_s[0] = 0u;
        sel_s[1] = k;
    }
    for (int shift = 24; shift >= 0; shift -= 8) {
        hist[t] = 0u;
        __syncthreads();
        const uint32_t prefix = sel_s[0], need = sel_s[1];
        // four independent loads per thread and round: the lists live in L2 / HBM and the loop is
        // latency-, not bandwidth-bound
        for (uint32_t base = 0; base < n; base += 1024u) {
            uint32_t o[4];
#pragma unroll
            for (uint32_t j = 0; j < 4u; ++j) {
                const uint32_t i = base + j * 256u + t;
                o[j] = i < n ? val(i) : 0u;
            }
#pragma unroll
            for (uint32_t j = 0; j < 4u; ++j) {
                const uint32_t i = base + j * 256u + t;
                if (i < n && (shift == 24 || (o[j] >> (shift + 8)) == prefix))
                    atomicAdd(&hist[(o[j] >> shift) & 255u], 1u);
            }
        }
        __syncthreads();
        const uint32_t mine = hist[t];
        uint32_t sfx = mine;  // sum of bins t .. end of this warp
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const uint32_t v = __shfl_down_sync(0xffffffffu, sfx, off);
            if (lane + off < 32u) sfx += v;
        }
        if (lane == 0) wtot[warp] = sfx;
        __syncthreads();
        for (uint32_t w = warp + 1; w < 8u; ++w) sfx += wtot[w];  // bins t .. 255
        const uint32_t above = sfx - mine;                         // bins t+1 .. 255
        if (above < need && need <= sfx) {                         // exactly one thread
            sel_s[0] = (prefix << 8) | t;
            sel_s[1] = need - above;
        }
        __syncthreads();
    }
    return sel_s[0];
}

// One thread scores one survivor row with the reference arithmetic: 128-byte blocks of the row
// are fetched as 8 independent float4 loads, the next block is in flight while the current one
// is folded (the rows are scattered, so the loop is bound by HBM latency otherwise); the query
// comes from shared memory as broadcast float4.  Same element order as scan_topk_kernel.
template <int METRIC, uint32_t BLK = 8>
__device__ __forceinline__ float tc_score_row(const float *q_s, const float *__restrict__ x,
                                              uint32_t dim, float qmag) {
    RowAcc<METRIC> acc;
    acc.reset();
    const uint32_t full = (dim / 8u) * 8u;  // elements in whole f32x8 groups
    const uint32_t n4 = full / 4u;
    const float4 *xv = reinterpret_cast<const float4 *>(x);
    const float4 *qv = reinterpret_cast<const float4 *>(q_s);
    float4 cur[BLK], nxt[BLK];
#pragma unroll
    for (uint32_t j = 0; j < BLK; ++j)
        cur[j] = (j < n4) ? __ldg(xv + j) : make_float4(0.f, 0.f, 0.f, 0.f);
    for (uint32_t b = 0; b < n4; b += BLK) {
#pragma unroll
        for (uint32_t j = 0; j < BLK; ++j)
            nxt[j] = (b + BLK + j < n4) ? __ldg(xv + b + BLK + j) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (uint32_t j = 0; j < BLK; ++j) {
            if (b + j < n4) {
                const float4 qq = qv[b + j];
                if (j & 1u) acc.template step<1>(cur[j], qq);
                else acc.template step<0>(cur[j], qq);
            }
        }
#pragma unroll
        for (uint32_t j = 0; j < BLK; ++j) cur[j] = nxt[j];
    }
    if (METRIC == kEuclidean) {
        float sum = acc.d[0];
        for (uint32_t i = full; i < dim; ++i) {
            const float df = __fsub_rn(q_s[i], x[i]);
            sum = __fadd_rn(sum, __fmul_rn(df, df));
        }
        return tc_l2_score(sum);
    }
    float dot = fold_lanes(acc.d), ssq = fold_lanes(acc.s);
    for (uint32_t i = full; i < dim; ++i) {
        const float xi = x[i];
        dot = __fadd_rn(dot, __fmul_rn(q_s[i], xi));
        if (METRIC == kCosine) ssq = __fadd_rn(ssq, __fmul_rn(xi, xi));
    }
    if (METRIC == kDot) return dot;
    const float rmag = __fsqrt_rn(ssq);
    return (qmag == 0.0f || rmag == 0.0f) ? 0.0f : __fdiv_rn(dot, __fmul_rn(qmag, rmag));
}

__global__ void __launch_bounds__(256) tc_refine_kernel(const TcRefineParams p) {
    extern __shared__ __align__(16) float q_s[];  // [dim]
    __shared__ uint32_t hist[256];
    __shared__ uint32_t sel_s[2];
    __shared__ uint32_t warp_cnt[8];
    __shared__ uint32_t out_pos_s, top_n;
    __shared__ uint32_t top_idx[kTcTopCap];
    __shared__ uint32_t top_ord[kTcTopCap];   // first the rows (for the prefetch), then the exact ords
    const uint32_t q = blockIdx.x, t = threadIdx.x, lane = t & 31u, warp = t >> 5;
    const uint32_t row_begin = p.ctl->row_begin, row_end = min(p.ctl->row_end, p.n_rows);
    if (row_begin >= row_end) return;  // no phase ran before this launch
    TcKept *list = p.kept + (size_t)q * kTcKeptCap;
    const uint32_t n_raw = p.kept_n[q];
    const uint32_t n = min(n_raw, kTcKeptCap);
    uint32_t tau = p.qmeta[q].tau_ord;
    if (n >= p.k) {
        const uint32_t sel = tc_radix_kth([&](uint32_t i) { return list[i].lb_ord; }, n, p.k, hist,
                                          sel_s, t);
        tau = max(tau, sel);
        // the entries with the k best lower bounds (ties may add a few)
        if (t == 0) top_n = 0u;
        const float *qv = p.queries + (size_t)q * p.dim;
        for (uint32_t i = t; i < p.dim; i += 256u) q_s[i] = __ldg(qv + i);
        __syncthreads();
        for (uint32_t base = 0; base < n; base += 1024u) {
            TcKept e[4];
#pragma unroll
            for (uint32_t j = 0; j < 4u; ++j) {
                const uint32_t i = base + j * 256u + t;
                e[j].row = e[j].lb_ord = e[j].ub_ord = 0u;
                if (i < n) e[j] = list[i];
            }
#pragma unroll
            for (uint32_t j = 0; j < 4u; ++j) {
                const uint32_t i = base + j * 256u + t;
                if (i < n && e[j].lb_ord >= sel) {
                    const uint32_t pos = atomicAdd(&top_n, 1u);
                    if (pos < kTcTopCap) {
                        top_idx[pos] = i;
                        top_ord[pos] = (e[j].lb_ord != e[j].ub_ord) ? e[j].row : 0xffffffffu;  // exact already
                    }
                }
            }
        }
        __syncthreads();
        const uint32_t cnt = top_n;
        if (cnt <= kTcTopCap) {
            // The k threads below each walk one scattered row as a chain of dependent 128-byte
            // steps: pull all of those rows into L2 first, with every thread of the CTA issuing
            // prefetches, so the walk pays L2 instead of HBM latency per step.
            {
                const uint32_t lines = (p.dim * 4u + 127u) / 128u;
                for (uint32_t w = t; w < cnt * lines; w += 256u) {
                    const uint32_t j = w / lines, seg = w - j * lines;
                    const uint32_t row = top_ord[j];
                    if (row != 0xffffffffu)
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(p.rows + (size_t)row * p.pitch + seg * 32u));
                }
            }
            __syncthreads();
            const float qmag = p.qmeta[q].qmag;
            for (uint32_t j = t; j < cnt; j += 256u) {
                const uint32_t i = top_idx[j];
                TcKept e = list[i];
                if (e.lb_ord != e.ub_ord) {
                    const float *x = p.rows + (size_t)e.row * p.pitch;
                    float sc;
                    if (p.metric == kEuclidean) sc = tc_score_row<kEuclidean>(q_s, x, p.dim, qmag);
                    else if (p.metric == kCosine) sc = tc_score_row<kCosine>(q_s, x, p.dim, qmag);
                    else sc = tc_score_row<kDot>(q_s, x, p.dim, qmag);
                    e.lb_ord = e.ub_ord = score_to_ord(__float_as_uint(sc));
                    list[i] = e;
                }
                top_ord[j] = e.lb_ord;
            }
            __syncthreads();
            const uint32_t ex = tc_radix_kth([&](uint32_t i) { return top_ord[i]; }, cnt, p.k, hist,
                                            sel_s, t);
            tau = max(tau, ex);
        }
    }
    // in-place stable compaction, 1024 entries per round (four independent loads per thread; the
    // writes of a round never pass the entries it has read)
    if (t == 0) out_pos_s = 0u;
    __syncthreads();
    for (uint32_t base = 0; base < n; base += 1024u) {
        TcKept e[4];
        bool keep[4];
        uint32_t mine = 0;
#pragma unroll
        for (uint32_t j = 0; j < 4u; ++j) {
            const uint32_t i = base + t * 4u + j;  // thread t owns 4 consecutive entries: order kept
            e[j].row = e[j].lb_ord = e[j].ub_ord = 0u;
            if (i < n) e[j] = list[i];
            keep[j] = i < n && e[j].ub_ord >= tau;
            mine += keep[j] ? 1u : 0u;
        }
        // exclusive prefix of `mine` over the 256 threads
        uint32_t incl = mine;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const uint32_t v = __shfl_up_sync(0xffffffffu, incl, off);
            if (lane >= (uint32_t)off) incl += v;
        }
        if (lane == 31u) warp_cnt[warp] = incl;
        __syncthreads();  // (also: every thread has read its entries before anyone writes)
        uint32_t off = out_pos_s + incl - mine;
        for (uint32_t w = 0; w < warp; ++w) off += warp_cnt[w];
#pragma unroll
        for (uint32_t j = 0; j < 4u; ++j)
            if (keep[j]) list[off++] = e[j];
        __syncthreads();
        if (t == 0) {
            uint32_t tot = 0;
            for (uint32_t w = 0; w < 8; ++w) tot += warp_cnt[w];
            out_pos_s += tot;
        }
        __syncthreads();
    }
    if (t == 0) {
        const uint32_t m = out_pos_s;
        p.kept_n[q] = m;
        TcQueryMeta qm = p.qmeta[q];
        if (n_raw > kTcKeptCap) qm.flags |= kTcFlagOverflow;
        qm.tau_ord = tau;
        p.qmeta[q] = qm;
        p.coef[q] = tc_make_coef(p.metric, tau, qm, p.dim);
        // rows the next phase may cover so that this query's list stays half empty even if the
        // pass rate does not drop (it does: tau has just tightened)
        const uint32_t prev = p.kept_prev[q];
        p.kept_prev[q] = m;
        if (!(qm.flags & (kTcFlagUnusable | kTcFlagOverflow))) {
            const uint64_t added = n > prev ? n - prev : 1u;
            const uint64_t room = (kTcKeptCap - m) / 2u;
            const uint64_t allowed = (uint64_t)p.allow_mul * room * (uint64_t)(row_end - row_begin) / added;
            atomicMin(&p.ctl->next_rows, allowed < 0xfffffff0ull ? (uint32_t)allowed : 0xfffffff0u);
        }
        __threadfence();
        if (atomicAdd(&p.ctl->ticket, 1u) == gridDim.x - 1) {
            __threadfence();
            unsigned long long span = *reinterpret_cast<volatile uint32_t *>(&p.ctl->next_rows);
            // the first phase keeps everything, so its pass rate says nothing: grow 4x;
            // afterwards at least 8 Ki rows and at most 15x what has been seen so far
            if (row_begin == 0u) span = (unsigned long long)p.grow0 * row_end;
            if (span < 8192ull) span = 8192ull;
            if (span > 15ull * row_end) span = 15ull * row_end;
            const unsigned long long nb = row_end;
            unsigned long long ne = (nb + span + kTcM - 1) / kTcM * kTcM;
            if (ne > p.n_rows) ne = p.n_rows;
            p.ctl->row_begin = (uint32_t)nb;
            p.ctl->row_end = (uint32_t)ne;
            p.ctl->next_rows = 0xffffffffu;
            p.ctl->ticket = 0u;
            p.ctl->phases += 1u;
        }
    }
}

// ---------------------------------------------------------------------------------------
// exact re-score of the survivors + final selection: one CTA per query
// ---------------------------------------------------------------------------------------
struct TcRescoreParams {
    const float *queries;     // [nq][dim]
    const float *rows;        // f32 mirror
    const TcKept *kept;
    const uint32_t *kept_n;
    const TcQueryMeta *qmeta;
    uint64_t *exact_keys;     // [nq][kTcKeptCap]
    uint64_t *out_rows;       // [nq][out_stride]
    float *out_scores;
    uint32_t *out_counts;
    uint32_t *stats;          // [0] total survivors re-scored (atomic)
    uint64_t row_base;
    uint32_t pitch;
    uint32_t dim;
    uint32_t k;
    uint32_t out_stride;
    int metric;
};

// final selection of one query over its exact keys (all 256 threads of the CTA)
__device__ __forceinline__ void tc_select_query(const TcRescoreParams &p, uint32_t q, uint32_t n,
                                                uint64_t *buf, uint64_t *thr_s, uint32_t *cnt_s,
                                                uint32_t *hist, uint32_t t) {
    const uint64_t *keys = p.exact_keys + (size_t)q * kTcKeptCap;
    if (t == 0 && p.stats) atomicAdd(p.stats, n);
    TopKState st;
    st.buf = buf;
    st.cnt_smem = cnt_s;
    st.thr_smem = thr_s;
    st.count = 0;
    st.k = p.k;
    st.cap = kCandCap;
    MergeScratch ms;
    ms.hist = hist;
    ms.sc = hist + 256;
    merge_published(st, t, keys, n, p.k, ms);
    TopKOutputs o;
    o.out_keys = nullptr;
    o.out_hits = nullptr;
    o.out_rows = p.out_rows + (size_t)q * p.out_stride;
    o.out_scores = p.out_scores + (size_t)q * p.out_stride;
    o.out_count = p.out_counts + q;
    o.row_base = p.row_base;
    o.accumulate_count = 0;
    write_outputs(st, t, p.k, o);
}

// Per-query re-score (query in shared memory, rows in list order).  Only used when the query
// rows are not 16-byte aligned (dim % 4 != 0); see tc_score_sorted_kernel for the fast path.
__global__ void __launch_bounds__(kRowsPerBlock) tc_rescore_kernel(const TcRescoreParams p) {
    extern __shared__ __align__(16) float q_s[];  // [dim]
    __shared__ __align__(16) uint64_t buf[kCandCap];
    __shared__ uint64_t thr_s;
    __shared__ uint32_t cnt_s;
    __shared__ uint32_t hist[256 + 16];
    const uint32_t q = blockIdx.x, t = threadIdx.x;
    const float *qv = p.queries + (size_t)q * p.dim;
    for (uint32_t i = t; i < p.dim; i += kRowsPerBlock) q_s[i] = __ldg(qv + i);
    if (t == 0) {
        thr_s = 0ull;
        cnt_s = 0u;
    }
    __syncthreads();
    const uint32_t n = min(p.kept_n[q], kTcKeptCap);
    const float qmag = p.qmeta[q].qmag;
    const TcKept *list = p.kept + (size_t)q * kTcKeptCap;
    uint64_t *keys = p.exact_keys + (size_t)q * kTcKeptCap;
    for (uint32_t i = t; i < n; i += kRowsPerBlock) {
        const uint32_t row = list[i].row;
        const float *x = p.rows + (size_t)row * p.pitch;
        float s;
        if (p.metric == kEuclidean) s = tc_score_row<kEuclidean>(q_s, x, p.dim, qmag);
        else if (p.metric == kCosine) s = tc_score_row<kCosine>(q_s, x, p.dim, qmag);
        else s = tc_score_row<kDot>(q_s, x, p.dim, qmag);
        keys[i] = make_key(__float_as_uint(s), row);
    }
    __threadfence();
    __syncthreads();
    tc_select_query(p, q, n, buf, &thr_s, &cnt_s, hist, t);
}

// ---- fast path: survivors of ALL queries re-scored in corpus order --------------------------
// The survivors of one query are spread over the whole mirror (tens of GB): co-resident threads
// that each walk a different far-away row touch hundreds of pages per SM, and the gather runs
// at ~1.2 TB/s (scripts/mb_gather.cu: address translation, not DRAM, is the limit).  Ordered by
// row, neighbouring threads walk neighbouring rows and the same gather reaches ~4.3 TB/s.  So:
// counting sort of the (row, query, slot) triples by row bucket, then one thread per sorted
// entry (its query read through L1/L2), then the per-query selection.
struct TcSortParams {
    const TcKept *kept;
    const uint32_t *kept_n;
    uint32_t *bucket;      // [n_buckets + 1] counts, then cursors
    uint2 *sorted;         // {row, query << 16 | slot}
    uint32_t *total;       // [1]
    uint32_t n_buckets;
    uint32_t shift;        // bucket = row >> shift
};

__global__ void __launch_bounds__(256) tc_sort_count_kernel(const TcSortParams p) {
    const uint32_t q = blockIdx.x;
    const uint32_t n = min(p.kept_n[q], kTcKeptCap);
    const TcKept *list = p.kept + (size_t)q * kTcKeptCap;
    for (uint32_t i = threadIdx.x; i < n; i += 256u) atomicAdd(&p.bucket[list[i].row >> p.shift], 1u);
}

__global__ void __launch_bounds__(1024) tc_sort_scan_kernel(const TcSortParams p) {
    __shared__ uint32_t part[1024];
    const uint32_t t = threadIdx.x;
    const uint32_t per = (p.n_buckets + 1023u) / 1024u;
    const uint32_t lo = min(t * per, p.n_buckets), hi = min(lo + per, p.n_buckets);
    uint32_t sum = 0;
    for (uint32_t b = lo; b < hi; ++b) sum += p.bucket[b];
    part[t] = sum;
    __syncthreads();
    for (uint32_t off = 1; off < 1024u; off <<= 1) {  // inclusive scan of the per-thread sums
        const uint32_t v = t >= off ? part[t - off] : 0u;
        __syncthreads();
        part[t] += v;
        __syncthreads();
    }
    uint32_t run = part[t] - sum;
    for (uint32_t b = lo; b < hi; ++b) {
        const uint32_t c = p.bucket[b];
        p.bucket[b] = run;  // exclusive offset: the scatter kernel's cursor
        run += c;
    }
    if (t == 1023u) *p.total = part[1023];
}

__global__ void __launch_bounds__(256) tc_sort_scatter_kernel(const TcSortParams p) {
    const uint32_t q = blockIdx.x;
    const uint32_t n = min(p.kept_n[q], kTcKeptCap);
    const TcKept *list = p.kept + (size_t)q * kTcKeptCap;
    for (uint32_t i = threadIdx.x; i < n; i += 256u) {
        const uint32_t row = list[i].row;
        const uint32_t pos = atomicAdd(&p.bucket[row >> p.shift], 1u);
        p.sorted[pos] = make_uint2(row, (q << 16) | i);
    }
}

__global__ void __launch_bounds__(256) tc_score_sorted_kernel(const TcRescoreParams p,
                                                             const uint2 *__restrict__ sorted,
                                                             const uint32_t *__restrict__ total) {
    const uint32_t n = *total;
    for (uint32_t e = blockIdx.x * 256u + threadIdx.x; e < n; e += gridDim.x * 256u) {
        const uint2 ent = sorted[e];
        const uint32_t row = ent.x, q = ent.y >> 16, i = ent.y & 0xffffu;
        const float *x = p.rows + (size_t)row * p.pitch;
        const float *qv = p.queries + (size_t)q * p.dim;  // 16-byte aligned: dim % 4 == 0
        const float qmag = p.qmeta[q].qmag;
        float s;
        if (p.metric == kEuclidean) s = tc_score_row<kEuclidean>(qv, x, p.dim, qmag);
        else if (p.metric == kCosine) s = tc_score_row<kCosine>(qv, x, p.dim, qmag);
        else s = tc_score_row<kDot>(qv, x, p.dim, qmag);
        p.exact_keys[(size_t)q * kTcKeptCap + i] = make_key(__float_as_uint(s), row);
    }
}

__global__ void __launch_bounds__(kRowsPerBlock) tc_select_kernel(const TcRescoreParams p) {
    __shared__ __align__(16) uint64_t buf[kCandCap];
    __shared__ uint64_t thr_s;
    __shared__ uint32_t cnt_s;
    __shared__ uint32_t hist[256 + 16];
    const uint32_t q = blockIdx.x, t = threadIdx.x;
    if (t == 0) {
        thr_s = 0ull;
        cnt_s = 0u;
    }
    __syncthreads();
    tc_select_query(p, q, min(p.kept_n[q], kTcKeptCap), buf, &thr_s, &cnt_s, hist, t);
}

// Which queries of a call must be redone by the exact path, decided ON THE DEVICE (asynchronous
// searches): the query was unusable / its list overflowed, or the phases of its pass did not
// reach the end of the corpus.  Same rule as the host's tc_query_flags.
__global__ void tc_redo_flags_kernel(const TcQueryMeta *qmeta, const TcCtl *ctl, uint32_t nq,
                                     uint32_t n_rows, uint32_t *redo) {
    for (uint32_t q = blockIdx.x * blockDim.x + threadIdx.x; q < nq; q += gridDim.x * blockDim.x) {
        uint32_t f = qmeta[q].flags;
        if (ctl[q / kTcMaxQ].row_begin < n_rows) f |= 4u;
        redo[q] = f;
    }
}

// sharded indexes: a shard's result as ShardHit[nq, k] for the cross-shard merge
__global__ void __launch_bounds__(256)
tc_pack_hits_kernel(const uint64_t *__restrict__ rows, const float *__restrict__ scores,
                    const uint32_t *__restrict__ counts, uint32_t k, ShardHit *hits) {
    const uint32_t q = blockIdx.x, n = counts[q];
    for (uint32_t i = threadIdx.x; i < k; i += 256u) {
        ShardHit h;
        h.global_row = 0ull;
        h.ord = 0u;
        h.score_bits = 0u;
        if (i < n) {
            const uint32_t sb = __float_as_uint(scores[(size_t)q * k + i]);
            h.global_row = rows[(size_t)q * k + i];
            h.ord = score_to_ord(sb);
            h.score_bits = sb;
            // an empty slot is {0,0,0}; a NaN hit has ord 0 but non-zero score bits (as
            // write_outputs encodes it)
        }
        hits[(size_t)q * k + i] = h;
    }
}

#endif  // __CUDACC__
}  // namespace nm
