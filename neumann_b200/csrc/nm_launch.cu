// nm_launch.cu — per-call workspaces and EVERY kernel launch of the library.  This is the
// only translation unit that includes the kernel headers (scan / batch / prefilter).
#include "nm_internal.hpp"
#include "scan_kernels.cuh"
#include "batch_kernels.cuh"
#include "prefilter_kernels.cuh"
#include "tc_prefilter_kernels.cuh"
#include "filter_kernels.cuh"

#include <chrono>
#include <thread>

namespace nmi {

size_t scan_smem_bytes(uint32_t n_stages, uint32_t q_floats) {
    return 1024 + (size_t)n_stages * nm::kStageBytes + (size_t)nm::kCandCap * 8 +
           (size_t)q_floats * 4 + 2 * nm::kMaxStages * 8 + 64 + nm::kMaxStages * 4;
}

int ws_acquire(Shard &sh, std::unique_ptr<Workspace> &out) {
    {
        std::lock_guard<std::mutex> g(sh.pool_mu);
        if (!sh.pool.empty()) {
            out = std::move(sh.pool.back());
            sh.pool.pop_back();
            return NM_OK;
        }
    }
    std::unique_ptr<Workspace> ws(new Workspace());
    ws->device = sh.device;
    CUDA_TRY(cudaStreamCreateWithFlags(&ws->stream, cudaStreamNonBlocking));
    CUDA_TRY(cudaEventCreate(&ws->ev0));
    CUDA_TRY(cudaEventCreate(&ws->ev1));
    CUDA_TRY(cudaMalloc(&ws->d_counter, kWsCounterWords * sizeof(uint32_t)));
    CUDA_TRY(cudaMemsetAsync(ws->d_counter, 0, kWsCounterWords * sizeof(uint32_t), ws->stream));
    out = std::move(ws);
    return NM_OK;
}

void ws_release(Shard &sh, std::unique_ptr<Workspace> &ws) {
    if (!ws) return;
    std::lock_guard<std::mutex> g(sh.pool_mu);
    sh.pool.push_back(std::move(ws));
}

int wait_async_searches(nm_index *idx) {
    for (auto &shp : idx->shards) {
        Shard &sh = *shp;
        std::lock_guard<std::mutex> pg(sh.pool_mu);
        if (sh.stream_ws.empty()) continue;
        CUDA_TRY(cudaSetDevice(sh.device));
        for (auto &e : sh.stream_ws) {
            Workspace &ws = *e.second;
            if (ws.async_pending) {
                CUDA_TRY(cudaEventSynchronize(ws.async_done));
                ws.async_pending = false;
            }
            // pipelined scans record no event (it would sit between two launches and undo the
            // overlap); their last CTA publishes the finished sequence number instead
            for (uint32_t done = 0; ws.pipe_seq != 0;) {
                CUDA_TRY(cudaMemcpyAsync(&done, ws.d_counter + 4, sizeof(done), cudaMemcpyDeviceToHost,
                                         sh.copy_stream));
                CUDA_TRY(cudaStreamSynchronize(sh.copy_stream));
                if ((int32_t)(done - ws.pipe_seq) >= 0) break;
                std::this_thread::sleep_for(std::chrono::microseconds(50));
            }
        }
    }
    return NM_OK;
}

int ws_ensure(Workspace &ws, const Shard &sh, uint32_t dim, uint32_t nq, uint32_t k,
              bool need_query, bool need_result, bool need_hits, int gather_ranks) {
    size_t qf = (size_t)nq * dim;
    if (need_query && ws.query_cap < qf) {
        if (ws.d_query) CUDA_TRY(cudaFree(ws.d_query));
        if (ws.h_query) CUDA_TRY(cudaFreeHost(ws.h_query));
        ws.query_cap = 0;
        CUDA_TRY(cudaMalloc(&ws.d_query, qf * 4));
        CUDA_TRY(cudaMallocHost(&ws.h_query, qf * 4));
        ws.query_cap = qf;
    }
    // two sets of per-CTA candidate lists: pipelined scans alternate between them
    size_t cand = (size_t)sh.sm_count * std::min<uint32_t>(k, nm::kMaxFastK);
    if (ws.cand_cap < cand) {
        if (ws.d_cand) CUDA_TRY(cudaFree(ws.d_cand));
        ws.cand_cap = 0;
        CUDA_TRY(cudaMalloc(&ws.d_cand, 2 * cand * 8));
        ws.cand_cap = cand;
    }
    if (need_result) {
        ResultLayout l = result_layout(nq, k);
        if (ws.result_cap < l.total) {
            if (ws.d_result) CUDA_TRY(cudaFree(ws.d_result));
            if (ws.h_result) CUDA_TRY(cudaFreeHost(ws.h_result));
            ws.result_cap = 0;
            CUDA_TRY(cudaMalloc(&ws.d_result, l.total));
            CUDA_TRY(cudaMallocHost(&ws.h_result, l.total));
            ws.result_cap = l.total;
        }
    }
    if (need_hits) {
        size_t h = (size_t)nq * k;
        if (ws.hits_cap < h) {
            if (ws.d_hits) CUDA_TRY(cudaFree(ws.d_hits));
            if (ws.h_hits) CUDA_TRY(cudaFreeHost(ws.h_hits));
            ws.hits_cap = 0;
            CUDA_TRY(cudaMalloc(&ws.d_hits, h * sizeof(nm::ShardHit)));
            CUDA_TRY(cudaMallocHost(&ws.h_hits, h * sizeof(nm::ShardHit)));
            ws.hits_cap = h;
        }
        size_t g = h * (size_t)gather_ranks;
        if (gather_ranks > 0 && ws.gather_cap < g) {
            if (ws.d_gather) CUDA_TRY(cudaFree(ws.d_gather));
            ws.gather_cap = 0;
            CUDA_TRY(cudaMalloc(&ws.d_gather, g * sizeof(nm::ShardHit)));
            ws.gather_cap = g;
        }
    }
    return NM_OK;
}

template <int METRIC>
int launch_scan_t(const Shard &sh, const nm::ScanParams &p, size_t smem, cudaStream_t stream) {
    // The dynamic-smem opt-in is a per-function, per-device attribute shared by all host
    // threads: raise it once to the architectural maximum and never lower it.
    static std::mutex mu;
    static bool configured[64] = {false};
    auto kern = nm::scan_topk_kernel<METRIC>;
    {
        std::lock_guard<std::mutex> g(mu);
        if (sh.device >= 64 || !configured[sh.device]) {
            CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          227 * 1024));
            if (sh.device < 64) configured[sh.device] = true;
        }
    }
    uint32_t n_rb = (p.n_rows + nm::kRowsPerBlock - 1) / nm::kRowsPerBlock;
    uint32_t grid = std::min<uint32_t>((uint32_t)sh.sm_count, n_rb);
    if (p.pdl_seq) {
        // programmatic dependent launch: this scan may start while the previous kernel on the
        // stream (the previous pipelined scan) is still finishing
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(grid);
        cfg.blockDim = dim3(nm::kScanThreads);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, sh.tmap, p));
        return NM_OK;
    }
    kern<<<grid, nm::kScanThreads, smem, stream>>>(sh.tmap, p);
    CUDA_TRY(cudaGetLastError());
    return NM_OK;
}

// One query over one shard: a single launch for k <= 1024, otherwise ceil(k/1024) chained
// passes, each admitting only keys below the previous pass's last key.
int launch_scan(nm_index *idx, const Shard &sh, Workspace &ws, const float *d_query, uint32_t k,
                int metric, uint64_t row_base, uint64_t *out_rows, float *out_scores,
                uint32_t *out_count, nm::ShardHit *out_hits, cudaStream_t stream,
                const nm::PeerXchg *xchg, const uint32_t *d_row_mask, const uint32_t *d_gate) {
    nm::ScanParams p;
    memset(&p, 0, sizeof(p));
    if (xchg) p.xchg = *xchg;
    p.row_mask = d_row_mask;
    p.gate = d_gate;
    p.query = d_query;
    p.cand = ws.d_cand;
    p.done_counter = ws.d_counter;
    p.row_base = row_base;
    p.n_rows = (uint32_t)sh.rows;
    p.dim = idx->dim;
    p.q_floats = (idx->dim + 31u) & ~31u;
    // stream through L2 with evict_first unless the whole shard fits comfortably in L2
    p.evict_first = (sh.rows * idx->pitch * 4ull > (64ull << 20)) ? 1u : 0u;
    uint32_t stages = nm::kMaxStages;
    const size_t limit = 227 * 1024;
    while (stages > 2 && scan_smem_bytes(stages, p.q_floats) > limit) --stages;
    if (scan_smem_bytes(stages, p.q_floats) > limit)
        return fail(NM_ERR_DIMENSION_MISMATCH,
                    "dimension %u does not fit the scan kernel's shared-memory query buffer",
                    idx->dim);
    p.n_stages = stages;
    size_t smem = scan_smem_bytes(stages, p.q_floats);
    const bool chained = k > (uint32_t)nm::kMaxFastK;
    const bool pipelined = ws.pipeline_next && !chained && !d_row_mask && !d_gate;
    if (pipelined) {
        p.pdl_seq = ++ws.pipe_seq;
        p.pdl_done = ws.d_counter + 4;
        if (p.pdl_seq & 1u) {
            p.cand = ws.d_cand + ws.cand_cap;
            p.done_counter = ws.d_counter + 2;
        }
    }
    if (chained) {
        if (ws.pass_keys_cap < k) {
            if (ws.d_pass_keys) {
                CUDA_TRY(cudaStreamSynchronize(stream));
                CUDA_TRY(cudaFree(ws.d_pass_keys));
            }
            ws.pass_keys_cap = 0;
            CUDA_TRY(cudaMalloc(&ws.d_pass_keys, (size_t)k * 8));
            ws.pass_keys_cap = k;
        }
        if (out_count) CUDA_TRY(cudaMemsetAsync(out_count, 0, 4, stream));
    }
    // never ask for more hits than the shard has rows (slots past that stay empty)
    // (the fused exchange needs the same k on every rank, whatever the local row count)
    const uint32_t k_need = xchg ? k : (uint32_t)std::min<uint64_t>(k, sh.rows);
    if (out_hits && k_need < k)
        CUDA_TRY(cudaMemsetAsync(out_hits + k_need, 0, (size_t)(k - k_need) * sizeof(nm::ShardHit),
                                 stream));
    for (uint32_t done = 0; done < k_need; done += nm::kMaxFastK) {
        const uint32_t kp = std::min<uint32_t>(nm::kMaxFastK, k_need - done);
        p.k = kp;
        p.out_keys = chained ? ws.d_pass_keys + done : nullptr;
        p.out_hits = out_hits ? out_hits + done : nullptr;
        p.out_rows = out_rows ? out_rows + done : nullptr;
        p.out_scores = out_scores ? out_scores + done : nullptr;
        p.out_count = out_count;
        p.accumulate_count = chained ? 1u : 0u;
        // the previous pass always has kMaxFastK slots; its last one is 0 when it ran dry
        p.key_ceiling = done ? ws.d_pass_keys + done - 1 : nullptr;
        idx->scan_launches++;
        int rc;
        switch (metric) {
        case NM_COSINE: rc = launch_scan_t<nm::kCosine>(sh, p, smem, stream); break;
        case NM_EUCLIDEAN: rc = launch_scan_t<nm::kEuclidean>(sh, p, smem, stream); break;
        default: rc = launch_scan_t<nm::kDot>(sh, p, smem, stream); break;
        }
        if (rc) return rc;
    }
    return NM_OK;
}


// ---- batched-query path ----------------------------------------------------------------
constexpr int kBatchMaxQB = 64;

template <int METRIC, int QB>
int launch_score_batch_t(const Shard &sh, const nm::BatchScoreParams &p, cudaStream_t stream) {
    static std::mutex mu;
    static bool configured[64] = {false};
    auto kern = nm::score_batch_kernel<METRIC, QB>;
    {
        std::lock_guard<std::mutex> g(mu);
        if (sh.device >= 64 || !configured[sh.device]) {
            CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          227 * 1024));
            if (sh.device < 64) configured[sh.device] = true;
        }
    }
    size_t smem = 1024 + (size_t)p.n_stages * nm::batch_stage_bytes<QB>() + 2 * nm::kMaxStages * 8 +
                  nm::kMaxStages * 4 + QB * 4 + 64;
    uint32_t n_rb = (p.n_rows + nm::kRowsPerBlock - 1) / nm::kRowsPerBlock;
    uint32_t grid = std::min<uint32_t>((uint32_t)sh.sm_count, n_rb);
    kern<<<grid, nm::kScanThreads, smem, stream>>>(sh.tmap, p);
    CUDA_TRY(cudaGetLastError());
    return NM_OK;
}

template <int QB>
uint32_t batch_stages() {
    uint32_t st = nm::kMaxStages;
    while (st > 2 && 1024 + (size_t)st * nm::batch_stage_bytes<QB>() + 512 + QB * 4 > 227 * 1024) --st;
    return st;
}

// Stages the single-query kernel can afford once the whole query sits in shared memory.
uint32_t single_query_stages(uint32_t dim) {
    const uint32_t q_floats = (dim + 31u) & ~31u;
    uint32_t stages = nm::kMaxStages;
    while (stages > 0 && scan_smem_bytes(stages, q_floats) > 227 * 1024) --stages;
    return stages;
}

bool batch_eligible(const nm_index *idx, const Shard &sh, uint32_t nq, uint32_t k, int metric) {
    // very long vectors: the single-query kernel keeps the query in shared memory and runs out
    // of stages; the batched kernels stream the query chunk by chunk, so they take over
    const bool long_rows = single_query_stages(idx->dim) < 4;
    if ((nq < kBatchMinQueries && !long_rows) || sh.rows == 0) return false;
    if (std::min<uint64_t>(k, sh.rows) > (uint64_t)nm::kMaxFastK) return false;
    // dot/cosine lanes need whole f32x8 groups; the scalar tail only exists on the 1-query path
    if (metric != NM_EUCLIDEAN && (idx->dim % 8u) != 0) return false;
    return true;
}

// All nq queries over one shard through the batched kernels.  Returns NM_OK, an error, or
// -1 when the score matrix cannot be allocated (caller falls back to single-query passes).
int scan_queries_batched(nm_index *idx, const Shard &sh, Workspace &ws, const float *d_queries,
                         uint32_t nq, uint32_t k, int metric, uint64_t row_base,
                         uint64_t *out_rows, float *out_scores, uint32_t *out_counts,
                         nm::ShardHit *out_hits, cudaStream_t stream, const uint32_t *d_gate) {
    const uint32_t dim = idx->dim;
    const uint32_t n_kc = (dim + 31u) / 32u;
    const uint32_t k_eff = (uint32_t)std::min<uint64_t>(k, sh.rows);
    const uint32_t n_rb = ((uint32_t)sh.rows + nm::kRowsPerBlock - 1) / nm::kRowsPerBlock;
    const uint64_t stride = ((uint64_t)sh.rows + 63u) & ~uint64_t(63);
    // queries per pass: 4 / 8 / 16 (HBM-bound up to ~8, then FP32-issue-bound), 64 for L2 only
    auto round_qb = [](uint32_t v) { return v <= 4u ? 4u : (v <= 8u ? 8u : 16u); };
    const uint32_t qb_max = (metric == NM_EUCLIDEAN && nq > 16) ? 64u : round_qb(nq);
    // scratch
    if (!ws.d_bctl) {
        CUDA_TRY(cudaMalloc(&ws.d_bctl, (2 + kBatchMaxQB) * sizeof(uint32_t)));
        CUDA_TRY(cudaMemsetAsync(ws.d_bctl, 0, (2 + kBatchMaxQB) * sizeof(uint32_t), stream));
        CUDA_TRY(cudaMalloc(&ws.d_qmag, kBatchMaxQB * sizeof(float)));
    }
    size_t need_qt = (size_t)n_kc * 32u * qb_max;
    size_t need_scores = (size_t)qb_max * stride;
    uint32_t ctas_per_q = std::max<uint32_t>(1u, std::min<uint32_t>(n_rb, (2u * sh.sm_count + qb_max - 1) / qb_max));
    size_t need_cand = (size_t)qb_max * ctas_per_q * k_eff;
    if (ws.qt_cap < need_qt || ws.scores_cap < need_scores || ws.bcand_cap < need_cand)
        CUDA_TRY(cudaStreamSynchronize(stream));  // earlier launches may still read the old buffers
    if (ws.qt_cap < need_qt) {
        if (ws.d_qt) CUDA_TRY(cudaFree(ws.d_qt));
        ws.qt_cap = 0;
        CUDA_TRY(cudaMalloc(&ws.d_qt, need_qt * 4));
        ws.qt_cap = need_qt;
    }
    if (ws.scores_cap < need_scores) {
        if (ws.d_scores) CUDA_TRY(cudaFree(ws.d_scores));
        ws.scores_cap = 0;
        if (cudaMalloc(&ws.d_scores, need_scores * 4) != cudaSuccess) {
            cudaGetLastError();
            ws.d_scores = nullptr;
            return -1;
        }
        ws.scores_cap = need_scores;
    }
    if (ws.bcand_cap < need_cand) {
        if (ws.d_bcand) CUDA_TRY(cudaFree(ws.d_bcand));
        ws.bcand_cap = 0;
        CUDA_TRY(cudaMalloc(&ws.d_bcand, need_cand * 8));
        ws.bcand_cap = need_cand;
    }
    if (out_hits && k_eff < k && !d_gate)
        CUDA_TRY(cudaMemsetAsync(out_hits, 0, (size_t)nq * k * sizeof(nm::ShardHit), stream));

    for (uint32_t q0 = 0; q0 < nq; q0 += qb_max) {
        const uint32_t nqp = std::min<uint32_t>(qb_max, nq - q0);
        const uint32_t qb = nqp > 16u ? 64u : round_qb(nqp);
        nm::prepare_batch_kernel<<<std::max<uint32_t>(8u, (n_kc * 32u * qb + 255u) / 256u), 256, 0, stream>>>(
            d_queries + (size_t)q0 * dim, nqp, dim, qb, n_kc, ws.d_qt, ws.d_qmag,
            d_gate ? d_gate + q0 : nullptr);
        CUDA_TRY(cudaGetLastError());
        nm::BatchScoreParams sp;
        memset(&sp, 0, sizeof(sp));
        sp.qt = ws.d_qt;
        sp.qmag = ws.d_qmag;
        sp.scores = ws.d_scores;
        sp.cursor = ws.d_bctl;
        sp.done = ws.d_bctl + 1;
        sp.score_stride = stride;
        sp.n_rows = (uint32_t)sh.rows;
        sp.dim = dim;
        sp.evict_first = (sh.rows * idx->pitch * 4ull > (64ull << 20)) ? 1u : 0u;
        sp.gate = d_gate ? d_gate + q0 : nullptr;
        sp.gate_n = nqp;
        int rc;
#define NM_SCORE_BATCH(METRIC_, QB_)                                          \
    do {                                                                      \
        sp.n_stages = batch_stages<QB_>();                                    \
        rc = launch_score_batch_t<METRIC_, QB_>(sh, sp, stream);              \
    } while (0)
#define NM_SCORE_BATCH_SMALL(METRIC_)                                         \
    do {                                                                      \
        if (qb == 4u) NM_SCORE_BATCH(METRIC_, 4);                             \
        else if (qb == 8u) NM_SCORE_BATCH(METRIC_, 8);                        \
        else NM_SCORE_BATCH(METRIC_, 16);                                     \
    } while (0)
        if (metric == NM_EUCLIDEAN) {
            if (qb == 64u) NM_SCORE_BATCH(nm::kEuclidean, 64);
            else NM_SCORE_BATCH_SMALL(nm::kEuclidean);
        } else if (metric == NM_COSINE) {
            NM_SCORE_BATCH_SMALL(nm::kCosine);
        } else {
            NM_SCORE_BATCH_SMALL(nm::kDot);
        }
#undef NM_SCORE_BATCH_SMALL
#undef NM_SCORE_BATCH
        if (rc) return rc;
        nm::BatchSelectParams bp;
        memset(&bp, 0, sizeof(bp));
        bp.scores = ws.d_scores;
        bp.score_stride = stride;
        bp.cand = ws.d_bcand;
        bp.tickets = ws.d_bctl + 2;
        bp.out_hits = out_hits ? out_hits + (size_t)q0 * k : nullptr;
        bp.out_rows = out_rows ? out_rows + (size_t)q0 * k : nullptr;
        bp.out_scores = out_scores ? out_scores + (size_t)q0 * k : nullptr;
        bp.out_counts = out_counts ? out_counts + q0 : nullptr;
        bp.row_base = row_base;
        bp.out_stride = k;
        bp.n_rows = (uint32_t)sh.rows;
        bp.k = k_eff;
        bp.gate = sp.gate;
        bp.gate_n = nqp;
        dim3 grid(ctas_per_q, nqp);
        nm::select_batch_kernel<<<grid, nm::kRowsPerBlock, 0, stream>>>(bp);
        CUDA_TRY(cudaGetLastError());
        idx->scan_launches += 3;
    }
    return NM_OK;
}


// ---- exact int8 pre-filter path (single shard, host-synchronous nm_search only) -----------
size_t prefilter_smem_bytes(uint32_t n_stages, uint32_t q_words) {
    return 1024 + (size_t)n_stages * nm::kStageBytes + (size_t)nm::kCandCap * 8 +
           (size_t)q_words * 4 + 2 * nm::kMaxStages * 8 + 256 + 64 +
           (size_t)nm::kKeptStage * sizeof(nm::KeptEntry);
}

bool prefilter_usable(const nm_index *idx, const Shard &sh, uint32_t nq, uint32_t k, int metric,
                      bool masked) {
    if (idx->prefilter.load() != 1 || masked) return false;
    if (!sh.tmap8_valid || sh.q8_rows != sh.rows || sh.rows == 0) return false;
    if (k > (uint32_t)nm::kMaxFastK || k > sh.rows) return false;
    if (idx->batching.load() && nq >= kBatchMinQueries &&
        (metric == NM_EUCLIDEAN || (idx->dim % 8u) == 0))
        return false;  // batches share a corpus pass in the batched kernels instead
    uint32_t q_words = ((idx->dim + 127u) / 128u) * 32u;
    return prefilter_smem_bytes(3, q_words) <= 227 * 1024;
}

int ws_ensure_prefilter(Workspace &ws, uint32_t nq) {
    if (!ws.d_kept) {
        CUDA_TRY(cudaMalloc(&ws.d_kept, (size_t)nm::kKeptCap * sizeof(nm::KeptEntry)));
        CUDA_TRY(cudaMalloc(&ws.d_exact_keys, (size_t)nm::kKeptCap * sizeof(uint64_t)));
    }
    if (ws.pf_ctl_cap < nq) {
        if (ws.d_pf_ctl) CUDA_TRY(cudaFree(ws.d_pf_ctl));
        if (ws.h_pf_ctl) CUDA_TRY(cudaFreeHost(ws.h_pf_ctl));
        ws.pf_ctl_cap = 0;
        CUDA_TRY(cudaMalloc(&ws.d_pf_ctl, (size_t)nq * 8 * sizeof(uint32_t)));
        CUDA_TRY(cudaMallocHost(&ws.h_pf_ctl, (size_t)nq * 8 * sizeof(uint32_t)));
        ws.pf_ctl_cap = nq;
    }
    return NM_OK;
}

// Two launches per query: int8 scan (intervals, kept list, k-th best lower bound), then exact
// re-score + selection.  ctl block q keeps the status for the host to inspect afterwards.
int launch_prefiltered(nm_index *idx, const Shard &sh, Workspace &ws, const float *d_query,
                       uint32_t q, uint32_t k, int metric, uint64_t *out_rows, float *out_scores,
                       uint32_t *out_count, cudaStream_t stream) {
    static std::mutex mu;
    static bool configured[64] = {false};
    {
        std::lock_guard<std::mutex> g(mu);
        if (sh.device >= 64 || !configured[sh.device]) {
            CUDA_TRY(cudaFuncSetAttribute(nm::prefilter_scan_kernel,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
            if (sh.device < 64) configured[sh.device] = true;
        }
    }
    uint32_t *ctl = ws.d_pf_ctl + (size_t)q * 8;
    CUDA_TRY(cudaMemsetAsync(ctl, 0, 8 * sizeof(uint32_t), stream));
    nm::PrefilterParams p;
    memset(&p, 0, sizeof(p));
    p.query = d_query;
    p.meta = sh.d_meta;
    p.cand = ws.d_cand;
    p.ctl = ctl;
    p.kept = ws.d_kept;
    p.n_rows = (uint32_t)sh.rows;
    p.dim = idx->dim;
    p.k = k;
    p.q_words = ((idx->dim + 127u) / 128u) * 32u;
    p.metric = metric == NM_COSINE ? nm::kCosine : (metric == NM_EUCLIDEAN ? nm::kEuclidean : nm::kDot);
    uint32_t stages = nm::kMaxStages;
    while (stages > 3 && prefilter_smem_bytes(stages, p.q_words) > 227 * 1024) --stages;
    p.n_stages = stages;
    const uint32_t n_rb = ((uint32_t)sh.rows + nm::kRowsPerBlock - 1) / nm::kRowsPerBlock;
    const uint32_t grid = std::min<uint32_t>((uint32_t)sh.sm_count, n_rb);
    nm::prefilter_scan_kernel<<<grid, nm::kScanThreads, prefilter_smem_bytes(stages, p.q_words),
                                stream>>>(sh.tmap8, p);
    CUDA_TRY(cudaGetLastError());
    nm::RescoreParams r;
    memset(&r, 0, sizeof(r));
    r.query = d_query;
    r.rows = sh.d_rows;
    r.pitch = idx->pitch;
    r.dim = idx->dim;
    r.kept = ws.d_kept;
    r.ctl = ctl;
    r.exact_keys = ws.d_exact_keys;
    r.out_rows = out_rows;
    r.out_scores = out_scores;
    r.out_count = out_count;
    r.row_base = sh.row_base;
    r.k = k;
    r.metric = p.metric;
    nm::prefilter_rescore_kernel<<<(uint32_t)sh.sm_count, nm::kRowsPerBlock, 0, stream>>>(r);
    CUDA_TRY(cudaGetLastError());
    idx->scan_launches += 2;
    return NM_OK;
}

// ---- tensor-core batch pre-filter (tc_prefilter_kernels.cuh) ------------------------------
constexpr uint32_t kTcSortBuckets = 8192;  // row buckets of the re-score ordering
constexpr uint32_t kTcMinRows = 65536;  // below this the exact kernels are cheaper

// shape part of tc_usable: would a batch of nq queries over a shard of this size take the path?
bool tc_shape_ok(const nm_index *idx, uint64_t shard_rows, uint32_t nq, uint32_t k) {
    if (nq < 2 || shard_rows < kTcMinRows || k > (uint32_t)nm::kMaxFastK) return false;
    if ((uint64_t)idx->dim * 16129ull >= 0x7fffffffull) return false;   // s32 accumulators
    if ((size_t)idx->dim * 4 > 160 * 1024) return false;                // rescore keeps q in smem
    return true;
}

bool tc_usable(const nm_index *idx, const Shard &sh, uint32_t nq, uint32_t k, int metric,
               bool masked) {
    (void)metric;
    if (!idx->prefilter.load() || !idx->tensor_core.load() || !idx->batching.load() || masked)
        return false;
    if (!sh.tmap8_valid || sh.q8_rows != sh.rows) return false;
    return tc_shape_ok(idx, sh.rows, nq, k);
}

// auxiliary words of the tensor-core path: [cap kept_n][cap kept_prev][4 stats][passes x TcCtl]
struct TcAux {
    uint32_t *kept_n, *kept_prev, *stats;
    nm::TcCtl *ctl;
    size_t tail_off, tail_bytes;  // stats + ctl: copied back with the results
};
static TcAux tc_aux(const Workspace &ws) {
    TcAux a;
    const size_t cap = ws.tc_nq_cap;
    a.kept_n = ws.d_tc_kept_n;
    a.kept_prev = ws.d_tc_kept_n + cap;
    a.stats = ws.d_tc_kept_n + 2 * cap;
    a.ctl = reinterpret_cast<nm::TcCtl *>(ws.d_tc_kept_n + 2 * cap + 8);
    a.tail_off = (2 * cap) * sizeof(uint32_t);
    a.tail_bytes = 8 * sizeof(uint32_t) + (cap / nm::kTcMaxQ + 1) * sizeof(nm::TcCtl);
    return a;
}

static int ws_ensure_tc(Workspace &ws, uint32_t dim, uint32_t nq, cudaStream_t stream) {
    const size_t q8_bytes = (size_t)nm::kTcMaxQ * q8_pitch(dim);
    if (ws.tc_q8_cap < q8_bytes || ws.tc_nq_cap < nq) CUDA_TRY(cudaStreamSynchronize(stream));
    if (ws.tc_q8_cap < q8_bytes) {
        if (ws.d_tc_q8) CUDA_TRY(cudaFree(ws.d_tc_q8));
        ws.tc_q8_cap = 0;
        CUDA_TRY(cudaMalloc(&ws.d_tc_q8, q8_bytes));
        ws.tc_q8_cap = q8_bytes;
    }
    if (ws.tc_nq_cap < nq) {
        const size_t cap = (std::max<size_t>(nq, nm::kTcMaxQ) + nm::kTcMaxQ - 1) / nm::kTcMaxQ * nm::kTcMaxQ;
        if (ws.d_tc_qmeta) CUDA_TRY(cudaFree(ws.d_tc_qmeta));
        if (ws.h_tc_qmeta) CUDA_TRY(cudaFreeHost(ws.h_tc_qmeta));
        if (ws.d_tc_coef) CUDA_TRY(cudaFree(ws.d_tc_coef));
        if (ws.d_tc_kept_n) CUDA_TRY(cudaFree(ws.d_tc_kept_n));
        if (ws.d_tc_redo) CUDA_TRY(cudaFree(ws.d_tc_redo));
        ws.tc_nq_cap = 0;
        ws.d_tc_qmeta = ws.h_tc_qmeta = ws.d_tc_coef = nullptr;
        ws.d_tc_kept_n = nullptr;
        ws.d_tc_redo = nullptr;
        const size_t tail = 8 * sizeof(uint32_t) + (cap / nm::kTcMaxQ + 1) * sizeof(nm::TcCtl);
        CUDA_TRY(cudaMalloc(&ws.d_tc_qmeta, cap * sizeof(nm::TcQueryMeta)));
        CUDA_TRY(cudaMallocHost(&ws.h_tc_qmeta, cap * sizeof(nm::TcQueryMeta) + tail));
        CUDA_TRY(cudaMalloc(&ws.d_tc_coef, cap * sizeof(float4)));
        CUDA_TRY(cudaMalloc(&ws.d_tc_kept_n, 2 * cap * sizeof(uint32_t) + tail));
        CUDA_TRY(cudaMalloc(&ws.d_tc_redo, cap * sizeof(uint32_t)));
        ws.tc_nq_cap = cap;
    }
    if (!ws.d_tc_kept) {
        CUDA_TRY(cudaMalloc(&ws.d_tc_bucket, (size_t)(kTcSortBuckets + 2) * sizeof(uint32_t)));
        CUDA_TRY(cudaMalloc(&ws.d_tc_sorted, (size_t)nm::kTcMaxQ * nm::kTcKeptCap * sizeof(uint2)));
        CUDA_TRY(cudaMalloc(&ws.d_tc_kept, (size_t)nm::kTcMaxQ * nm::kTcKeptCap * sizeof(nm::TcKept)));
        CUDA_TRY(cudaMalloc(&ws.d_tc_keys, (size_t)nm::kTcMaxQ * nm::kTcKeptCap * sizeof(uint64_t)));
    }
    return NM_OK;
}

// One pass per 256 queries: prepare, then kTcMaxPhases (gemm, refine) pairs whose row ranges
// are chosen ON DEVICE (the first 2048 rows keep everything; each refine sizes the next range
// from the pass rate it saw; pairs past the end exit at once), then the exact re-score.
constexpr uint32_t kTcMaxPhases = 12;

int scan_queries_tc(nm_index *idx, const Shard &sh, Workspace &ws, const float *d_queries,
                    uint32_t nq, uint32_t k, int metric, uint64_t row_base, uint64_t *out_rows,
                    float *out_scores, uint32_t *out_counts, cudaStream_t stream, int *debug_dots,
                    const uint32_t *d_row_mask) {
    static std::mutex mu;
    static bool configured[64] = {false};
    {
        std::lock_guard<std::mutex> g(mu);
        if (sh.device >= 64 || !configured[sh.device]) {
            CUDA_TRY(cudaFuncSetAttribute(nm::tc_gemm_filter_kernel<1>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
            CUDA_TRY(cudaFuncSetAttribute(nm::tc_gemm_filter_kernel<2>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
            CUDA_TRY(cudaFuncSetAttribute(nm::tc_rescore_kernel,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
            CUDA_TRY(cudaFuncSetAttribute(nm::tc_refine_kernel,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
            if (sh.device < 64) configured[sh.device] = true;
        }
    }
    static const bool screen = [] {
        const char *e = getenv("NM_TC_SCREEN");
        return !(e && e[0] == '0');
    }();
    const uint32_t dim = idx->dim, pitch8 = q8_pitch(dim);
    const uint32_t rows = (uint32_t)sh.rows;
    const uint32_t k_eff = std::min<uint32_t>(k, rows);
    int rc = ws_ensure_tc(ws, dim, nq, stream);
    if (rc) return rc;
    auto *qmeta = static_cast<nm::TcQueryMeta *>(ws.d_tc_qmeta);
    auto *coef = static_cast<float4 *>(ws.d_tc_coef);
    auto *kept = static_cast<nm::TcKept *>(ws.d_tc_kept);
    const TcAux aux = tc_aux(ws);
    CUDA_TRY(cudaMemsetAsync(aux.stats, 0, 8 * sizeof(uint32_t), stream));
    const int kmetric = metric == NM_COSINE ? nm::kCosine
                                            : (metric == NM_EUCLIDEAN ? nm::kEuclidean : nm::kDot);
    // cta_group::2 (pairs of CTAs on one 256-row super tile) unless NM_TC_PAIR=0
    static const bool pair = [] {
        const char *e = getenv("NM_TC_PAIR");
        return !(e && e[0] == '0');
    }();
    const uint32_t ctas = pair ? 2u : 1u;
    const uint32_t max_tiles = (rows + nm::kTcM * ctas - 1) / (nm::kTcM * ctas);
    const uint32_t grid = ctas * std::min<uint32_t>((uint32_t)sh.sm_count / ctas, max_tiles);
    uint32_t pass = 0;
    for (uint32_t q0 = 0; q0 < nq; q0 += nm::kTcMaxQ, ++pass) {
        const uint32_t nqp = std::min<uint32_t>(nm::kTcMaxQ, nq - q0);
        const uint32_t n_pad = pair ? ((nqp + 31u) & ~31u) : ((nqp + 15u) & ~15u);  // UMMA N
        nm::TcCtl *ctl = aux.ctl + pass;
        nm::tc_prepare_queries_kernel<<<n_pad, 256, 0, stream>>>(
            d_queries + (size_t)q0 * dim, nqp, dim, pitch8, ws.d_tc_q8, qmeta + q0, coef + q0,
            aux.kept_n + q0, aux.kept_prev + q0, ctl, rows, d_row_mask ? 1u : 0u);
        CUDA_TRY(cudaGetLastError());
        CUtensorMap tmap_q;
        rc = encode_tmap_u8(&tmap_q, ws.d_tc_q8, dim, n_pad, pitch8, 128, n_pad / ctas);
        if (rc) return rc;
        nm::TcGemmParams gp;
        memset(&gp, 0, sizeof(gp));
        gp.meta = sh.d_meta;
        gp.norms = sh.d_norms;
        gp.qmeta = qmeta + q0;
        gp.coef = coef + q0;
        gp.kept = kept;
        gp.kept_n = aux.kept_n + q0;
        gp.dump = debug_dots ? debug_dots + (size_t)q0 * rows : nullptr;
        gp.dump_stride = rows;
        gp.ctl = ctl;
        gp.n_rows = rows;
        gp.dim = dim;
        gp.nq = nqp;
        gp.n_pad = n_pad;
        gp.evict_first = ((uint64_t)rows * pitch8 > (64ull << 20)) ? 1u : 0u;
        gp.screen = screen ? 1u : 0u;
        gp.shift = 0;
        while (((16300ull * dim) >> gp.shift) >= (1ull << 22) - 64) ++gp.shift;
        gp.metric = kmetric;
        gp.row_mask = d_row_mask;
        nm::TcRefineParams rp;
        memset(&rp, 0, sizeof(rp));
        rp.kept = kept;
        rp.kept_n = aux.kept_n + q0;
        rp.kept_prev = aux.kept_prev + q0;
        rp.qmeta = qmeta + q0;
        rp.coef = coef + q0;
        rp.ctl = ctl;
        rp.queries = d_queries + (size_t)q0 * dim;
        rp.rows = sh.d_rows;
        rp.pitch = idx->pitch;
        rp.n_rows = rows;
        rp.dim = dim;
        rp.k = k_eff;
        rp.metric = kmetric;
        static const uint32_t grow0 = [] {
            const char *e = getenv("NM_TC_GROW0");
            return e ? (uint32_t)std::max(1, atoi(e)) : 4u;
        }();
        static const uint32_t allow_mul = [] {
            const char *e = getenv("NM_TC_ALLOW");
            return e ? (uint32_t)std::max(1, atoi(e)) : 1u;
        }();
        rp.grow0 = grow0;
        rp.allow_mul = allow_mul;
        for (uint32_t ph = 0; ph < kTcMaxPhases; ++ph) {
            if (pair) {
                cudaLaunchConfig_t cfg = {};
                cfg.gridDim = dim3(grid);
                cfg.blockDim = dim3(nm::kTcThreads);
                cfg.dynamicSmemBytes = nm::tc_gemm_smem_bytes();
                cfg.stream = stream;
                cudaLaunchAttribute attr[1];
                attr[0].id = cudaLaunchAttributeClusterDimension;
                attr[0].val.clusterDim.x = 2;
                attr[0].val.clusterDim.y = 1;
                attr[0].val.clusterDim.z = 1;
                cfg.attrs = attr;
                cfg.numAttrs = 1;
                CUDA_TRY(cudaLaunchKernelEx(&cfg, nm::tc_gemm_filter_kernel<2>, sh.tmap8_tc, tmap_q, gp));
            } else {
                nm::tc_gemm_filter_kernel<1><<<grid, nm::kTcThreads, nm::tc_gemm_smem_bytes(), stream>>>(
                    sh.tmap8_tc, tmap_q, gp);
            }
            CUDA_TRY(cudaGetLastError());
            nm::tc_refine_kernel<<<nqp, 256, (size_t)dim * 4, stream>>>(rp);
            CUDA_TRY(cudaGetLastError());
        }
        nm::TcRescoreParams sp;
        memset(&sp, 0, sizeof(sp));
        sp.queries = d_queries + (size_t)q0 * dim;
        sp.rows = sh.d_rows;
        sp.kept = kept;
        sp.kept_n = aux.kept_n + q0;
        sp.qmeta = qmeta + q0;
        sp.exact_keys = ws.d_tc_keys;
        sp.out_rows = out_rows + (size_t)q0 * k;
        sp.out_scores = out_scores + (size_t)q0 * k;
        sp.out_counts = out_counts + q0;
        sp.stats = aux.stats;
        sp.row_base = row_base;
        sp.pitch = idx->pitch;
        sp.dim = dim;
        sp.k = k_eff;
        sp.out_stride = k;
        sp.metric = kmetric;
        if (dim % 4u == 0u && (size_t)nqp <= 0xffffu) {
            // survivors of all queries, re-scored in corpus order (see tc_score_sorted_kernel)
            nm::TcSortParams so;
            memset(&so, 0, sizeof(so));
            so.kept = kept;
            so.kept_n = aux.kept_n + q0;
            so.bucket = ws.d_tc_bucket;
            so.sorted = static_cast<uint2 *>(ws.d_tc_sorted);
            so.total = ws.d_tc_bucket + kTcSortBuckets + 1;
            so.shift = 8;
            while (((uint64_t)rows >> so.shift) + 1 > kTcSortBuckets) ++so.shift;
            so.n_buckets = (uint32_t)(((uint64_t)rows >> so.shift) + 1);
            CUDA_TRY(cudaMemsetAsync(so.bucket, 0, (size_t)(so.n_buckets + 1) * sizeof(uint32_t), stream));
            nm::tc_sort_count_kernel<<<nqp, 256, 0, stream>>>(so);
            CUDA_TRY(cudaGetLastError());
            nm::tc_sort_scan_kernel<<<1, 1024, 0, stream>>>(so);
            CUDA_TRY(cudaGetLastError());
            nm::tc_sort_scatter_kernel<<<nqp, 256, 0, stream>>>(so);
            CUDA_TRY(cudaGetLastError());
            nm::tc_score_sorted_kernel<<<(uint32_t)sh.sm_count * 8u, 256, 0, stream>>>(sp, so.sorted,
                                                                                      so.total);
            CUDA_TRY(cudaGetLastError());
            nm::tc_select_kernel<<<nqp, nm::kRowsPerBlock, 0, stream>>>(sp);
            CUDA_TRY(cudaGetLastError());
            idx->scan_launches += 4;
        } else {
            nm::tc_rescore_kernel<<<nqp, nm::kRowsPerBlock, (size_t)dim * 4, stream>>>(sp);
            CUDA_TRY(cudaGetLastError());
        }
        idx->scan_launches += 2 + 2 * kTcMaxPhases;
    }
    // flags, phase control and statistics come back with the results
    CUDA_TRY(cudaMemcpyAsync(ws.h_tc_qmeta, ws.d_tc_qmeta, (size_t)nq * sizeof(nm::TcQueryMeta),
                             cudaMemcpyDeviceToHost, stream));
    CUDA_TRY(cudaMemcpyAsync(static_cast<uint8_t *>(ws.h_tc_qmeta) +
                                 ws.tc_nq_cap * sizeof(nm::TcQueryMeta),
                             reinterpret_cast<const uint8_t *>(ws.d_tc_kept_n) + aux.tail_off,
                             aux.tail_bytes, cudaMemcpyDeviceToHost, stream));
    return NM_OK;
}

// after the stream has been waited for: which queries must be redone exactly
// (a pass whose phases did not reach the end of the corpus flags all of its queries)
uint32_t tc_query_flags(const Workspace &ws, uint32_t q, uint32_t rows) {
    const uint8_t *tail = static_cast<const uint8_t *>(ws.h_tc_qmeta) +
                          ws.tc_nq_cap * sizeof(nm::TcQueryMeta);
    const nm::TcCtl *ctl = reinterpret_cast<const nm::TcCtl *>(tail + 8 * sizeof(uint32_t));
    uint32_t f = static_cast<const nm::TcQueryMeta *>(ws.h_tc_qmeta)[q].flags;
    if (ctl[q / nm::kTcMaxQ].row_begin < rows) f |= 4u;
    return f;
}
uint32_t tc_survivors(const Workspace &ws) {
    return *reinterpret_cast<const uint32_t *>(static_cast<const uint8_t *>(ws.h_tc_qmeta) +
                                               ws.tc_nq_cap * sizeof(nm::TcQueryMeta));
}
uint32_t tc_phases(const Workspace &ws) {
    const uint8_t *tail = static_cast<const uint8_t *>(ws.h_tc_qmeta) +
                          ws.tc_nq_cap * sizeof(nm::TcQueryMeta);
    return reinterpret_cast<const nm::TcCtl *>(tail + 8 * sizeof(uint32_t))[0].phases;
}

// ---- the same for sharded indexes: this shard's hits as ShardHit[nq, k] (global rows) -------
// Enqueue: tensor-core pass into scratch, pack to hits, flags to the host.  Finish (after the
// caller has other shards in flight): wait, redo flagged queries with the exact scan.
int scan_queries_tc_hits_enqueue(nm_index *idx, const Shard &sh, Workspace &ws,
                                 const float *d_queries, uint32_t nq, uint32_t k, int metric,
                                 uint64_t row_base, nm::ShardHit *out_hits, cudaStream_t stream,
                                 const uint32_t *d_row_mask) {
    const ResultLayout l = result_layout(nq, k);
    if (ws.tc_out_cap < l.total) {
        CUDA_TRY(cudaStreamSynchronize(stream));
        if (ws.d_tc_out) CUDA_TRY(cudaFree(ws.d_tc_out));
        ws.tc_out_cap = 0;
        CUDA_TRY(cudaMalloc(&ws.d_tc_out, l.total));
        ws.tc_out_cap = l.total;
    }
    uint64_t *t_rows = reinterpret_cast<uint64_t *>(ws.d_tc_out + l.rows_off);
    float *t_scores = reinterpret_cast<float *>(ws.d_tc_out + l.scores_off);
    uint32_t *t_counts = reinterpret_cast<uint32_t *>(ws.d_tc_out + l.counts_off);
    int rc = scan_queries_tc(idx, sh, ws, d_queries, nq, k, metric, row_base, t_rows, t_scores,
                             t_counts, stream, nullptr, d_row_mask);
    if (rc) return rc;
    nm::tc_pack_hits_kernel<<<nq, 256, 0, stream>>>(t_rows, t_scores, t_counts, k, out_hits);
    CUDA_TRY(cudaGetLastError());
    return NM_OK;
}

int scan_queries_tc_hits_finish(nm_index *idx, const Shard &sh, Workspace &ws,
                                const float *d_queries, uint32_t nq, uint32_t k, int metric,
                                uint64_t row_base, nm::ShardHit *out_hits, cudaStream_t stream,
                                const uint32_t *d_row_mask) {
    CUDA_TRY(cudaStreamSynchronize(stream));
    idx->tc_survivors += tc_survivors(ws);
    for (uint32_t q = 0; q < nq; ++q) {
        if (tc_query_flags(ws, q, (uint32_t)sh.rows) == 0) continue;
        idx->tc_fallbacks++;
        int rc = launch_scan(idx, sh, ws, d_queries + (size_t)q * idx->dim, k, metric, row_base,
                             nullptr, nullptr, nullptr, out_hits + (size_t)q * k, stream, nullptr,
                             d_row_mask);
        if (rc) return rc;
    }
    return NM_OK;
}

// nq queries over one shard: batched kernels when that pays, else one (chained) scan per query.
int scan_queries(nm_index *idx, const Shard &sh, Workspace &ws, const float *d_queries,
                 uint32_t nq, uint32_t k, int metric, uint64_t row_base, uint64_t *out_rows,
                 float *out_scores, uint32_t *out_counts, nm::ShardHit *out_hits,
                 cudaStream_t stream, const uint32_t *d_gate) {
    if ((idx->batching.load() || single_query_stages(idx->dim) < 2) &&
        batch_eligible(idx, sh, nq, k, metric)) {
        int rc = scan_queries_batched(idx, sh, ws, d_queries, nq, k, metric, row_base, out_rows,
                                      out_scores, out_counts, out_hits, stream, d_gate);
        if (rc != -1) return rc;
    }
    for (uint32_t q = 0; q < nq; ++q) {
        int rc = launch_scan(idx, sh, ws, d_queries + (size_t)q * idx->dim, k, metric, row_base,
                             out_rows ? out_rows + (size_t)q * k : nullptr,
                             out_scores ? out_scores + (size_t)q * k : nullptr,
                             out_counts ? out_counts + q : nullptr,
                             out_hits ? out_hits + (size_t)q * k : nullptr, stream, nullptr, nullptr,
                             d_gate ? d_gate + q : nullptr);
        if (rc) return rc;
    }
    return NM_OK;
}

int tc_redo_flags(const Workspace &ws, uint32_t nq, uint32_t rows, uint32_t **d_redo, cudaStream_t stream) {
    const TcAux aux = tc_aux(ws);
    nm::tc_redo_flags_kernel<<<(nq + 255u) / 256u, 256, 0, stream>>>(
        static_cast<const nm::TcQueryMeta *>(ws.d_tc_qmeta), aux.ctl, nq, rows, ws.d_tc_redo);
    CUDA_TRY(cudaGetLastError());
    *d_redo = ws.d_tc_redo;
    return NM_OK;
}



int launch_merge_shards(nm_index *idx, const nm::ShardHit *d_gather, uint32_t nq, uint32_t k,
                        uint64_t *out_rows, float *out_scores, uint32_t *out_counts,
                        cudaStream_t stream) {
    const uint64_t total = (uint64_t)idx->n_ranks * k;
    if (total > 4096) {
        // large gathers (k up to the corpus size): rank-based merge of the sorted shard lists
        const dim3 grid((uint32_t)((total + nm::kMergeThreads - 1) / nm::kMergeThreads), nq);
        nm::merge_shards_rank_kernel<<<grid, nm::kMergeThreads, 0, stream>>>(
            d_gather, (uint32_t)idx->n_ranks, k, nq * k, out_rows, out_scores, out_counts);
        CUDA_TRY(cudaGetLastError());
        idx->merge_launches++;
        return NM_OK;
    }
    const uint32_t n_sort = pow2_ceil((uint32_t)total);
    const size_t msmem = (size_t)n_sort * 8;
    nm::merge_shards_kernel<<<nq, nm::kMergeThreads, msmem, stream>>>(
        d_gather, (uint32_t)idx->n_ranks, k, nq * k, n_sort, out_rows, out_scores, out_counts);
    CUDA_TRY(cudaGetLastError());
    idx->merge_launches++;
    return NM_OK;
}

int launch_exchange_empty(const nm::PeerXchg &x, uint32_t k, uint64_t *scratch, uint64_t *out_rows,
                          float *out_scores, uint32_t *out_count, cudaStream_t stream) {
    nm::exchange_empty_shard_kernel<<<1, nm::kRowsPerBlock, 0, stream>>>(x, k, scratch, out_rows,
                                                                         out_scores, out_count);
    CUDA_TRY(cudaGetLastError());
    return NM_OK;
}

int launch_filter_mask(const Shard &sh, const nm::FilterOpDev *d_ops, uint32_t n_ops, uint32_t max_depth,
                       uint64_t n_rows, uint32_t *d_mask, uint64_t n_words, cudaStream_t stream) {
    // one warp per 256-row block of the mask, 8 warps per CTA; giant shards loop
    if (n_words == 0) return NM_OK;
    const uint32_t blocks = (uint32_t)std::min<uint64_t>((n_words / 8 + 7) / 8, (uint64_t)sh.sm_count * 64);
    if (max_depth <= 32)
        nm::filter_mask_kernel<uint32_t><<<blocks, 256, 0, stream>>>(d_ops, n_ops, n_rows, d_mask, n_words);
    else
        nm::filter_mask_kernel<uint64_t><<<blocks, 256, 0, stream>>>(d_ops, n_ops, n_rows, d_mask, n_words);
    CUDA_TRY(cudaGetLastError());
    return NM_OK;
}

int launch_column_move(uint8_t *tags, uint64_t *vals, uint64_t dst, uint64_t src, cudaStream_t stream) {
    nm::column_move_kernel<<<1, 1, 0, stream>>>(tags, vals, dst, src);
    CUDA_TRY(cudaGetLastError());
    return NM_OK;
}

int launch_fill_synthetic(const Shard &sh, float *rows, uint64_t n, uint32_t dim, uint32_t pitch,
                          uint64_t seed, uint64_t global_row0, cudaStream_t stream) {
    nm::fill_synthetic_kernel<<<sh.sm_count * 8, 256, 0, stream>>>(rows, n, dim, pitch, seed,
                                                                   global_row0);
    CUDA_TRY(cudaGetLastError());
    return NM_OK;
}

int launch_quantize(const Shard &sh, const float *rows, uint32_t pitch, uint32_t dim, uint64_t first,
                    uint64_t n, int8_t *q8, uint32_t pitch8, nm::RowMeta *meta, float2 *norms,
                    uint32_t *flag, cudaStream_t stream) {
    uint32_t blocks = (uint32_t)std::min<uint64_t>((n + 7) / 8, (uint64_t)sh.sm_count * 16);
    nm::quantize_rows_kernel<<<blocks, 256, 0, stream>>>(rows, pitch, dim, first, n, q8, pitch8, meta,
                                                         norms, flag);
    CUDA_TRY(cudaGetLastError());
    return NM_OK;
}

}  // namespace nmi
