// nm_search.cu — nm_search / nm_search_masked / nm_search_device: validation, routing between
// the single-query scan, the batched kernels, the int8 pre-filter, the fused peer exchange and
// the NCCL all-gather path, and the host-side merge of in-process multi-device indexes.
#include "nm_internal.hpp"
#include "nm_trace.hpp"

#include <cstdlib>

using namespace nmi;

namespace {

// Wait for the results of one search.  cudaStreamSynchronize parks the calling thread and its
// wake-up costs tens of microseconds once the wait is longer than the driver's spin window —
// about 60 us per query on a 4 ms scan.  Default: poll cudaStreamQuery (one busy core for the
// duration of the scan, the usual latency/CPU trade of a serving thread).  NM_WAIT=block
// restores the blocking wait.
int wait_stream(cudaStream_t stream) {
    static const bool block = [] {
        const char *m = getenv("NM_WAIT");
        return m && strcmp(m, "block") == 0;
    }();
    if (block) {
        CUDA_TRY(cudaStreamSynchronize(stream));
        return NM_OK;
    }
    for (;;) {
        cudaError_t e = cudaStreamQuery(stream);
        if (e == cudaSuccess) return NM_OK;
        if (e != cudaErrorNotReady)
            return fail(NM_ERR_STORAGE, "CUDA error %s while waiting for a search: %s",
                        cudaGetErrorName(e), cudaGetErrorString(e));
#if defined(__x86_64__)
        __builtin_ia32_pause();
#endif
    }
}

// Small result blocks are written by the kernels straight into the pinned host block (mapped
// into the device's address space under UVA): no device-to-host copy operation between the last
// kernel and the caller — its engine hand-over costs more than 136 bytes of posted PCIe writes.
// Stream completion makes the writes visible to the host.  NM_ZERO_COPY_RESULTS=0 switches it off.
constexpr size_t kZeroCopyResultBytes = 4096;
bool zero_copy_results(size_t total) {
    static const bool on = [] {
        const char *m = getenv("NM_ZERO_COPY_RESULTS");
        return !(m && m[0] == '0');
    }();
    return on && total <= kZeroCopyResultBytes;
}

// Device-time bracket of a synchronous search (nm_stats.last_scan_ms).  Two timing events cost
// ~8 us of a 36 us call on a small corpus (each makes the front end drain and write a timestamp),
// so they are recorded only while nm_index_set_profiling is on; otherwise last_scan_ms reads 0.
static inline int time_begin(nm_index *idx, Workspace &ws) {
    ws.timed = idx->profiling.load() != 0;
    if (ws.timed) CUDA_TRY(cudaEventRecord(ws.ev0, ws.stream));
    return NM_OK;
}
static inline int time_end(Workspace &ws) {
    if (ws.timed) CUDA_TRY(cudaEventRecord(ws.ev1, ws.stream));
    return NM_OK;
}
static inline int time_read(Workspace &ws, float *ms) {
    *ms = 0.f;
    if (ws.timed) CUDA_TRY(cudaEventElapsedTime(ms, ws.ev0, ws.ev1));
    return NM_OK;
}

struct HostHit {
    uint32_t ord;
    uint32_t score_bits;
    uint64_t row;
    uint64_t pos;
};

int validate_search(const nm_index *idx, const void *queries, uint32_t nq, uint32_t k, int metric,
                    const void *out_rows, const void *out_scores, const void *out_counts) {
    if (!idx) return fail(NM_ERR_INVALID_ARGUMENT, "null index");
    if (nq == 0 || idx->dim == 0) return fail(NM_ERR_EMPTY_VECTOR, "empty query");
    if (k == 0) return fail(NM_ERR_INVALID_TOP_K, "top_k must be >= 1");
    if (!queries || !out_rows || !out_scores || !out_counts)
        return fail(NM_ERR_INVALID_ARGUMENT, "null buffer");
    if (metric < 0 || metric > 2) return fail(NM_ERR_INVALID_ARGUMENT, "unknown metric %d", metric);
    return NM_OK;
}



// Packed result block device -> pinned host -> caller buffers; one D2H copy, one sync.
int download_results(nm_index *idx, const Shard &sh, Workspace &ws, const ResultLayout &l,
                     uint32_t nq, uint32_t k, uint64_t *out_rows, float *out_scores,
                     uint32_t *out_counts, bool on_host = false) {
    NM_TRACE("wait_download");
    if (!on_host)  // else: the kernels wrote into ws.h_result themselves
        CUDA_TRY(cudaMemcpyAsync(ws.h_result, ws.d_result, l.total, cudaMemcpyDeviceToHost, ws.stream));
    {
        int wrc = wait_stream(ws.stream);
        if (wrc) return wrc;
    }
    float ms = 0.f;
    if (int trc = time_read(ws, &ms)) return trc;
    idx->last_scan_ms = ms;
    const uint32_t *hc = reinterpret_cast<const uint32_t *>(ws.h_result + l.counts_off);
    const uint64_t *hr = reinterpret_cast<const uint64_t *>(ws.h_result + l.rows_off);
    const float *hs = reinterpret_cast<const float *>(ws.h_result + l.scores_off);
    for (uint32_t q = 0; q < nq; ++q) {
        if (hc[q] == 0xffffffffu)  // poisoned by the exchange watchdog
            return fail(NM_ERR_STORAGE, "peer exchange timed out waiting for another rank");
        out_counts[q] = hc[q];
        memcpy(out_rows + (size_t)q * k, hr + (size_t)q * k, (size_t)hc[q] * 8);
        memcpy(out_scores + (size_t)q * k, hs + (size_t)q * k, (size_t)hc[q] * 4);
    }
    idx->searches += nq;
    idx->rows_scanned += (uint64_t)nq * sh.rows;
    idx->bytes_streamed += (uint64_t)nq * sh.rows * idx->dim * 4;
    idx->h2d_bytes += (uint64_t)nq * idx->dim * 4;
    idx->d2h_bytes += l.total;  // zero-copy or copied: the same bytes cross the bus
    return NM_OK;
}

// Rank-independent routing decision for collective searches (every rank must agree).
bool collective_uses_fused_exchange(const nm_index *idx, uint32_t nq, uint32_t k, int metric,
                                    bool masked = false) {
    if (!idx->xchg_ok || k > (uint32_t)nm::kMaxFastK) return false;
    // masked single queries: fused scan + exchange; masked batches: per-shard hit lists (tensor-core
    // pre-filter where a shard can, else query by query) + ONE all-gather
    if (masked) return nq < kBatchMinQueries;
    const bool long_rows = single_query_stages(idx->dim) < 4;
    const bool would_batch = (idx->batching.load() || single_query_stages(idx->dim) < 2) &&
                             (nq >= kBatchMinQueries || long_rows) &&
                             (metric == NM_EUCLIDEAN || (idx->dim % 8u) == 0);
    return !would_batch;
}

// One fused launch per query: scan + peer-memory exchange + merge, results written in place.
int collective_fused(nm_index *idx, const Shard &sh, Workspace &ws, const float *d_queries,
                     uint32_t nq, uint32_t k, int metric, uint64_t *out_rows, float *out_scores,
                     uint32_t *out_counts, cudaStream_t stream, const uint32_t *d_mask = nullptr) {
    for (uint32_t q = 0; q < nq; ++q) {
        nm::PeerXchg x = make_xchg(idx, ++idx->xchg_seq);
        if (sh.rows == 0) {
            int rc = launch_exchange_empty(x, k, ws.d_cand, out_rows + (size_t)q * k,
                                           out_scores + (size_t)q * k, out_counts + q, stream);
            if (rc) return rc;
            idx->merge_launches++;
        } else {
            int rc = launch_scan(idx, sh, ws, d_queries + (size_t)q * idx->dim, k, metric,
                                 idx->comm_row_base, out_rows + (size_t)q * k,
                                 out_scores + (size_t)q * k, out_counts + q, nullptr, stream, &x, d_mask);
            if (rc) return rc;
        }
    }
    return NM_OK;
}

}  // namespace

extern "C" {

// one masked scan per query into this shard's hit list (masked searches do not batch)
static int masked_hits(nm_index *idx, const Shard &sh, Workspace &ws, const float *d_queries, uint32_t nq,
                       uint32_t k, int metric, uint64_t row_base, nm::ShardHit *out_hits,
                       cudaStream_t stream, const uint32_t *d_mask) {
    for (uint32_t q = 0; q < nq; ++q) {
        int rc = launch_scan(idx, sh, ws, d_queries + (size_t)q * idx->dim, k, metric, row_base, nullptr,
                             nullptr, nullptr, out_hits + (size_t)q * k, stream, nullptr, d_mask);
        if (rc) return rc;
    }
    return NM_OK;
}

static int search_impl(nm_index *idx, const float *queries, uint32_t nq, uint32_t k, int metric,
                       const MaskSpec &mspec, uint64_t *out_rows, float *out_scores,
                       uint32_t *out_counts) {
    NM_TRACE(mspec.prog ? "nm_search_filtered" : (mspec.host_mask ? "nm_search_masked" : "nm_search"));
    int rc = validate_search(idx, queries, nq, k, metric, out_rows, out_scores, out_counts);
    if (rc) return rc;
    const bool masked = mspec.any();
    if (nq >= 2) {
        rc = q8_auto_prepare(idx, nq, k);  // auto mode: first eligible batch builds the int8 copy
        if (rc) return rc;
    }
    std::shared_lock<std::shared_mutex> g(idx->mu);
    const size_t G = idx->shards.size();
    const uint32_t dim = idx->dim;
    const bool collective = idx->comm != nullptr;
    if (collective && G != 1)
        return fail(NM_ERR_CONFIGURATION, "a communicator needs a single-device index per rank");
    std::vector<std::shared_ptr<MaskEntry>> mask_holds(G);  // cached masks stay alive until the sync

    // ---- fast path: one device, no communicator: the kernel writes the final result ----
    if (G == 1 && !collective) {
        Shard &sh = *idx->shards[0];
        if (sh.rows == 0) {
            for (uint32_t q = 0; q < nq; ++q) out_counts[q] = 0;
            return NM_OK;
        }
        CUDA_TRY(cudaSetDevice(sh.device));
        std::unique_ptr<Workspace> ws;
        rc = ws_acquire(sh, ws);
        if (rc) return rc;
        struct Releaser {
            Shard &s;
            std::unique_ptr<Workspace> &w;
            ~Releaser() { ws_release(s, w); }
        } rel{sh, ws};
        rc = ws_ensure(*ws, sh, dim, nq, k, true, true, false, 0);
        if (rc) return rc;
        ResultLayout l = result_layout(nq, k);
        {
            NM_TRACE("stage_query");
            memcpy(ws->h_query, queries, (size_t)nq * dim * 4);
            CUDA_TRY(cudaMemcpyAsync(ws->d_query, ws->h_query, (size_t)nq * dim * 4,
                                     cudaMemcpyHostToDevice, ws->stream));
        }
        const bool tc_path = tc_usable(idx, sh, nq, k, metric, false);
        // the pre-filter paths copy their status words back with the results and may redo queries
        // in place: they keep the device-side block
        const bool host_block = zero_copy_results(l.total) && !tc_path &&
                                !prefilter_usable(idx, sh, nq, k, metric, masked);
        uint8_t *r_base = host_block ? ws->h_result : ws->d_result;
        uint64_t *r_rows = reinterpret_cast<uint64_t *>(r_base + l.rows_off);
        float *r_scores = reinterpret_cast<float *>(r_base + l.scores_off);
        uint32_t *r_counts = reinterpret_cast<uint32_t *>(r_base + l.counts_off);
        const uint32_t *d_mask = nullptr;
        if (masked) {
            NM_TRACE("filter_mask");
            rc = shard_mask(idx, sh, *ws, mspec, sh.row_base, ws->stream, &d_mask, &mask_holds[0]);
            if (rc) return rc;
        }
        NM_TRACE("scan");
        if (int trc = time_begin(idx, *ws)) return trc;
        if (masked && !tc_path) {
            for (uint32_t q = 0; q < nq; ++q) {
                rc = launch_scan(idx, sh, *ws, ws->d_query + (size_t)q * dim, k, metric, sh.row_base,
                                 r_rows + (size_t)q * k, r_scores + (size_t)q * k, r_counts + q,
                                 nullptr, ws->stream, nullptr, d_mask);
                if (rc) return rc;
            }
        } else if (tc_path) {
            // tensor-core pre-filter: one int8 GEMM pass per 256 queries + exact re-score; the
            // per-query flags come back with the results, flagged queries are redone exactly.
            // Masked batches too: ineligible rows never enter the kept lists.
            rc = scan_queries_tc(idx, sh, *ws, ws->d_query, nq, k, metric, sh.row_base, r_rows,
                                 r_scores, r_counts, ws->stream, nullptr, d_mask);
            if (rc) return rc;
            if (int trc = time_end(*ws)) return trc;
            CUDA_TRY(cudaMemcpyAsync(ws->h_result, ws->d_result, l.total, cudaMemcpyDeviceToHost,
                                     ws->stream));
            rc = wait_stream(ws->stream);
            if (rc) return rc;
            idx->tc_queries += nq;
            idx->tc_survivors += tc_survivors(*ws);
            uint32_t n_redo = 0;
            for (uint32_t q = 0; q < nq; ++q)
                if (tc_query_flags(*ws, q, (uint32_t)sh.rows) != 0) ++n_redo;
            idx->tc_fallbacks += n_redo;
            const bool redo = n_redo != 0;
            if (n_redo * 2 > nq && !masked) {
                // most of the batch (e.g. a corpus ordered against the running threshold): the
                // exact batched kernels redo all of it in shared corpus passes
                rc = scan_queries(idx, sh, *ws, ws->d_query, nq, k, metric, sh.row_base, r_rows,
                                  r_scores, r_counts, nullptr, ws->stream);
                if (rc) return rc;
            } else if (redo) {
                for (uint32_t q = 0; q < nq; ++q) {
                    if (tc_query_flags(*ws, q, (uint32_t)sh.rows) == 0) continue;
                    rc = launch_scan(idx, sh, *ws, ws->d_query + (size_t)q * dim, k, metric,
                                     sh.row_base, r_rows + (size_t)q * k, r_scores + (size_t)q * k,
                                     r_counts + q, nullptr, ws->stream, nullptr, d_mask);
                    if (rc) return rc;
                }
            }
            if (!redo) {
                float ms = 0.f;
                if (int trc = time_read(*ws, &ms)) return trc;
                idx->last_scan_ms = ms;
                const uint32_t *hc = reinterpret_cast<const uint32_t *>(ws->h_result + l.counts_off);
                const uint64_t *hr = reinterpret_cast<const uint64_t *>(ws->h_result + l.rows_off);
                const float *hs = reinterpret_cast<const float *>(ws->h_result + l.scores_off);
                for (uint32_t q = 0; q < nq; ++q) {
                    out_counts[q] = hc[q];
                    memcpy(out_rows + (size_t)q * k, hr + (size_t)q * k, (size_t)hc[q] * 8);
                    memcpy(out_scores + (size_t)q * k, hs + (size_t)q * k, (size_t)hc[q] * 4);
                }
                idx->searches += nq;
                idx->rows_scanned += (uint64_t)nq * sh.rows;
                idx->bytes_streamed += (uint64_t)((nq + 255u) / 256u) * sh.rows *
                                       (q8_pitch(dim) + sizeof(nm::RowMeta));
                idx->h2d_bytes += (uint64_t)nq * dim * 4;
                idx->d2h_bytes += l.total + (uint64_t)nq * 48;
                return NM_OK;
            }
        } else if (prefilter_usable(idx, sh, nq, k, metric, masked)) {
            rc = ws_ensure_prefilter(*ws, nq);
            if (rc) return rc;
            for (uint32_t q = 0; q < nq; ++q) {
                rc = launch_prefiltered(idx, sh, *ws, ws->d_query + (size_t)q * dim, q, k, metric,
                                        r_rows + (size_t)q * k, r_scores + (size_t)q * k,
                                        r_counts + q, ws->stream);
                if (rc) return rc;
            }
            // status words come back with the results (one sync); queries whose candidate list
            // overflowed (or whose query is not finite) are redone with the exact f32 scan
            CUDA_TRY(cudaMemcpyAsync(ws->h_pf_ctl, ws->d_pf_ctl, (size_t)nq * 8 * sizeof(uint32_t),
                                     cudaMemcpyDeviceToHost, ws->stream));
            if (int trc = time_end(*ws)) return trc;
            CUDA_TRY(cudaMemcpyAsync(ws->h_result, ws->d_result, l.total, cudaMemcpyDeviceToHost,
                                     ws->stream));
            rc = wait_stream(ws->stream);
            if (rc) return rc;
            idx->pf_queries += nq;
            bool redo = false;
            for (uint32_t q = 0; q < nq; ++q) {
                idx->pf_kept += ws->h_pf_ctl[(size_t)q * 8 + 7];
                if (ws->h_pf_ctl[(size_t)q * 8 + 3] == 0) continue;
                idx->pf_fallbacks++;
                redo = true;
                rc = launch_scan(idx, sh, *ws, ws->d_query + (size_t)q * dim, k, metric, sh.row_base,
                                 r_rows + (size_t)q * k, r_scores + (size_t)q * k, r_counts + q,
                                 nullptr, ws->stream);
                if (rc) return rc;
            }
            if (!redo) {
                // results are already on the host: finish without a second copy
                float ms = 0.f;
                if (int trc = time_read(*ws, &ms)) return trc;
                idx->last_scan_ms = ms;
                const uint32_t *hc = reinterpret_cast<const uint32_t *>(ws->h_result + l.counts_off);
                const uint64_t *hr = reinterpret_cast<const uint64_t *>(ws->h_result + l.rows_off);
                const float *hs = reinterpret_cast<const float *>(ws->h_result + l.scores_off);
                for (uint32_t q = 0; q < nq; ++q) {
                    out_counts[q] = hc[q];
                    memcpy(out_rows + (size_t)q * k, hr + (size_t)q * k, (size_t)hc[q] * 8);
                    memcpy(out_scores + (size_t)q * k, hs + (size_t)q * k, (size_t)hc[q] * 4);
                }
                idx->searches += nq;
                idx->rows_scanned += (uint64_t)nq * sh.rows;
                idx->bytes_streamed += (uint64_t)nq * sh.rows * (q8_pitch(dim) + sizeof(nm::RowMeta));
                idx->h2d_bytes += (uint64_t)nq * dim * 4;
                idx->d2h_bytes += l.total + (uint64_t)nq * 32;
                return NM_OK;
            }
        } else {
            rc = scan_queries(idx, sh, *ws, ws->d_query, nq, k, metric, sh.row_base, r_rows, r_scores,
                              r_counts, nullptr, ws->stream);
            if (rc) return rc;
        }
        if (int trc = time_end(*ws)) return trc;
        return download_results(idx, sh, *ws, l, nq, k, out_rows, out_scores, out_counts, host_block);
    }

    // ---- collective path: one shard per process.  Single queries: ONE fused launch (scan +
    //      peer-memory exchange + merge).  Batches / k > 1024: scan, ONE ncclAllGather of the
    //      per-shard hits, merge kernel. ----
    if (collective) {
        Shard &sh = *idx->shards[0];
        CUDA_TRY(cudaSetDevice(sh.device));
        std::lock_guard<std::mutex> cg(idx->comm_mu);
        if (idx->xchg_async_stream) {  // asynchronous collective searches come first
            CUDA_TRY(cudaStreamSynchronize(idx->xchg_async_stream));
            idx->xchg_async_stream = nullptr;
        }
        std::unique_ptr<Workspace> ws;
        rc = ws_acquire(sh, ws);
        if (rc) return rc;
        struct Releaser {
            Shard &s;
            std::unique_ptr<Workspace> &w;
            ~Releaser() { ws_release(s, w); }
        } rel{sh, ws};
        rc = ws_ensure(*ws, sh, dim, nq, k, true, true, true, idx->n_ranks);
        if (rc) return rc;
        ResultLayout l = result_layout(nq, k);
        const bool host_block = zero_copy_results(l.total);  // (rank-local choice)
        uint8_t *r_base = host_block ? ws->h_result : ws->d_result;
        uint64_t *r_rows = reinterpret_cast<uint64_t *>(r_base + l.rows_off);
        float *r_scores = reinterpret_cast<float *>(r_base + l.scores_off);
        uint32_t *r_counts = reinterpret_cast<uint32_t *>(r_base + l.counts_off);
        memcpy(ws->h_query, queries, (size_t)nq * dim * 4);
        CUDA_TRY(cudaMemcpyAsync(ws->d_query, ws->h_query, (size_t)nq * dim * 4,
                                 cudaMemcpyHostToDevice, ws->stream));
        // masks of a collective index cover this rank's own rows (local row r = bit r)
        const uint32_t *d_mask = nullptr;
        rc = shard_mask(idx, sh, *ws, mspec, 0, ws->stream, &d_mask, &mask_holds[0]);
        if (rc) return rc;
        if (int trc = time_begin(idx, *ws)) return trc;
        if (collective_uses_fused_exchange(idx, nq, k, metric, masked)) {
            rc = collective_fused(idx, sh, *ws, ws->d_query, nq, k, metric, r_rows, r_scores,
                                  r_counts, ws->stream, d_mask);
            if (rc) return rc;
            if (int trc = time_end(*ws)) return trc;
        } else {
            if (sh.rows == 0) {
                CUDA_TRY(cudaMemsetAsync(ws->d_hits, 0, (size_t)nq * k * sizeof(nm::ShardHit),
                                         ws->stream));
            } else if (tc_usable(idx, sh, nq, k, metric, false)) {
                // this shard's hits through the tensor-core pre-filter (a rank-local choice:
                // the exchange format is the same); masked / filtered batches too
                rc = scan_queries_tc_hits_enqueue(idx, sh, *ws, ws->d_query, nq, k, metric,
                                                  idx->comm_row_base, ws->d_hits, ws->stream, d_mask);
                if (rc) return rc;
                rc = scan_queries_tc_hits_finish(idx, sh, *ws, ws->d_query, nq, k, metric,
                                                 idx->comm_row_base, ws->d_hits, ws->stream, d_mask);
                if (rc) return rc;
                idx->tc_queries += nq;
            } else if (masked) {
                rc = masked_hits(idx, sh, *ws, ws->d_query, nq, k, metric, idx->comm_row_base,
                                 ws->d_hits, ws->stream, d_mask);
                if (rc) return rc;
            } else {
                rc = scan_queries(idx, sh, *ws, ws->d_query, nq, k, metric, idx->comm_row_base,
                                  nullptr, nullptr, nullptr, ws->d_hits, ws->stream);
                if (rc) return rc;
            }
            if (int trc = time_end(*ws)) return trc;
            NM_TRACE("exchange_merge");
            NCCL_TRY(nccl().AllGather(ws->d_hits, ws->d_gather,
                                      (size_t)nq * k * sizeof(nm::ShardHit), ncclChar, idx->comm,
                                      ws->stream));
            rc = launch_merge_shards(idx, ws->d_gather, nq, k, r_rows, r_scores, r_counts, ws->stream);
            if (rc) return rc;
        }
        return download_results(idx, sh, *ws, l, nq, k, out_rows, out_scores, out_counts, host_block);
    }

    // ---- several devices in this process: scan each shard, merge the per-shard top-k on
    //      the host (G*k hits), exactly ResultMerger::merge_top_k -------------------------
    std::vector<std::unique_ptr<Workspace>> wss(G);
    std::vector<char> tc_shard(G, 0);
    std::vector<const uint32_t *> d_masks(G, nullptr);
    bool any_tc = false;
    struct ReleaseAll {
        nm_index *idx;
        std::vector<std::unique_ptr<Workspace>> &w;
        ~ReleaseAll() {
            for (size_t s = 0; s < w.size(); ++s) ws_release(*idx->shards[s], w[s]);
        }
    } rel{idx, wss};
    for (size_t s = 0; s < G; ++s) {
        Shard &sh = *idx->shards[s];
        CUDA_TRY(cudaSetDevice(sh.device));
        rc = ws_acquire(sh, wss[s]);
        if (rc) return rc;
        Workspace &ws = *wss[s];
        rc = ws_ensure(ws, sh, dim, nq, k, true, false, true, 0);
        if (rc) return rc;
        memcpy(ws.h_query, queries, (size_t)nq * dim * 4);
        CUDA_TRY(cudaMemcpyAsync(ws.d_query, ws.h_query, (size_t)nq * dim * 4,
                                 cudaMemcpyHostToDevice, ws.stream));
        const uint32_t *d_mask = nullptr;
        rc = shard_mask(idx, sh, ws, mspec, sh.row_base, ws.stream, &d_mask, &mask_holds[s]);
        if (rc) return rc;
        d_masks[s] = d_mask;
        if (int trc = time_begin(idx, ws)) return trc;
        if (sh.rows == 0) {
            CUDA_TRY(cudaMemsetAsync(ws.d_hits, 0, (size_t)nq * k * sizeof(nm::ShardHit), ws.stream));
        } else if (tc_usable(idx, sh, nq, k, metric, false)) {
            tc_shard[s] = true;
            any_tc = true;
            rc = scan_queries_tc_hits_enqueue(idx, sh, ws, ws.d_query, nq, k, metric, sh.row_base,
                                              ws.d_hits, ws.stream, d_mask);
            if (rc) return rc;
            continue;  // finished below, once every shard has its work in flight
        } else if (masked) {
            rc = masked_hits(idx, sh, ws, ws.d_query, nq, k, metric, sh.row_base, ws.d_hits, ws.stream,
                             d_mask);
            if (rc) return rc;
        } else {
            rc = scan_queries(idx, sh, ws, ws.d_query, nq, k, metric, sh.row_base, nullptr, nullptr,
                              nullptr, ws.d_hits, ws.stream);
            if (rc) return rc;
        }
        if (int trc = time_end(ws)) return trc;
        CUDA_TRY(cudaMemcpyAsync(ws.h_hits, ws.d_hits, (size_t)nq * k * sizeof(nm::ShardHit),
                                 cudaMemcpyDeviceToHost, ws.stream));
    }
    for (size_t s = 0; s < G; ++s) {
        if (!tc_shard[s]) continue;
        Shard &sh = *idx->shards[s];
        Workspace &ws = *wss[s];
        CUDA_TRY(cudaSetDevice(sh.device));
        rc = scan_queries_tc_hits_finish(idx, sh, ws, ws.d_query, nq, k, metric, sh.row_base,
                                         ws.d_hits, ws.stream, d_masks[s]);
        if (rc) return rc;
        if (int trc = time_end(ws)) return trc;
        CUDA_TRY(cudaMemcpyAsync(ws.h_hits, ws.d_hits, (size_t)nq * k * sizeof(nm::ShardHit),
                                 cudaMemcpyDeviceToHost, ws.stream));
    }
    if (any_tc) idx->tc_queries += nq;
    float max_ms = 0.f;
    for (size_t s = 0; s < G; ++s) {
        CUDA_TRY(cudaSetDevice(idx->shards[s]->device));
        rc = wait_stream(wss[s]->stream);
        if (rc) return rc;
        float ms = 0.f;
        if (int trc = time_read(*wss[s], &ms)) return trc;
        max_ms = std::max(max_ms, ms);
    }
    idx->last_scan_ms = max_ms;
    std::vector<HostHit> all;
    for (uint32_t q = 0; q < nq; ++q) {
        all.clear();
        for (size_t s = 0; s < G; ++s) {
            const nm::ShardHit *h = wss[s]->h_hits + (size_t)q * k;
            for (uint32_t i = 0; i < k; ++i) {
                if (h[i].ord == 0 && h[i].score_bits == 0) continue;
                all.push_back(HostHit{h[i].ord, h[i].score_bits, h[i].global_row, all.size()});
            }
        }
        std::stable_sort(all.begin(), all.end(),
                         [](const HostHit &a, const HostHit &b) { return a.ord > b.ord; });
        uint32_t m = (uint32_t)std::min<size_t>(k, all.size());
        out_counts[q] = m;
        for (uint32_t i = 0; i < m; ++i) {
            out_rows[(size_t)q * k + i] = all[i].row;
            memcpy(&out_scores[(size_t)q * k + i], &all[i].score_bits, 4);
        }
    }
    uint64_t rows = idx->total_rows();
    idx->searches += nq;
    idx->rows_scanned += (uint64_t)nq * rows;
    idx->bytes_streamed += (uint64_t)nq * rows * dim * 4;
    idx->h2d_bytes += (uint64_t)nq * dim * 4 * G;
    idx->d2h_bytes += (uint64_t)nq * k * sizeof(nm::ShardHit) * G;
    return NM_OK;
}

// Single-query calls from concurrent host threads are coalesced: the first caller becomes the
// leader and runs its query at once; callers arriving while the GPU is busy queue up, and the
// next leader round takes every queued query with the same (k, metric) through ONE search_impl
// call (>= 2 of them share a corpus pass in the batched kernels).  Results are bit-identical
// to isolated calls (tests/test_gpu_engine.py::test_concurrent_searches_are_coalesced).
static int search_coalesced(nm_index *idx, const float *query, uint32_t k, int metric,
                            uint64_t *out_rows, float *out_scores, uint32_t *out_count) {
    nm_index::PendingSearch me;
    me.query = query;
    me.k = k;
    me.metric = metric;
    me.out_rows = out_rows;
    me.out_scores = out_scores;
    me.out_count = out_count;
    std::unique_lock<std::mutex> lk(idx->co_mu);
    idx->co_pending.push_back(&me);
    for (;;) {
        if (me.done) break;
        if (idx->co_leader) {  // someone else is driving: wait for my result or for my turn
            idx->co_cv.wait(lk);
            continue;
        }
        // become the leader for one round: everything queued with my (k, metric)
        std::vector<nm_index::PendingSearch *> batch;
        const size_t cap = (size_t)std::max(1, idx->coalesce_max.load());
        batch.reserve(std::min(cap, idx->co_pending.size()));  // (may throw: not the leader yet)
        idx->co_leader = true;
        for (auto it = idx->co_pending.begin(); it != idx->co_pending.end() && batch.size() < cap;) {
            if ((*it)->k == me.k && (*it)->metric == me.metric) {
                batch.push_back(*it);
                it = idx->co_pending.erase(it);
            } else {
                ++it;
            }
        }
        lk.unlock();
        const uint32_t nb = (uint32_t)batch.size();
        const uint32_t dim = idx->dim;
        int rc;
        // nothing may leave this block by exception: co_leader has to be handed back, and the C
        // ABI never throws (std::bad_alloc from the staging vectors becomes NM_ERR_STORAGE)
        try {
            if (nb == 1) {
                rc = search_impl(idx, batch[0]->query, 1, k, metric, MaskSpec(), batch[0]->out_rows,
                                 batch[0]->out_scores, batch[0]->out_count);
            } else {
                std::vector<float> qs((size_t)nb * dim);
                std::vector<uint64_t> rows((size_t)nb * k);
                std::vector<float> scores((size_t)nb * k);
                std::vector<uint32_t> counts(nb);
                for (uint32_t i = 0; i < nb; ++i)
                    memcpy(&qs[(size_t)i * dim], batch[i]->query, (size_t)dim * 4);
                rc = search_impl(idx, qs.data(), nb, k, metric, MaskSpec(), rows.data(), scores.data(),
                                 counts.data());
                if (rc == NM_OK)
                    for (uint32_t i = 0; i < nb; ++i) {
                        *batch[i]->out_count = counts[i];
                        memcpy(batch[i]->out_rows, &rows[(size_t)i * k], (size_t)counts[i] * 8);
                        memcpy(batch[i]->out_scores, &scores[(size_t)i * k], (size_t)counts[i] * 4);
                    }
            }
        } catch (const std::exception &e) {
            rc = fail(NM_ERR_STORAGE, "coalesced search failed: %s", e.what());
        } catch (...) {
            rc = fail(NM_ERR_STORAGE, "coalesced search failed");
        }
        const std::string err = rc ? std::string(nmi::last_error()) : std::string();
        idx->co_batches++;
        idx->co_queries += nb;
        lk.lock();
        for (auto *r : batch) {
            r->rc = rc;
            r->error = err;
            r->done = true;
        }
        idx->co_leader = false;
        idx->co_cv.notify_all();
    }
    lk.unlock();
    if (me.rc) return fail(me.rc, "%s", me.error.c_str());
    return NM_OK;
}

int nm_search(nm_index *idx, const float *queries, uint32_t nq, uint32_t k, int metric,
              uint64_t *out_rows, float *out_scores, uint32_t *out_counts) {
    if (idx && nq == 1 && idx->coalesce_max.load() > 1 && queries && out_rows && out_scores &&
        out_counts && k > 0 && metric >= 0 && metric <= 2 && idx->comm == nullptr)
        return search_coalesced(idx, queries, k, metric, out_rows, out_scores, out_counts);
    return search_impl(idx, queries, nq, k, metric, MaskSpec(), out_rows, out_scores, out_counts);
}

int nm_debug_tc_dots(nm_index *idx, const float *queries, uint32_t nq, int32_t *out) {
    if (!idx || !queries || !out || nq == 0) return fail(NM_ERR_INVALID_ARGUMENT, "null argument");
    std::shared_lock<std::shared_mutex> g(idx->mu);
    if (idx->shards.size() != 1 || idx->comm)
        return fail(NM_ERR_CONFIGURATION, "nm_debug_tc_dots needs a single-device index");
    Shard &sh = *idx->shards[0];
    nq = std::min<uint32_t>(nq, 256u);
    if (!sh.tmap8_valid || sh.q8_rows != sh.rows || sh.rows == 0)
        return fail(NM_ERR_CONFIGURATION, "nm_debug_tc_dots needs the int8 copy (nm_index_set_prefilter(idx, 1))");
    CUDA_TRY(cudaSetDevice(sh.device));
    std::unique_ptr<Workspace> ws;
    int rc = ws_acquire(sh, ws);
    if (rc) return rc;
    struct Releaser {
        Shard &s;
        std::unique_ptr<Workspace> &w;
        ~Releaser() { ws_release(s, w); }
    } rel{sh, ws};
    const uint32_t k = 1;
    rc = ws_ensure(*ws, sh, idx->dim, nq, k, true, true, false, 0);
    if (rc) return rc;
    ResultLayout l = result_layout(nq, k);
    memcpy(ws->h_query, queries, (size_t)nq * idx->dim * 4);
    CUDA_TRY(cudaMemcpyAsync(ws->d_query, ws->h_query, (size_t)nq * idx->dim * 4,
                             cudaMemcpyHostToDevice, ws->stream));
    int *d_dots = nullptr;
    CUDA_TRY(cudaMalloc(&d_dots, (size_t)nq * sh.rows * 4));
    rc = scan_queries_tc(idx, sh, *ws, ws->d_query, nq, k, NM_DOT_PRODUCT, sh.row_base,
                         reinterpret_cast<uint64_t *>(ws->d_result + l.rows_off),
                         reinterpret_cast<float *>(ws->d_result + l.scores_off),
                         reinterpret_cast<uint32_t *>(ws->d_result + l.counts_off), ws->stream,
                         d_dots);
    if (rc == NM_OK) {
        cudaError_t e = cudaMemcpyAsync(out, d_dots, (size_t)nq * sh.rows * 4,
                                        cudaMemcpyDeviceToHost, ws->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ws->stream);
        if (e != cudaSuccess)
            rc = fail(NM_ERR_STORAGE, "CUDA error %s in nm_debug_tc_dots: %s", cudaGetErrorName(e),
                      cudaGetErrorString(e));
    }
    cudaFree(d_dots);
    return rc;
}

int nm_index_set_coalescing(nm_index *idx, int max_batch) {
    if (!idx) return fail(NM_ERR_INVALID_ARGUMENT, "null index");
    if (max_batch < 1) return fail(NM_ERR_INVALID_ARGUMENT, "max_batch must be >= 1");
    idx->coalesce_max = max_batch;
    return NM_OK;
}

int nm_search_masked(nm_index *idx, const float *queries, uint32_t nq, uint32_t k, int metric,
                     const uint64_t *row_mask, uint64_t *out_rows, float *out_scores,
                     uint32_t *out_counts) {
    if (!row_mask) return fail(NM_ERR_INVALID_ARGUMENT, "null row mask");
    MaskSpec spec;
    spec.host_mask = row_mask;
    return search_impl(idx, queries, nq, k, metric, spec, out_rows, out_scores, out_counts);
}

int nm_search_filtered(nm_index *idx, const float *queries, uint32_t nq, uint32_t k, int metric,
                       const nm_filter_op *program, uint32_t n_ops, const uint32_t *tables,
                       uint32_t n_table_words, uint64_t *out_rows, float *out_scores,
                       uint32_t *out_counts) {
    int rc = validate_filter_program(program, n_ops, tables, n_table_words);
    if (rc) return rc;
    MaskSpec spec;
    spec.prog = program;
    spec.n_ops = n_ops;
    spec.tables = tables;
    spec.n_table_words = n_table_words;
    return search_impl(idx, queries, nq, k, metric, spec, out_rows, out_scores, out_counts);
}

int nm_search_device(nm_index *idx, const float *d_queries, uint32_t nq, uint32_t k, int metric,
                     uint64_t *d_out_rows, float *d_out_scores, uint32_t *d_out_counts,
                     void *stream_v) {
    NM_TRACE("nm_search_device");
    int rc = validate_search(idx, d_queries, nq, k, metric, d_out_rows, d_out_scores, d_out_counts);
    if (rc) return rc;
    if (nq >= 2 && idx->comm == nullptr) {
        rc = q8_auto_prepare(idx, nq, k);  // auto mode: first eligible batch builds the int8 copy
        if (rc) return rc;
    }
    std::shared_lock<std::shared_mutex> g(idx->mu);
    if (idx->shards.size() != 1)
        return fail(NM_ERR_CONFIGURATION, "nm_search_device needs a single-device index");
    Shard &sh = *idx->shards[0];
    const uint32_t dim = idx->dim;
    const bool collective = idx->comm != nullptr;
    CUDA_TRY(cudaSetDevice(sh.device));
    // Caller stream: use the workspace bound to that stream and return without synchronising
    // (stream order protects the scratch).  NULL stream: pooled workspace + synchronise.
    std::unique_ptr<Workspace> pooled;
    Workspace *ws = nullptr;
    cudaStream_t stream = nullptr;
    if (stream_v) {
        stream = (cudaStream_t)stream_v;
        std::lock_guard<std::mutex> pg(sh.pool_mu);
        for (auto &e : sh.stream_ws)
            if (e.first == stream) ws = e.second.get();
        if (!ws) {
            std::unique_ptr<Workspace> nw(new Workspace());
            nw->device = sh.device;
            CUDA_TRY(cudaMalloc(&nw->d_counter, kWsCounterWords * sizeof(uint32_t)));
            CUDA_TRY(cudaMemsetAsync(nw->d_counter, 0, kWsCounterWords * sizeof(uint32_t), stream));
            CUDA_TRY(cudaEventCreateWithFlags(&nw->async_done, cudaEventDisableTiming));
            ws = nw.get();
            sh.stream_ws.emplace_back(stream, std::move(nw));
        }
    } else {
        rc = ws_acquire(sh, pooled);
        if (rc) return rc;
        ws = pooled.get();
        stream = ws->stream;
    }
    struct Releaser {
        Shard &s;
        std::unique_ptr<Workspace> &w;
        ~Releaser() { ws_release(s, w); }
    } rel{sh, pooled};
    // growing a stream-bound workspace frees buffers earlier launches may still read
    {
        size_t need_cand = (size_t)sh.sm_count * std::min<uint32_t>(k, nm::kMaxFastK),
               need_hits = (size_t)nq * k;
        bool grow = ws->cand_cap < need_cand ||
                    (collective && (ws->hits_cap < need_hits ||
                                    ws->gather_cap < need_hits * (size_t)idx->n_ranks));
        if (grow && stream_v) CUDA_TRY(cudaStreamSynchronize(stream));
    }
    rc = ws_ensure(*ws, sh, dim, nq, k, false, false, collective, collective ? idx->n_ranks : 0);
    if (rc) return rc;
    // nm_index_set_pipelining: single-query scans of consecutive asynchronous calls overlap
    // (programmatic dependent launch); profiling events between the launches would undo that
    const bool pipelined = stream_v && idx->pipelining.load() && !idx->profiling.load();
    ws->pipeline_next = pipelined;
    struct PipeReset {
        Workspace *w;
        ~PipeReset() { w->pipeline_next = false; }
    } pipe_reset{ws};
    // The peer exchange and NCCL order collective searches by ONE stream: when the caller
    // switches streams, the earlier stream's searches are finished first (called with comm_mu
    // held, in the same critical section that enqueues the search).
    auto adopt_stream = [&]() -> int {
        if (!stream_v) return NM_OK;
        if (idx->xchg_async_stream && idx->xchg_async_stream != stream)
            CUDA_TRY(cudaStreamSynchronize(idx->xchg_async_stream));
        idx->xchg_async_stream = stream;
        return NM_OK;
    };
    std::pair<cudaEvent_t, cudaEvent_t> *prof = nullptr;
    if (idx->profiling.load() && sh.rows) {
        if (ws->prof_used == ws->prof_events.size()) {
            cudaEvent_t a, b;
            CUDA_TRY(cudaEventCreate(&a));
            CUDA_TRY(cudaEventCreate(&b));
            ws->prof_events.emplace_back(a, b);
        }
        prof = &ws->prof_events[ws->prof_used++];
        CUDA_TRY(cudaEventRecord(prof->first, stream));
    }
    if (!collective) {
        if (sh.rows == 0) {
            CUDA_TRY(cudaMemsetAsync(d_out_counts, 0, (size_t)nq * 4, stream));
        } else if (tc_usable(idx, sh, nq, k, metric, false)) {
            // batches: tensor-core pre-filter, then the exact kernels as CONDITIONAL launches that
            // only run for queries the device flagged (nothing is read back to the host here)
            rc = scan_queries_tc(idx, sh, *ws, d_queries, nq, k, metric, sh.row_base, d_out_rows,
                                 d_out_scores, d_out_counts, stream);
            if (rc) return rc;
            uint32_t *d_redo = nullptr;
            rc = tc_redo_flags(*ws, nq, (uint32_t)sh.rows, &d_redo, stream);
            if (rc) return rc;
            rc = scan_queries(idx, sh, *ws, d_queries, nq, k, metric, sh.row_base, d_out_rows,
                              d_out_scores, d_out_counts, nullptr, stream, d_redo);
            if (rc) return rc;
            idx->tc_queries += nq;
        } else {
            rc = scan_queries(idx, sh, *ws, d_queries, nq, k, metric, sh.row_base, d_out_rows,
                              d_out_scores, d_out_counts, nullptr, stream);
            if (rc) return rc;
        }
        if (prof) CUDA_TRY(cudaEventRecord(prof->second, stream));
    } else if (collective_uses_fused_exchange(idx, nq, k, metric)) {
        // (collective calls on one index must be issued in the same order on every rank)
        std::lock_guard<std::mutex> cg(idx->comm_mu);
        rc = adopt_stream();
        if (rc) return rc;
        rc = collective_fused(idx, sh, *ws, d_queries, nq, k, metric, d_out_rows, d_out_scores,
                              d_out_counts, stream);
        if (rc) return rc;
        if (prof) CUDA_TRY(cudaEventRecord(prof->second, stream));
    } else {
        std::lock_guard<std::mutex> cg(idx->comm_mu);
        rc = adopt_stream();
        if (rc) return rc;
        if (sh.rows == 0) {
            CUDA_TRY(cudaMemsetAsync(ws->d_hits, 0, (size_t)nq * k * sizeof(nm::ShardHit), stream));
        } else {
            rc = scan_queries(idx, sh, *ws, d_queries, nq, k, metric, idx->comm_row_base, nullptr,
                              nullptr, nullptr, ws->d_hits, stream);
            if (rc) return rc;
        }
        if (prof) CUDA_TRY(cudaEventRecord(prof->second, stream));
        NCCL_TRY(nccl().AllGather(ws->d_hits, ws->d_gather,
                                  (size_t)nq * k * sizeof(nm::ShardHit), ncclChar, idx->comm,
                                  stream));
        rc = launch_merge_shards(idx, ws->d_gather, nq, k, d_out_rows, d_out_scores, d_out_counts,
                                 stream);
        if (rc) return rc;
    }
    if (!stream_v) {
        CUDA_TRY(cudaStreamSynchronize(stream));
    } else if (!pipelined) {
        // lets mutations wait for this call (wait_async_searches)
        std::lock_guard<std::mutex> pg(sh.pool_mu);
        CUDA_TRY(cudaEventRecord(ws->async_done, stream));
        ws->async_pending = true;
    }
    idx->searches += nq;
    idx->rows_scanned += (uint64_t)nq * sh.rows;
    idx->bytes_streamed += (uint64_t)nq * sh.rows * dim * 4;
    return NM_OK;
}

int nm_index_set_pipelining(nm_index *idx, int enable) {
    if (!idx) return fail(NM_ERR_INVALID_ARGUMENT, "null index");
    idx->pipelining = enable ? 1 : 0;
    return NM_OK;
}

int nm_index_release_stream(nm_index *idx, void *stream_v) {
    if (!idx) return fail(NM_ERR_INVALID_ARGUMENT, "null index");
    std::unique_lock<std::shared_mutex> g(idx->mu);
    cudaStream_t stream = (cudaStream_t)stream_v;
    for (auto &shp : idx->shards) {
        Shard &sh = *shp;
        CUDA_TRY(cudaSetDevice(sh.device));
        std::unique_ptr<Workspace> victim;
        {
            std::lock_guard<std::mutex> pg(sh.pool_mu);
            for (auto it = sh.stream_ws.begin(); it != sh.stream_ws.end(); ++it)
                if (it->first == stream) {
                    victim = std::move(it->second);
                    sh.stream_ws.erase(it);
                    break;
                }
        }
        if (victim) CUDA_TRY(cudaStreamSynchronize(stream));  // its scratch may still be in use
    }
    {
        std::lock_guard<std::mutex> cg(idx->comm_mu);
        if (idx->xchg_async_stream == stream) idx->xchg_async_stream = nullptr;
    }
    return NM_OK;
}

}  // extern "C"
