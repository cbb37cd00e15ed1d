// nm_index.cu — C ABI (include/neumann_b200.h) over the sm_100a scan kernels.
//
// Owns: the device mirror of the reference's `emb:` rows (row-major f32, pitch = dim rounded up
// to 4 floats so every row is 16-byte aligned for TMA), pinned double-buffered staging for
// host -> device loads, per-call workspaces (stream, query, candidates, results) so that
// nm_search is re-entrant, and the optional NCCL communicator for row-range sharding across
// processes.  There is deliberately no CPU code path for the scan.
#include "../../include/neumann_b200.h"
#include "scan_kernels.cuh"
#include "batch_kernels.cuh"
#include "prefilter_kernels.cuh"

#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <memory>
#include <mutex>
#include <shared_mutex>
#include <string>
#include <thread>
#include <vector>

namespace {

thread_local std::string g_last_error;

int fail(int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return code;
}

#define CUDA_TRY(expr)                                                                     \
    do {                                                                                   \
        cudaError_t _e = (expr);                                                           \
        if (_e != cudaSuccess)                                                             \
            return fail(NM_ERR_STORAGE, "CUDA error %s at %s:%d: %s", cudaGetErrorName(_e), \
                        __FILE__, __LINE__, cudaGetErrorString(_e));                       \
    } while (0)

// ---- cuTensorMapEncodeTiled through the runtime (no link-time libcuda dependency) --------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *,
                                  const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                  const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_tiled() {
    static EncodeTiledFn fn = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) !=
                cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}

// ---- NCCL, loaded lazily so the library also loads on hosts without it -------------------
struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t,
                              cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};

NcclApi &nccl() {
    static NcclApi api = [] {
        NcclApi a;
        a.handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!a.handle) a.handle = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!a.handle) return a;
        a.GetUniqueId = (decltype(a.GetUniqueId))dlsym(a.handle, "ncclGetUniqueId");
        a.CommInitRank = (decltype(a.CommInitRank))dlsym(a.handle, "ncclCommInitRank");
        a.CommDestroy = (decltype(a.CommDestroy))dlsym(a.handle, "ncclCommDestroy");
        a.AllGather = (decltype(a.AllGather))dlsym(a.handle, "ncclAllGather");
        a.GetErrorString = (decltype(a.GetErrorString))dlsym(a.handle, "ncclGetErrorString");
        a.ok = a.GetUniqueId && a.CommInitRank && a.CommDestroy && a.AllGather &&
               a.GetErrorString;
        return a;
    }();
    return api;
}

#define NCCL_TRY(expr)                                                                  \
    do {                                                                                \
        ncclResult_t _r = (expr);                                                       \
        if (_r != ncclSuccess)                                                          \
            return fail(NM_ERR_STORAGE, "NCCL error at %s:%d: %s", __FILE__, __LINE__,  \
                        nccl().GetErrorString(_r));                                     \
    } while (0)

constexpr size_t kStagingBytes = 64u << 20;  // per pinned staging buffer (two per shard)

// Pageable -> pinned staging copy, split over a few host threads (one core tops out near
// 11 GB/s, well below what the DMA engine takes from pinned memory).
void parallel_memcpy(void *dst, const void *src, size_t bytes) {
    unsigned hw = std::thread::hardware_concurrency();
    unsigned n = std::min<unsigned>(8u, hw ? hw : 1u);
    if (bytes < (4u << 20) || n < 2) {
        memcpy(dst, src, bytes);
        return;
    }
    std::vector<std::thread> th;
    const size_t per = ((bytes / n) + 4095) & ~size_t(4095);
    for (unsigned i = 1; i < n; ++i) {
        size_t off = per * i;
        if (off >= bytes) break;
        size_t len = std::min(per, bytes - off);
        th.emplace_back([=] { memcpy((char *)dst + off, (const char *)src + off, len); });
    }
    memcpy(dst, src, std::min(per, bytes));
    for (auto &t : th) t.join();
}

// Per-call scratch on one device.  Pooled per shard so concurrent nm_search calls never share
// a stream, a candidate buffer or the "last CTA" ticket.
struct Workspace {
    int device = -1;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    float *d_query = nullptr;
    float *h_query = nullptr;  // pinned
    size_t query_cap = 0;      // floats
    uint64_t *d_cand = nullptr;
    size_t cand_cap = 0;  // keys
    uint32_t *d_counter = nullptr;
    uint32_t *d_mask = nullptr;       // row bitmask of a pre-filtered search, padded to row blocks
    size_t mask_cap = 0;              // u32 words
    uint64_t *d_pass_keys = nullptr;  // [k] merged keys of the chained passes (k > 1024)
    size_t pass_keys_cap = 0;
    // packed result block: [counts u32 x nq (8-aligned)] [rows u64 x nq*k] [scores f32 x nq*k]
    uint8_t *d_result = nullptr;
    uint8_t *h_result = nullptr;  // pinned
    size_t result_cap = 0;
    nm::ShardHit *d_hits = nullptr;  // [nq, k] this shard's hits
    nm::ShardHit *h_hits = nullptr;  // pinned
    size_t hits_cap = 0;
    nm::ShardHit *d_gather = nullptr;  // [n_ranks, nq, k]
    size_t gather_cap = 0;
    // pre-filter path (prefilter_kernels.cuh)
    nm::KeptEntry *d_kept = nullptr;
    uint64_t *d_exact_keys = nullptr;
    uint32_t *d_pf_ctl = nullptr;   // [nq][8]
    uint32_t *h_pf_ctl = nullptr;   // pinned
    size_t pf_ctl_cap = 0;          // queries
    // batched-query path (batch_kernels.cuh)
    float *d_qt = nullptr;         // [n_kc][32][QB] transposed query chunks
    size_t qt_cap = 0;             // floats
    float *d_qmag = nullptr;       // [64]
    float *d_scores = nullptr;     // [QB][score_stride]
    size_t scores_cap = 0;         // floats
    uint64_t *d_bcand = nullptr;   // [QB][ctas_per_query][k]
    size_t bcand_cap = 0;          // keys
    uint32_t *d_bctl = nullptr;    // [0] cursor [1] done [2..2+64) tickets
    // profiling ring (nm_index_set_profiling): event pairs around the scan launches of
    // asynchronous nm_search_device calls, resolved lazily by nm_index_stats
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_events;
    size_t prof_used = 0;

    ~Workspace() {
        if (device < 0) return;
        cudaSetDevice(device);
        if (d_query) cudaFree(d_query);
        if (h_query) cudaFreeHost(h_query);
        if (d_cand) cudaFree(d_cand);
        if (d_counter) cudaFree(d_counter);
        if (d_mask) cudaFree(d_mask);
        if (d_pass_keys) cudaFree(d_pass_keys);
        if (d_kept) cudaFree(d_kept);
        if (d_exact_keys) cudaFree(d_exact_keys);
        if (d_pf_ctl) cudaFree(d_pf_ctl);
        if (h_pf_ctl) cudaFreeHost(h_pf_ctl);
        if (d_qt) cudaFree(d_qt);
        if (d_qmag) cudaFree(d_qmag);
        if (d_scores) cudaFree(d_scores);
        if (d_bcand) cudaFree(d_bcand);
        if (d_bctl) cudaFree(d_bctl);
        if (d_result) cudaFree(d_result);
        if (h_result) cudaFreeHost(h_result);
        if (d_hits) cudaFree(d_hits);
        if (h_hits) cudaFreeHost(h_hits);
        if (d_gather) cudaFree(d_gather);
        if (ev0) cudaEventDestroy(ev0);
        if (ev1) cudaEventDestroy(ev1);
        for (auto &pe : prof_events) {
            cudaEventDestroy(pe.first);
            cudaEventDestroy(pe.second);
        }
        if (stream) cudaStreamDestroy(stream);
    }
};

struct Shard {
    int device = 0;
    int sm_count = 0;
    float *d_rows = nullptr;
    uint64_t rows = 0;      // local rows
    uint64_t capacity = 0;  // local rows allocated
    uint64_t row_base = 0;  // global index of local row 0 (within this process)
    CUtensorMap tmap;
    bool tmap_valid = false;
    // int8 pre-filter copy (prefilter_kernels.cuh): [capacity8, pitch8] bytes + 16 B per row
    int8_t *d_q8 = nullptr;
    nm::RowMeta *d_meta = nullptr;
    uint32_t *d_q8_flag = nullptr;
    uint64_t q8_capacity = 0;
    uint64_t q8_rows = 0;  // rows [0, q8_rows) are quantised and current
    CUtensorMap tmap8;
    bool tmap8_valid = false;
    cudaStream_t copy_stream = nullptr;
    float *staging[2] = {nullptr, nullptr};
    cudaEvent_t staging_done[2] = {nullptr, nullptr};
    std::mutex pool_mu;
    std::vector<std::unique_ptr<Workspace>> pool;
    // Workspaces bound to a caller stream (nm_search_device): work on one stream is ordered,
    // so the same scratch can be reused by consecutive asynchronous calls without a sync.
    std::vector<std::pair<cudaStream_t, std::unique_ptr<Workspace>>> stream_ws;
};

}  // namespace

struct nm_index {
    uint32_t dim = 0;
    uint32_t pitch = 0;  // floats per row in device memory
    std::vector<std::unique_ptr<Shard>> shards;
    mutable std::shared_mutex mu;
    // cross-process sharding
    ncclComm_t comm = nullptr;
    int n_ranks = 1, rank = 0;
    uint64_t comm_row_base = 0;
    std::mutex comm_mu;  // collective searches are issued one at a time, in call order
    // peer-memory exchange (CUDA IPC): the fused single-query path writes hits straight into
    // the peers' mailboxes; NCCL stays for the batched / k > 1024 paths and as a fallback
    void *xchg_mem = nullptr;                    // local [flags 256 B | mailbox]
    void *xchg_peer[nm::kMaxRanks] = {nullptr};  // mapped peer buffers (own rank = xchg_mem)
    bool xchg_ok = false;
    uint32_t xchg_seq = 0;
    // counters
    std::atomic<uint64_t> searches{0}, rows_scanned{0}, bytes_streamed{0}, scan_launches{0},
        merge_launches{0}, h2d_bytes{0}, d2h_bytes{0};
    std::atomic<double> last_scan_ms{0.0};
    std::atomic<int> profiling{0};
    std::atomic<int> prefilter{0};  // nm_index_set_prefilter: 1 = exact int8 pre-filter
    std::atomic<uint64_t> pf_queries{0}, pf_fallbacks{0}, pf_kept{0};
    std::atomic<int> batching{1};  // nm_index_set_batching: 0 forces one scan per query
    double profiled_scan_ms = 0.0;  // guarded by mu (exclusive) in nm_index_stats
    uint64_t profiled_scans = 0;
    uint64_t total_rows() const {
        uint64_t n = 0;
        for (auto &s : shards) n += s->rows;
        return n;
    }
};

namespace {

size_t scan_smem_bytes(uint32_t n_stages, uint32_t q_floats) {
    return 1024 + (size_t)n_stages * nm::kStageBytes + (size_t)nm::kCandCap * 8 +
           (size_t)q_floats * 4 + 2 * nm::kMaxStages * 8 + 64 + nm::kMaxStages * 4;
}

int build_tmap(nm_index *idx, Shard &sh) {
    sh.tmap_valid = false;
    if (sh.rows == 0) return NM_OK;
    EncodeTiledFn enc = get_encode_tiled();
    if (!enc) return fail(NM_ERR_STORAGE, "cuTensorMapEncodeTiled unavailable (driver too old?)");
    cuuint64_t gdim[2] = {idx->dim, sh.rows};
    cuuint64_t gstride[1] = {(cuuint64_t)idx->pitch * 4};
    cuuint32_t box[2] = {nm::kChunkFloats, nm::kRowsPerBlock};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&sh.tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, sh.d_rows, gdim, gstride, box,
                     estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        return fail(NM_ERR_STORAGE, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    sh.tmap_valid = true;
    return NM_OK;
}


// ---- int8 pre-filter copy upkeep ---------------------------------------------------------
uint32_t q8_pitch(uint32_t dim) { return (dim + 15u) & ~15u; }

int build_tmap8(nm_index *idx, Shard &sh) {
    sh.tmap8_valid = false;
    if (sh.rows == 0 || !sh.d_q8) return NM_OK;
    EncodeTiledFn enc = get_encode_tiled();
    if (!enc) return fail(NM_ERR_STORAGE, "cuTensorMapEncodeTiled unavailable (driver too old?)");
    cuuint64_t gdim[2] = {idx->dim, sh.rows};
    cuuint64_t gstride[1] = {(cuuint64_t)q8_pitch(idx->dim)};
    cuuint32_t box[2] = {128, nm::kRowsPerBlock};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&sh.tmap8, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, sh.d_q8, gdim, gstride, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        return fail(NM_ERR_STORAGE, "cuTensorMapEncodeTiled (int8) failed with CUresult %d", (int)r);
    sh.tmap8_valid = true;
    return NM_OK;
}

// Bring the int8 copy of rows [first, first+n) up to date (call with the index write lock held,
// after the f32 mirror holds the new data).  No-op while the pre-filter is off.
int q8_refresh(nm_index *idx, Shard &sh, uint64_t first, uint64_t n) {
    if (!idx->prefilter.load()) return NM_OK;
    const uint32_t pitch8 = q8_pitch(idx->dim);
    if (sh.rows > sh.q8_capacity) {
        uint64_t cap = std::max<uint64_t>(sh.rows, sh.capacity);
        int8_t *nq = nullptr;
        nm::RowMeta *nmeta = nullptr;
        CUDA_TRY(cudaMalloc(&nq, cap * pitch8));
        CUDA_TRY(cudaMalloc(&nmeta, cap * sizeof(nm::RowMeta)));
        if (sh.d_q8 && sh.q8_rows) {
            CUDA_TRY(cudaMemcpyAsync(nq, sh.d_q8, sh.q8_rows * pitch8, cudaMemcpyDeviceToDevice,
                                     sh.copy_stream));
            CUDA_TRY(cudaMemcpyAsync(nmeta, sh.d_meta, sh.q8_rows * sizeof(nm::RowMeta),
                                     cudaMemcpyDeviceToDevice, sh.copy_stream));
            CUDA_TRY(cudaStreamSynchronize(sh.copy_stream));
        }
        if (sh.d_q8) CUDA_TRY(cudaFree(sh.d_q8));
        if (sh.d_meta) CUDA_TRY(cudaFree(sh.d_meta));
        sh.d_q8 = nq;
        sh.d_meta = nmeta;
        sh.q8_capacity = cap;
    }
    if (!sh.d_q8_flag) {
        CUDA_TRY(cudaMalloc(&sh.d_q8_flag, sizeof(uint32_t)));
        CUDA_TRY(cudaMemsetAsync(sh.d_q8_flag, 0, sizeof(uint32_t), sh.copy_stream));
    }
    // anything the copy has never seen is (re)quantised together with the requested range
    if (sh.q8_rows < first) {
        n += first - sh.q8_rows;
        first = sh.q8_rows;
    }
    if (first + n > sh.rows) n = sh.rows > first ? sh.rows - first : 0;
    if (n) {
        uint32_t blocks = (uint32_t)std::min<uint64_t>((n + 7) / 8, (uint64_t)sh.sm_count * 16);
        nm::quantize_rows_kernel<<<blocks, 256, 0, sh.copy_stream>>>(
            sh.d_rows, idx->pitch, idx->dim, first, n, sh.d_q8, pitch8, sh.d_meta, sh.d_q8_flag);
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaStreamSynchronize(sh.copy_stream));
    }
    sh.q8_rows = sh.rows;
    return build_tmap8(idx, sh);
}

int shard_reserve(nm_index *idx, Shard &sh, uint64_t rows, bool keep) {
    if (rows <= sh.capacity) return NM_OK;
    if (rows > nm::kMaxLocalRows)
        return fail(NM_ERR_INVALID_ARGUMENT, "shard would hold %llu rows; limit is %u per device",
                    (unsigned long long)rows, nm::kMaxLocalRows);
    uint64_t cap = rows;
    if (keep && sh.capacity) cap = std::max<uint64_t>(rows, sh.capacity + sh.capacity / 2);
    float *p = nullptr;
    CUDA_TRY(cudaMalloc(&p, cap * idx->pitch * sizeof(float)));
    if (keep && sh.rows) {
        CUDA_TRY(cudaMemcpyAsync(p, sh.d_rows, sh.rows * idx->pitch * sizeof(float),
                                 cudaMemcpyDeviceToDevice, sh.copy_stream));
        CUDA_TRY(cudaStreamSynchronize(sh.copy_stream));
    }
    if (sh.d_rows) CUDA_TRY(cudaFree(sh.d_rows));
    sh.d_rows = p;
    sh.capacity = cap;
    return NM_OK;
}

// Host rows [n, dim] -> device rows [first, first+n) of the shard.  Pinned sources are DMA'd
// directly; pageable sources go through two pinned staging buffers so the host memcpy of
// chunk i+1 overlaps the DMA of chunk i.
int shard_upload(nm_index *idx, Shard &sh, uint64_t first, const float *src, uint64_t n) {
    if (n == 0) return NM_OK;
    const size_t row_bytes = (size_t)idx->dim * 4, pitch_bytes = (size_t)idx->pitch * 4;
    float *dst = sh.d_rows + first * idx->pitch;
    cudaPointerAttributes attr;
    bool pinned = cudaPointerGetAttributes(&attr, src) == cudaSuccess &&
                  attr.type == cudaMemoryTypeHost;
    cudaGetLastError();
    if (idx->pitch != idx->dim)
        CUDA_TRY(cudaMemsetAsync(dst, 0, n * pitch_bytes, sh.copy_stream));
    if (pinned) {
        CUDA_TRY(cudaMemcpy2DAsync(dst, pitch_bytes, src, row_bytes, row_bytes, n,
                                   cudaMemcpyHostToDevice, sh.copy_stream));
    } else {
        if (!sh.staging[0]) {
            for (int b = 0; b < 2; ++b) {
                CUDA_TRY(cudaMallocHost(&sh.staging[b], kStagingBytes));
                CUDA_TRY(cudaEventCreateWithFlags(&sh.staging_done[b], cudaEventDisableTiming));
            }
        }
        const uint64_t rows_per_chunk = std::max<uint64_t>(1, kStagingBytes / row_bytes);
        if (row_bytes > kStagingBytes)
            return fail(NM_ERR_DIMENSION_MISMATCH, "dimension %u too large for staging", idx->dim);
        int b = 0;
        for (uint64_t r = 0; r < n; r += rows_per_chunk, b ^= 1) {
            uint64_t m = std::min(rows_per_chunk, n - r);
            CUDA_TRY(cudaEventSynchronize(sh.staging_done[b]));
            parallel_memcpy(sh.staging[b], src + r * idx->dim, m * row_bytes);
            CUDA_TRY(cudaMemcpy2DAsync(dst + r * idx->pitch, pitch_bytes, sh.staging[b], row_bytes,
                                       row_bytes, m, cudaMemcpyHostToDevice, sh.copy_stream));
            CUDA_TRY(cudaEventRecord(sh.staging_done[b], sh.copy_stream));
        }
    }
    CUDA_TRY(cudaStreamSynchronize(sh.copy_stream));
    idx->h2d_bytes += n * row_bytes;
    return NM_OK;
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
    CUDA_TRY(cudaMalloc(&ws->d_counter, 2 * sizeof(uint32_t)));
    CUDA_TRY(cudaMemsetAsync(ws->d_counter, 0, 2 * sizeof(uint32_t), ws->stream));
    out = std::move(ws);
    return NM_OK;
}

void ws_release(Shard &sh, std::unique_ptr<Workspace> &ws) {
    if (!ws) return;
    std::lock_guard<std::mutex> g(sh.pool_mu);
    sh.pool.push_back(std::move(ws));
}

struct ResultLayout {
    size_t counts_off, rows_off, scores_off, total;
};
ResultLayout result_layout(uint32_t nq, uint32_t k) {
    ResultLayout l;
    l.counts_off = 0;
    l.rows_off = ((size_t)nq * 4 + 15) & ~size_t(15);
    l.scores_off = l.rows_off + (size_t)nq * k * 8;
    l.total = l.scores_off + (size_t)nq * k * 4;
    return l;
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
    size_t cand = (size_t)sh.sm_count * std::min<uint32_t>(k, nm::kMaxFastK);
    if (ws.cand_cap < cand) {
        if (ws.d_cand) CUDA_TRY(cudaFree(ws.d_cand));
        ws.cand_cap = 0;
        CUDA_TRY(cudaMalloc(&ws.d_cand, cand * 8));
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
    kern<<<grid, nm::kScanThreads, smem, stream>>>(sh.tmap, p);
    CUDA_TRY(cudaGetLastError());
    return NM_OK;
}

// One query over one shard: a single launch for k <= 1024, otherwise ceil(k/1024) chained
// passes, each admitting only keys below the previous pass's last key.
int launch_scan(nm_index *idx, const Shard &sh, Workspace &ws, const float *d_query, uint32_t k,
                int metric, uint64_t row_base, uint64_t *out_rows, float *out_scores,
                uint32_t *out_count, nm::ShardHit *out_hits, cudaStream_t stream,
                const nm::PeerXchg *xchg = nullptr, const uint32_t *d_row_mask = nullptr) {
    nm::ScanParams p;
    memset(&p, 0, sizeof(p));
    if (xchg) p.xchg = *xchg;
    p.row_mask = d_row_mask;
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
constexpr uint32_t kBatchMinQueries = 8;   // below this, nq single-query passes are cheaper
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
                         nm::ShardHit *out_hits, cudaStream_t stream) {
    const uint32_t dim = idx->dim;
    const uint32_t n_kc = (dim + 31u) / 32u;
    const uint32_t k_eff = (uint32_t)std::min<uint64_t>(k, sh.rows);
    const uint32_t n_rb = ((uint32_t)sh.rows + nm::kRowsPerBlock - 1) / nm::kRowsPerBlock;
    const uint64_t stride = ((uint64_t)sh.rows + 63u) & ~uint64_t(63);
    const uint32_t qb_max = (metric == NM_EUCLIDEAN) ? (nq > 16 ? 64u : 16u) : 16u;
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
    if (out_hits && k_eff < k)
        CUDA_TRY(cudaMemsetAsync(out_hits, 0, (size_t)nq * k * sizeof(nm::ShardHit), stream));

    for (uint32_t q0 = 0; q0 < nq; q0 += qb_max) {
        const uint32_t nqp = std::min<uint32_t>(qb_max, nq - q0);
        const uint32_t qb = (qb_max == 64u && nqp <= 16u) ? 16u : qb_max;
        nm::prepare_batch_kernel<<<std::max<uint32_t>(8u, (n_kc * 32u * qb + 255u) / 256u), 256, 0, stream>>>(
            d_queries + (size_t)q0 * dim, nqp, dim, qb, n_kc, ws.d_qt, ws.d_qmag);
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
        int rc;
        if (metric == NM_EUCLIDEAN) {
            if (qb == 64u) {
                sp.n_stages = batch_stages<64>();
                rc = launch_score_batch_t<nm::kEuclidean, 64>(sh, sp, stream);
            } else {
                sp.n_stages = batch_stages<16>();
                rc = launch_score_batch_t<nm::kEuclidean, 16>(sh, sp, stream);
            }
        } else if (metric == NM_COSINE) {
            sp.n_stages = batch_stages<16>();
            rc = launch_score_batch_t<nm::kCosine, 16>(sh, sp, stream);
        } else {
            sp.n_stages = batch_stages<16>();
            rc = launch_score_batch_t<nm::kDot, 16>(sh, sp, stream);
        }
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
           (size_t)q_words * 4 + 2 * nm::kMaxStages * 8 + 256 +
           (size_t)nm::kKeptStage * sizeof(nm::KeptEntry);
}

bool prefilter_usable(const nm_index *idx, const Shard &sh, uint32_t nq, uint32_t k, int metric,
                      const uint64_t *row_mask) {
    if (!idx->prefilter.load() || row_mask || metric == NM_EUCLIDEAN) return false;
    if (!sh.tmap8_valid || sh.q8_rows != sh.rows || sh.rows == 0) return false;
    if (k > (uint32_t)nm::kMaxFastK || k > sh.rows) return false;
    if (idx->batching.load() && nq >= kBatchMinQueries && (idx->dim % 8u) == 0) return false;
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
    p.metric = metric == NM_COSINE ? nm::kCosine : nm::kDot;
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

// nq queries over one shard: batched kernels when that pays, else one (chained) scan per query.
int scan_queries(nm_index *idx, const Shard &sh, Workspace &ws, const float *d_queries,
                 uint32_t nq, uint32_t k, int metric, uint64_t row_base, uint64_t *out_rows,
                 float *out_scores, uint32_t *out_counts, nm::ShardHit *out_hits,
                 cudaStream_t stream) {
    if ((idx->batching.load() || single_query_stages(idx->dim) < 2) &&
        batch_eligible(idx, sh, nq, k, metric)) {
        int rc = scan_queries_batched(idx, sh, ws, d_queries, nq, k, metric, row_base, out_rows,
                                      out_scores, out_counts, out_hits, stream);
        if (rc != -1) return rc;
    }
    for (uint32_t q = 0; q < nq; ++q) {
        int rc = launch_scan(idx, sh, ws, d_queries + (size_t)q * idx->dim, k, metric, row_base,
                             out_rows ? out_rows + (size_t)q * k : nullptr,
                             out_scores ? out_scores + (size_t)q * k : nullptr,
                             out_counts ? out_counts + q : nullptr,
                             out_hits ? out_hits + (size_t)q * k : nullptr, stream);
        if (rc) return rc;
    }
    return NM_OK;
}


constexpr size_t kXchgFlagBytes = 256;
size_t xchg_bytes(int n_ranks) {
    return kXchgFlagBytes + (size_t)2 * n_ranks * nm::kMaxFastK * sizeof(nm::ShardHit);
}

nm::PeerXchg make_xchg(const nm_index *idx, uint32_t seq) {
    nm::PeerXchg x;
    memset(&x, 0, sizeof(x));
    x.n_ranks = (uint32_t)idx->n_ranks;
    x.rank = (uint32_t)idx->rank;
    x.seq = seq;
    x.kcap = nm::kMaxFastK;
    for (int r = 0; r < idx->n_ranks; ++r) {
        uint8_t *base = static_cast<uint8_t *>(idx->xchg_peer[r]);
        x.flags[r] = reinterpret_cast<uint32_t *>(base);
        x.mailbox[r] = reinterpret_cast<nm::ShardHit *>(base + kXchgFlagBytes);
    }
    return x;
}

// Map every rank's exchange buffer into this process.  Failure is not fatal: the index then
// keeps using ncclAllGather + merge_shards_kernel.
void setup_peer_exchange(nm_index *idx, cudaStream_t stream) {
    idx->xchg_ok = false;
    const char *off = getenv("NM_DISABLE_PEER_EXCHANGE");
    if (off && off[0] == '1') return;
    if (idx->n_ranks < 2 || idx->n_ranks > nm::kMaxRanks) return;
    const int n = idx->n_ranks;
    bool ok = true;
    cudaIpcMemHandle_t mine;
    cudaIpcMemHandle_t *d_handles = nullptr;
    std::vector<cudaIpcMemHandle_t> all((size_t)n);
    // every rank must reach the all-gather below, so failures only clear `ok`
    if (cudaMalloc(&idx->xchg_mem, xchg_bytes(n)) != cudaSuccess) ok = false;
    if (ok && cudaMemset(idx->xchg_mem, 0, xchg_bytes(n)) != cudaSuccess) ok = false;
    memset(&mine, 0, sizeof(mine));
    if (ok && cudaIpcGetMemHandle(&mine, idx->xchg_mem) != cudaSuccess) ok = false;
    if (!ok) memset(&mine, 0, sizeof(mine));
    if (cudaMalloc(&d_handles, sizeof(mine) * n) != cudaSuccess) {
        cudaGetLastError();
        return;  // cannot even exchange: peers time out in NCCL, nothing we can do here
    }
    cudaMemcpyAsync(d_handles + idx->rank, &mine, sizeof(mine), cudaMemcpyHostToDevice, stream);
    ncclResult_t nr = nccl().AllGather(d_handles + idx->rank, d_handles, sizeof(mine), ncclChar,
                                       idx->comm, stream);
    cudaMemcpyAsync(all.data(), d_handles, sizeof(mine) * n, cudaMemcpyDeviceToHost, stream);
    if (cudaStreamSynchronize(stream) != cudaSuccess || nr != ncclSuccess) ok = false;
    cudaFree(d_handles);
    cudaIpcMemHandle_t zero;
    memset(&zero, 0, sizeof(zero));
    for (int r = 0; r < n && ok; ++r)
        if (memcmp(&all[r], &zero, sizeof(zero)) == 0) ok = false;  // some rank failed
    for (int r = 0; r < n && ok; ++r) {
        if (r == idx->rank) {
            idx->xchg_peer[r] = idx->xchg_mem;
        } else if (cudaIpcOpenMemHandle(&idx->xchg_peer[r], all[r],
                                        cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
            idx->xchg_peer[r] = nullptr;
            ok = false;
        }
    }
    // agree on the outcome: everyone uses the exchange or nobody does
    int *d_flag = nullptr;
    std::vector<int> flags((size_t)n, 0);
    int my = ok ? 1 : 0;
    if (cudaMalloc(&d_flag, sizeof(int) * n) == cudaSuccess) {
        cudaMemcpyAsync(d_flag + idx->rank, &my, sizeof(int), cudaMemcpyHostToDevice, stream);
        nr = nccl().AllGather(d_flag + idx->rank, d_flag, sizeof(int), ncclChar, idx->comm, stream);
        cudaMemcpyAsync(flags.data(), d_flag, sizeof(int) * n, cudaMemcpyDeviceToHost, stream);
        if (cudaStreamSynchronize(stream) != cudaSuccess || nr != ncclSuccess) ok = false;
        cudaFree(d_flag);
        for (int r = 0; r < n; ++r) ok = ok && flags[r] == 1;
    } else {
        ok = false;
    }
    cudaGetLastError();
    idx->xchg_ok = ok;
    idx->xchg_seq = 0;
}

void teardown_peer_exchange(nm_index *idx) {
    for (int r = 0; r < nm::kMaxRanks; ++r) {
        if (idx->xchg_peer[r] && idx->xchg_peer[r] != idx->xchg_mem)
            cudaIpcCloseMemHandle(idx->xchg_peer[r]);
        idx->xchg_peer[r] = nullptr;
    }
    idx->xchg_ok = false;
}

uint32_t pow2_ceil(uint32_t v) {
    uint32_t n = 2;
    while (n < v) n <<= 1;
    return n;
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

}  // namespace


// Packed result block device -> pinned host -> caller buffers; one D2H copy, one sync.
int download_results(nm_index *idx, const Shard &sh, Workspace &ws, const ResultLayout &l,
                     uint32_t nq, uint32_t k, uint64_t *out_rows, float *out_scores,
                     uint32_t *out_counts) {
    CUDA_TRY(cudaMemcpyAsync(ws.h_result, ws.d_result, l.total, cudaMemcpyDeviceToHost, ws.stream));
    CUDA_TRY(cudaStreamSynchronize(ws.stream));
    float ms = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&ms, ws.ev0, ws.ev1));
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
    idx->d2h_bytes += l.total;
    return NM_OK;
}

// Rank-independent routing decision for collective searches (every rank must agree).
bool collective_uses_fused_exchange(const nm_index *idx, uint32_t nq, uint32_t k, int metric) {
    if (!idx->xchg_ok || k > (uint32_t)nm::kMaxFastK) return false;
    const bool long_rows = single_query_stages(idx->dim) < 4;
    const bool would_batch = (idx->batching.load() || single_query_stages(idx->dim) < 2) &&
                             (nq >= kBatchMinQueries || long_rows) &&
                             (metric == NM_EUCLIDEAN || (idx->dim % 8u) == 0);
    return !would_batch;
}

// One fused launch per query: scan + peer-memory exchange + merge, results written in place.
int collective_fused(nm_index *idx, const Shard &sh, Workspace &ws, const float *d_queries,
                     uint32_t nq, uint32_t k, int metric, uint64_t *out_rows, float *out_scores,
                     uint32_t *out_counts, cudaStream_t stream) {
    for (uint32_t q = 0; q < nq; ++q) {
        nm::PeerXchg x = make_xchg(idx, ++idx->xchg_seq);
        if (sh.rows == 0) {
            nm::exchange_empty_shard_kernel<<<1, nm::kRowsPerBlock, 0, stream>>>(
                x, k, ws.d_cand, out_rows + (size_t)q * k, out_scores + (size_t)q * k,
                out_counts + q);
            CUDA_TRY(cudaGetLastError());
            idx->merge_launches++;
        } else {
            int rc = launch_scan(idx, sh, ws, d_queries + (size_t)q * idx->dim, k, metric,
                                 idx->comm_row_base, out_rows + (size_t)q * k,
                                 out_scores + (size_t)q * k, out_counts + q, nullptr, stream, &x);
            if (rc) return rc;
        }
    }
    return NM_OK;
}

// ======================================================================================
// C ABI
// ======================================================================================
extern "C" {

int nm_abi_version(void) { return NM_ABI_VERSION; }

const char *nm_last_error(void) { return g_last_error.c_str(); }

int nm_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int nm_index_create(uint32_t dim, const int *devices, int n_dev, nm_index **out) {
    if (!out) return fail(NM_ERR_INVALID_ARGUMENT, "null out pointer");
    *out = nullptr;
    if (dim == 0) return fail(NM_ERR_EMPTY_VECTOR, "dimension must be >= 1");
    int avail = nm_device_count();
    if (avail == 0)
        return fail(NM_ERR_STORAGE, "no CUDA device visible: the SIMILAR scan has no CPU path");
    std::vector<int> devs;
    if (!devices || n_dev <= 0) {
        int cur = 0;
        CUDA_TRY(cudaGetDevice(&cur));
        devs.push_back(cur);
    } else {
        for (int i = 0; i < n_dev; ++i) {
            if (devices[i] < 0 || devices[i] >= avail)
                return fail(NM_ERR_INVALID_ARGUMENT, "device %d out of range (0..%d)", devices[i],
                            avail - 1);
            devs.push_back(devices[i]);
        }
    }
    std::unique_ptr<nm_index> idx(new nm_index());
    idx->dim = dim;
    idx->pitch = (dim + 3u) & ~3u;
    for (int d : devs) {
        cudaDeviceProp prop;
        CUDA_TRY(cudaGetDeviceProperties(&prop, d));
        if (prop.major < 10)
            return fail(NM_ERR_STORAGE, "device %d is sm_%d%d; this library is built for sm_100a", d,
                        prop.major, prop.minor);
        std::unique_ptr<Shard> sh(new Shard());
        sh->device = d;
        sh->sm_count = prop.multiProcessorCount;
        CUDA_TRY(cudaSetDevice(d));
        CUDA_TRY(cudaStreamCreateWithFlags(&sh->copy_stream, cudaStreamNonBlocking));
        idx->shards.push_back(std::move(sh));
    }
    *out = idx.release();
    return NM_OK;
}

void nm_index_destroy(nm_index *idx) {
    if (!idx) return;
    if (idx->comm && nccl().ok) {
        cudaSetDevice(idx->shards[0]->device);
        cudaDeviceSynchronize();
        teardown_peer_exchange(idx);
        nccl().CommDestroy(idx->comm);
        if (idx->xchg_mem) cudaFree(idx->xchg_mem);
    }
    for (auto &sh : idx->shards) {
        cudaSetDevice(sh->device);
        cudaDeviceSynchronize();  // asynchronous nm_search_device work may still be in flight
        sh->pool.clear();
        sh->stream_ws.clear();
        if (sh->d_rows) cudaFree(sh->d_rows);
        if (sh->d_q8) cudaFree(sh->d_q8);
        if (sh->d_meta) cudaFree(sh->d_meta);
        if (sh->d_q8_flag) cudaFree(sh->d_q8_flag);
        for (int b = 0; b < 2; ++b) {
            if (sh->staging[b]) cudaFreeHost(sh->staging[b]);
            if (sh->staging_done[b]) cudaEventDestroy(sh->staging_done[b]);
        }
        if (sh->copy_stream) cudaStreamDestroy(sh->copy_stream);
    }
    delete idx;
}

int nm_index_clear(nm_index *idx) {
    if (!idx) return fail(NM_ERR_INVALID_ARGUMENT, "null index");
    std::unique_lock<std::shared_mutex> g(idx->mu);
    for (auto &sh : idx->shards) {
        sh->rows = 0;
        sh->row_base = 0;
        sh->tmap_valid = false;
        sh->q8_rows = 0;
        sh->tmap8_valid = false;
    }
    return NM_OK;
}

int nm_index_load(nm_index *idx, const float *rows, uint64_t n) {
    if (!idx) return fail(NM_ERR_INVALID_ARGUMENT, "null index");
    if (n && !rows) return fail(NM_ERR_INVALID_ARGUMENT, "null rows");
    std::unique_lock<std::shared_mutex> g(idx->mu);
    const uint64_t G = idx->shards.size();
    for (uint64_t s = 0; s < G; ++s) {
        Shard &sh = *idx->shards[s];
        uint64_t lo = n * s / G, hi = n * (s + 1) / G;
        CUDA_TRY(cudaSetDevice(sh.device));
        sh.rows = 0;
        int rc = shard_reserve(idx, sh, hi - lo, false);
        if (rc) return rc;
        rc = shard_upload(idx, sh, 0, rows + lo * idx->dim, hi - lo);
        if (rc) return rc;
        sh.rows = hi - lo;
        sh.row_base = lo;
        rc = build_tmap(idx, sh);
        if (rc) return rc;
        sh.q8_rows = 0;
        rc = q8_refresh(idx, sh, 0, sh.rows);
        if (rc) return rc;
    }
    return NM_OK;
}

int nm_index_append(nm_index *idx, const float *rows, uint64_t n) {
    if (!idx) return fail(NM_ERR_INVALID_ARGUMENT, "null index");
    if (n == 0) return NM_OK;
    if (!rows) return fail(NM_ERR_INVALID_ARGUMENT, "null rows");
    std::unique_lock<std::shared_mutex> g(idx->mu);
    Shard &sh = *idx->shards.back();
    CUDA_TRY(cudaSetDevice(sh.device));
    int rc = shard_reserve(idx, sh, sh.rows + n, true);
    if (rc) return rc;
    rc = shard_upload(idx, sh, sh.rows, rows, n);
    if (rc) return rc;
    sh.rows += n;
    rc = build_tmap(idx, sh);
    if (rc) return rc;
    return q8_refresh(idx, sh, sh.rows - n, n);
}

static int locate_row(nm_index *idx, uint64_t row, Shard **out, uint64_t *local) {
    for (auto &sh : idx->shards) {
        if (row >= sh->row_base && row < sh->row_base + sh->rows) {
            *out = sh.get();
            *local = row - sh->row_base;
            return NM_OK;
        }
    }
    return fail(NM_ERR_INVALID_ARGUMENT, "row %llu out of range", (unsigned long long)row);
}

int nm_index_update(nm_index *idx, uint64_t row, const float *vec) {
    if (!idx || !vec) return fail(NM_ERR_INVALID_ARGUMENT, "null argument");
    std::unique_lock<std::shared_mutex> g(idx->mu);
    Shard *sh = nullptr;
    uint64_t local = 0;
    int rc = locate_row(idx, row, &sh, &local);
    if (rc) return rc;
    CUDA_TRY(cudaSetDevice(sh->device));
    rc = shard_upload(idx, *sh, local, vec, 1);
    if (rc) return rc;
    return q8_refresh(idx, *sh, local, 1);
}

int nm_index_swap_remove(nm_index *idx, uint64_t row, uint64_t *moved_from) {
    if (!idx) return fail(NM_ERR_INVALID_ARGUMENT, "null index");
    std::unique_lock<std::shared_mutex> g(idx->mu);
    Shard *sh = nullptr;
    uint64_t local = 0;
    int rc = locate_row(idx, row, &sh, &local);
    if (rc) return rc;
    // the globally last row lives in the last non-empty shard
    Shard *last = nullptr;
    for (auto it = idx->shards.rbegin(); it != idx->shards.rend(); ++it)
        if ((*it)->rows) {
            last = it->get();
            break;
        }
    uint64_t last_global = last->row_base + last->rows - 1;
    if (moved_from) *moved_from = last_global;
    if (last_global != row) {
        const size_t bytes = (size_t)idx->pitch * 4;
        const float *src = last->d_rows + (last->rows - 1) * idx->pitch;
        float *dst = sh->d_rows + local * idx->pitch;
        CUDA_TRY(cudaSetDevice(sh->device));
        if (last->device == sh->device)
            CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, sh->copy_stream));
        else
            CUDA_TRY(cudaMemcpyPeerAsync(dst, sh->device, src, last->device, bytes,
                                         sh->copy_stream));
        CUDA_TRY(cudaStreamSynchronize(sh->copy_stream));
    }
    last->rows -= 1;
    CUDA_TRY(cudaSetDevice(last->device));
    rc = build_tmap(idx, *last);
    if (rc) return rc;
    last->q8_rows = std::min(last->q8_rows, last->rows);
    rc = q8_refresh(idx, *last, last->rows, 0);
    if (rc) return rc;
    if (last_global != row) {
        CUDA_TRY(cudaSetDevice(sh->device));
        rc = q8_refresh(idx, *sh, local, 1);  // the moved row took this slot
    }
    return rc;
}

int nm_index_get_row(nm_index *idx, uint64_t row, float *out_vec) {
    if (!idx || !out_vec) return fail(NM_ERR_INVALID_ARGUMENT, "null argument");
    std::shared_lock<std::shared_mutex> g(idx->mu);
    Shard *sh = nullptr;
    uint64_t local = 0;
    int rc = locate_row(idx, row, &sh, &local);
    if (rc) return rc;
    CUDA_TRY(cudaSetDevice(sh->device));
    CUDA_TRY(cudaMemcpy(out_vec, sh->d_rows + local * idx->pitch, (size_t)idx->dim * 4,
                        cudaMemcpyDeviceToHost));
    return NM_OK;
}

uint64_t nm_index_rows(const nm_index *idx) {
    if (!idx) return 0;
    std::shared_lock<std::shared_mutex> g(idx->mu);
    return idx->total_rows();
}
uint32_t nm_index_dim(const nm_index *idx) { return idx ? idx->dim : 0; }
int nm_index_device_count(const nm_index *idx) { return idx ? (int)idx->shards.size() : 0; }

int nm_index_fill_synthetic(nm_index *idx, uint64_t n, uint64_t seed, uint64_t row_offset) {
    if (!idx) return fail(NM_ERR_INVALID_ARGUMENT, "null index");
    std::unique_lock<std::shared_mutex> g(idx->mu);
    const uint64_t G = idx->shards.size();
    for (uint64_t s = 0; s < G; ++s) {
        Shard &sh = *idx->shards[s];
        uint64_t lo = n * s / G, hi = n * (s + 1) / G;
        CUDA_TRY(cudaSetDevice(sh.device));
        sh.rows = 0;
        int rc = shard_reserve(idx, sh, hi - lo, false);
        if (rc) return rc;
        if (hi > lo) {
            nm::fill_synthetic_kernel<<<sh.sm_count * 8, 256, 0, sh.copy_stream>>>(
                sh.d_rows, hi - lo, idx->dim, idx->pitch, seed, row_offset + lo);
            CUDA_TRY(cudaGetLastError());
            CUDA_TRY(cudaStreamSynchronize(sh.copy_stream));
        }
        sh.rows = hi - lo;
        sh.row_base = lo;
        rc = build_tmap(idx, sh);
        if (rc) return rc;
        sh.q8_rows = 0;
        rc = q8_refresh(idx, sh, 0, sh.rows);
        if (rc) return rc;
    }
    return NM_OK;
}

// --------------------------------------------------------------------------------------
// search
// --------------------------------------------------------------------------------------
static int search_impl(nm_index *idx, const float *queries, uint32_t nq, uint32_t k, int metric,
                       const uint64_t *row_mask, uint64_t *out_rows, float *out_scores,
                       uint32_t *out_counts) {
    int rc = validate_search(idx, queries, nq, k, metric, out_rows, out_scores, out_counts);
    if (rc) return rc;
    std::shared_lock<std::shared_mutex> g(idx->mu);
    const size_t G = idx->shards.size();
    const uint32_t dim = idx->dim;
    const bool collective = idx->comm != nullptr;
    if (collective && G != 1)
        return fail(NM_ERR_CONFIGURATION, "a communicator needs a single-device index per rank");
    if (row_mask && (collective || G != 1))
        return fail(NM_ERR_CONFIGURATION,
                    "nm_search_masked needs a single-device index without a communicator");

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
        memcpy(ws->h_query, queries, (size_t)nq * dim * 4);
        CUDA_TRY(cudaMemcpyAsync(ws->d_query, ws->h_query, (size_t)nq * dim * 4,
                                 cudaMemcpyHostToDevice, ws->stream));
        uint64_t *r_rows = reinterpret_cast<uint64_t *>(ws->d_result + l.rows_off);
        float *r_scores = reinterpret_cast<float *>(ws->d_result + l.scores_off);
        uint32_t *r_counts = reinterpret_cast<uint32_t *>(ws->d_result + l.counts_off);
        if (row_mask) {
            // stage the bitmask (bit r of word r/64 == bit r%32 of u32 word r/32 on little
            // endian hosts), padded with zeros to whole row blocks
            const size_t words = (((size_t)sh.rows + 255) / 256) * 8;
            const size_t src_bytes = (((size_t)sh.rows + 63) / 64) * 8;
            if (ws->mask_cap < words) {
                if (ws->d_mask) CUDA_TRY(cudaFree(ws->d_mask));
                ws->mask_cap = 0;
                CUDA_TRY(cudaMalloc(&ws->d_mask, words * 4));
                ws->mask_cap = words;
            }
            CUDA_TRY(cudaMemsetAsync(ws->d_mask, 0, words * 4, ws->stream));
            CUDA_TRY(cudaMemcpyAsync(ws->d_mask, row_mask, src_bytes, cudaMemcpyHostToDevice,
                                     ws->stream));
            idx->h2d_bytes += src_bytes;
        }
        CUDA_TRY(cudaEventRecord(ws->ev0, ws->stream));
        if (row_mask) {
            for (uint32_t q = 0; q < nq; ++q) {
                rc = launch_scan(idx, sh, *ws, ws->d_query + (size_t)q * dim, k, metric, sh.row_base,
                                 r_rows + (size_t)q * k, r_scores + (size_t)q * k, r_counts + q,
                                 nullptr, ws->stream, nullptr, ws->d_mask);
                if (rc) return rc;
            }
        } else if (prefilter_usable(idx, sh, nq, k, metric, row_mask)) {
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
            CUDA_TRY(cudaEventRecord(ws->ev1, ws->stream));
            CUDA_TRY(cudaMemcpyAsync(ws->h_result, ws->d_result, l.total, cudaMemcpyDeviceToHost,
                                     ws->stream));
            CUDA_TRY(cudaStreamSynchronize(ws->stream));
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
                CUDA_TRY(cudaEventElapsedTime(&ms, ws->ev0, ws->ev1));
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
        CUDA_TRY(cudaEventRecord(ws->ev1, ws->stream));
        return download_results(idx, sh, *ws, l, nq, k, out_rows, out_scores, out_counts);
    }

    // ---- collective path: one shard per process.  Single queries: ONE fused launch (scan +
    //      peer-memory exchange + merge).  Batches / k > 1024: scan, ONE ncclAllGather of the
    //      per-shard hits, merge kernel. ----
    if (collective) {
        Shard &sh = *idx->shards[0];
        CUDA_TRY(cudaSetDevice(sh.device));
        std::lock_guard<std::mutex> cg(idx->comm_mu);
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
        uint64_t *r_rows = reinterpret_cast<uint64_t *>(ws->d_result + l.rows_off);
        float *r_scores = reinterpret_cast<float *>(ws->d_result + l.scores_off);
        uint32_t *r_counts = reinterpret_cast<uint32_t *>(ws->d_result + l.counts_off);
        memcpy(ws->h_query, queries, (size_t)nq * dim * 4);
        CUDA_TRY(cudaMemcpyAsync(ws->d_query, ws->h_query, (size_t)nq * dim * 4,
                                 cudaMemcpyHostToDevice, ws->stream));
        CUDA_TRY(cudaEventRecord(ws->ev0, ws->stream));
        if (collective_uses_fused_exchange(idx, nq, k, metric)) {
            rc = collective_fused(idx, sh, *ws, ws->d_query, nq, k, metric, r_rows, r_scores,
                                  r_counts, ws->stream);
            if (rc) return rc;
            CUDA_TRY(cudaEventRecord(ws->ev1, ws->stream));
        } else {
            if (sh.rows == 0) {
                CUDA_TRY(cudaMemsetAsync(ws->d_hits, 0, (size_t)nq * k * sizeof(nm::ShardHit),
                                         ws->stream));
            } else {
                rc = scan_queries(idx, sh, *ws, ws->d_query, nq, k, metric, idx->comm_row_base,
                                  nullptr, nullptr, nullptr, ws->d_hits, ws->stream);
                if (rc) return rc;
            }
            CUDA_TRY(cudaEventRecord(ws->ev1, ws->stream));
            NCCL_TRY(nccl().AllGather(ws->d_hits, ws->d_gather,
                                      (size_t)nq * k * sizeof(nm::ShardHit), ncclChar, idx->comm,
                                      ws->stream));
            uint32_t total = (uint32_t)idx->n_ranks * k;
            uint32_t n_sort = pow2_ceil(total);
            size_t msmem = (size_t)n_sort * 8;
            if (msmem > 200 * 1024)
                return fail(NM_ERR_INVALID_TOP_K, "n_ranks*k = %u too large for the merge kernel",
                            total);
            if (msmem > 48 * 1024)
                CUDA_TRY(cudaFuncSetAttribute(nm::merge_shards_kernel,
                                              cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              200 * 1024));
            nm::merge_shards_kernel<<<nq, nm::kMergeThreads, msmem, ws->stream>>>(
                ws->d_gather, (uint32_t)idx->n_ranks, k, nq * k, n_sort, r_rows, r_scores, r_counts);
            CUDA_TRY(cudaGetLastError());
            idx->merge_launches++;
        }
        return download_results(idx, sh, *ws, l, nq, k, out_rows, out_scores, out_counts);
    }

    // ---- several devices in this process: scan each shard, merge the per-shard top-k on
    //      the host (G*k hits), exactly ResultMerger::merge_top_k -------------------------
    std::vector<std::unique_ptr<Workspace>> wss(G);
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
        CUDA_TRY(cudaEventRecord(ws.ev0, ws.stream));
        if (sh.rows == 0) {
            CUDA_TRY(cudaMemsetAsync(ws.d_hits, 0, (size_t)nq * k * sizeof(nm::ShardHit), ws.stream));
        } else {
            rc = scan_queries(idx, sh, ws, ws.d_query, nq, k, metric, sh.row_base, nullptr, nullptr,
                              nullptr, ws.d_hits, ws.stream);
            if (rc) return rc;
        }
        CUDA_TRY(cudaEventRecord(ws.ev1, ws.stream));
        CUDA_TRY(cudaMemcpyAsync(ws.h_hits, ws.d_hits, (size_t)nq * k * sizeof(nm::ShardHit),
                                 cudaMemcpyDeviceToHost, ws.stream));
    }
    float max_ms = 0.f;
    for (size_t s = 0; s < G; ++s) {
        CUDA_TRY(cudaSetDevice(idx->shards[s]->device));
        CUDA_TRY(cudaStreamSynchronize(wss[s]->stream));
        float ms = 0.f;
        CUDA_TRY(cudaEventElapsedTime(&ms, wss[s]->ev0, wss[s]->ev1));
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

int nm_search(nm_index *idx, const float *queries, uint32_t nq, uint32_t k, int metric,
              uint64_t *out_rows, float *out_scores, uint32_t *out_counts) {
    return search_impl(idx, queries, nq, k, metric, nullptr, out_rows, out_scores, out_counts);
}

int nm_search_masked(nm_index *idx, const float *queries, uint32_t nq, uint32_t k, int metric,
                     const uint64_t *row_mask, uint64_t *out_rows, float *out_scores,
                     uint32_t *out_counts) {
    if (!row_mask) return fail(NM_ERR_INVALID_ARGUMENT, "null row mask");
    return search_impl(idx, queries, nq, k, metric, row_mask, out_rows, out_scores, out_counts);
}

int nm_search_device(nm_index *idx, const float *d_queries, uint32_t nq, uint32_t k, int metric,
                     uint64_t *d_out_rows, float *d_out_scores, uint32_t *d_out_counts,
                     void *stream_v) {
    int rc = validate_search(idx, d_queries, nq, k, metric, d_out_rows, d_out_scores, d_out_counts);
    if (rc) return rc;
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
            CUDA_TRY(cudaMalloc(&nw->d_counter, 2 * sizeof(uint32_t)));
            CUDA_TRY(cudaMemsetAsync(nw->d_counter, 0, 2 * sizeof(uint32_t), stream));
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
        } else {
            rc = scan_queries(idx, sh, *ws, d_queries, nq, k, metric, sh.row_base, d_out_rows,
                              d_out_scores, d_out_counts, nullptr, stream);
            if (rc) return rc;
        }
        if (prof) CUDA_TRY(cudaEventRecord(prof->second, stream));
    } else if (collective_uses_fused_exchange(idx, nq, k, metric)) {
        // (collective calls on one index must be issued in the same order on every rank)
        std::lock_guard<std::mutex> cg(idx->comm_mu);
        rc = collective_fused(idx, sh, *ws, d_queries, nq, k, metric, d_out_rows, d_out_scores,
                              d_out_counts, stream);
        if (rc) return rc;
        if (prof) CUDA_TRY(cudaEventRecord(prof->second, stream));
    } else {
        std::lock_guard<std::mutex> cg(idx->comm_mu);
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
        uint32_t total = (uint32_t)idx->n_ranks * k;
        uint32_t n_sort = pow2_ceil(total);
        size_t msmem = (size_t)n_sort * 8;
        if (msmem > 200 * 1024)
            return fail(NM_ERR_INVALID_TOP_K, "n_ranks*k = %u too large for the merge kernel", total);
        if (msmem > 48 * 1024)
            CUDA_TRY(cudaFuncSetAttribute(nm::merge_shards_kernel,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        nm::merge_shards_kernel<<<nq, nm::kMergeThreads, msmem, stream>>>(
            ws->d_gather, (uint32_t)idx->n_ranks, k, nq * k, n_sort, d_out_rows, d_out_scores,
            d_out_counts);
        CUDA_TRY(cudaGetLastError());
        idx->merge_launches++;
    }
    if (!stream_v) CUDA_TRY(cudaStreamSynchronize(stream));
    idx->searches += nq;
    idx->rows_scanned += (uint64_t)nq * sh.rows;
    idx->bytes_streamed += (uint64_t)nq * sh.rows * dim * 4;
    return NM_OK;
}

int nm_index_set_prefilter(nm_index *idx, int mode) {
    if (!idx) return fail(NM_ERR_INVALID_ARGUMENT, "null index");
    if (mode != 0 && mode != 1) return fail(NM_ERR_INVALID_ARGUMENT, "unknown pre-filter mode %d", mode);
    std::unique_lock<std::shared_mutex> g(idx->mu);
    idx->prefilter = mode;
    for (auto &shp : idx->shards) {
        Shard &sh = *shp;
        CUDA_TRY(cudaSetDevice(sh.device));
        if (mode == 0) {
            if (sh.d_q8) CUDA_TRY(cudaFree(sh.d_q8));
            if (sh.d_meta) CUDA_TRY(cudaFree(sh.d_meta));
            sh.d_q8 = nullptr;
            sh.d_meta = nullptr;
            sh.q8_capacity = sh.q8_rows = 0;
            sh.tmap8_valid = false;
        } else {
            sh.q8_rows = 0;
            int rc = q8_refresh(idx, sh, 0, sh.rows);
            if (rc) return rc;
        }
    }
    return NM_OK;
}

int nm_index_set_batching(nm_index *idx, int enable) {
    if (!idx) return fail(NM_ERR_INVALID_ARGUMENT, "null index");
    idx->batching = enable ? 1 : 0;
    return NM_OK;
}

int nm_index_set_profiling(nm_index *idx, int enable) {
    if (!idx) return fail(NM_ERR_INVALID_ARGUMENT, "null index");
    idx->profiling = enable ? 1 : 0;
    return NM_OK;
}

// --------------------------------------------------------------------------------------
// communicator
// --------------------------------------------------------------------------------------
int nm_comm_create_id(void *out_id) {
    if (!out_id) return fail(NM_ERR_INVALID_ARGUMENT, "null id buffer");
    if (!nccl().ok) return fail(NM_ERR_STORAGE, "libnccl.so.2 could not be loaded");
    static_assert(sizeof(ncclUniqueId) == NM_COMM_ID_BYTES, "ncclUniqueId size");
    ncclUniqueId id;
    NCCL_TRY(nccl().GetUniqueId(&id));
    memcpy(out_id, &id, sizeof(id));
    return NM_OK;
}

int nm_index_attach_comm(nm_index *idx, const void *id, int n_ranks, int rank, uint64_t row_base) {
    if (!idx || !id) return fail(NM_ERR_INVALID_ARGUMENT, "null argument");
    if (n_ranks < 1 || rank < 0 || rank >= n_ranks)
        return fail(NM_ERR_INVALID_ARGUMENT, "bad rank %d of %d", rank, n_ranks);
    if (!nccl().ok) return fail(NM_ERR_STORAGE, "libnccl.so.2 could not be loaded");
    std::unique_lock<std::shared_mutex> g(idx->mu);
    if (idx->shards.size() != 1)
        return fail(NM_ERR_CONFIGURATION, "a communicator needs a single-device index per rank");
    if (idx->comm) return fail(NM_ERR_CONFIGURATION, "communicator already attached");
    CUDA_TRY(cudaSetDevice(idx->shards[0]->device));
    ncclUniqueId uid;
    memcpy(&uid, id, sizeof(uid));
    NCCL_TRY(nccl().CommInitRank(&idx->comm, n_ranks, uid, rank));
    idx->n_ranks = n_ranks;
    idx->rank = rank;
    idx->comm_row_base = row_base;
    setup_peer_exchange(idx, idx->shards[0]->copy_stream);
    return NM_OK;
}

int nm_index_detach_comm(nm_index *idx) {
    if (!idx) return fail(NM_ERR_INVALID_ARGUMENT, "null index");
    std::unique_lock<std::shared_mutex> g(idx->mu);
    if (idx->comm) {
        CUDA_TRY(cudaSetDevice(idx->shards[0]->device));
        CUDA_TRY(cudaDeviceSynchronize());
        teardown_peer_exchange(idx);
        NCCL_TRY(nccl().CommDestroy(idx->comm));  // collective: every rank has unmapped by now
        idx->comm = nullptr;
        if (idx->xchg_mem) cudaFree(idx->xchg_mem);
        idx->xchg_mem = nullptr;
    }
    idx->n_ranks = 1;
    idx->rank = 0;
    idx->comm_row_base = 0;
    return NM_OK;
}

int nm_index_stats(nm_index *idx, nm_stats *out) {
    if (!idx || !out) return fail(NM_ERR_INVALID_ARGUMENT, "null argument");
    out->searches = idx->searches;
    out->rows_scanned = idx->rows_scanned;
    out->bytes_streamed = idx->bytes_streamed;
    out->scan_launches = idx->scan_launches;
    out->merge_launches = idx->merge_launches;
    out->h2d_bytes = idx->h2d_bytes;
    out->d2h_bytes = idx->d2h_bytes;
    out->last_scan_ms = idx->last_scan_ms;
    out->prefilter_queries = idx->pf_queries;
    out->prefilter_fallbacks = idx->pf_fallbacks;
    out->prefilter_kept = idx->pf_kept;
    {
        // fold finished profiling event pairs into the totals (waits for the streams)
        std::unique_lock<std::shared_mutex> g(idx->mu);
        for (auto &sh : idx->shards) {
            cudaSetDevice(sh->device);
            std::lock_guard<std::mutex> pg(sh->pool_mu);
            std::vector<Workspace *> all_ws;
            for (auto &e : sh->stream_ws) all_ws.push_back(e.second.get());
            for (auto &w : sh->pool) all_ws.push_back(w.get());
            for (Workspace *wsp : all_ws) {
                Workspace &ws = *wsp;
                for (size_t i = 0; i < ws.prof_used; ++i) {
                    float ms = 0.f;
                    if (cudaEventSynchronize(ws.prof_events[i].second) == cudaSuccess &&
                        cudaEventElapsedTime(&ms, ws.prof_events[i].first,
                                             ws.prof_events[i].second) == cudaSuccess) {
                        idx->profiled_scan_ms += ms;
                        idx->profiled_scans += 1;
                    }
                }
                ws.prof_used = 0;
            }
        }
        cudaGetLastError();
        out->profiled_scan_ms = idx->profiled_scan_ms;
        out->profiled_scans = idx->profiled_scans;
    }
    return NM_OK;
}

}  // extern "C"
