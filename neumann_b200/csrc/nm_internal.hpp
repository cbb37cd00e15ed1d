// nm_internal.hpp — shared declarations of the host-side translation units:
//   nm_core.cu    mirror lifecycle, staging, mutations, int8 copy upkeep, statistics
//   nm_launch.cu  every kernel launch (the only unit that includes the kernel headers)
//   nm_search.cu  nm_search / nm_search_masked / nm_search_device and their routing
//   nm_comm.cu    NCCL loader, communicator, CUDA-IPC peer exchange
#pragma once
#include "../../include/neumann_b200.h"
#include "nm_types.hpp"
#include "nm_vmm.hpp"

#include <cuda.h>
#include <cuda_runtime.h>
#include <nccl.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <shared_mutex>
#include <string>
#include <vector>

namespace nmi {

// thread-local last error + nm_status in one call
int fail(int code, const char *fmt, ...);
const char *last_error();

#define CUDA_TRY(expr)                                                                          \
    do {                                                                                        \
        cudaError_t _e = (expr);                                                                \
        if (_e != cudaSuccess)                                                                  \
            return nmi::fail(NM_ERR_STORAGE, "CUDA error %s at %s:%d: %s", cudaGetErrorName(_e), \
                             __FILE__, __LINE__, cudaGetErrorString(_e));                       \
    } while (0)

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

NcclApi &nccl();

#define NCCL_TRY(expr)                                                                       \
    do {                                                                                     \
        ncclResult_t _r = (expr);                                                            \
        if (_r != ncclSuccess)                                                               \
            return nmi::fail(NM_ERR_STORAGE, "NCCL error at %s:%d: %s", __FILE__, __LINE__,  \
                             nmi::nccl().GetErrorString(_r));                                \
    } while (0)

// Per-call scratch on one device.  Pooled per shard so concurrent nm_search calls never share
// a stream, a candidate buffer or the "last CTA" ticket.
struct Workspace {
    int device = -1;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    bool timed = false;  // ev0 / ev1 were recorded for the call in flight (profiling on)
    float *d_query = nullptr;
    float *h_query = nullptr;  // pinned
    size_t query_cap = 0;      // floats
    uint64_t *d_cand = nullptr;
    size_t cand_cap = 0;  // keys
    // [0..1] ticket + row-block cursor, [2..3] the same for odd pipelined scans, [4] highest
    // finished pipelined sequence number (kWsCounterWords words)
    uint32_t *d_counter = nullptr;
    uint32_t pipe_seq = 0;          // pipelined scans issued on this (stream-bound) workspace
    bool pipeline_next = false;     // set by nm_search_device: the next launch_scan may overlap
    cudaEvent_t async_done = nullptr;  // recorded after every non-pipelined asynchronous call
    bool async_pending = false;
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
    // tensor-core batch pre-filter (tc_prefilter_kernels.cuh)
    int8_t *d_tc_q8 = nullptr;      // [256][pitch8] int8 queries of the current pass
    size_t tc_q8_cap = 0;           // bytes
    void *d_tc_qmeta = nullptr;     // [nq] nm::TcQueryMeta
    void *h_tc_qmeta = nullptr;     // pinned
    void *d_tc_coef = nullptr;      // [nq] float4
    uint32_t *d_tc_kept_n = nullptr;  // [nq] + [1] statistics
    size_t tc_nq_cap = 0;           // queries
    void *d_tc_kept = nullptr;      // [256][kTcKeptCap] nm::TcKept
    uint64_t *d_tc_keys = nullptr;  // [256][kTcKeptCap]
    uint8_t *d_tc_out = nullptr;      // scratch result block of a sharded tensor-core pass
    size_t tc_out_cap = 0;
    uint32_t *d_tc_redo = nullptr;    // [tc_nq_cap] device-side redo flags (asynchronous searches)
    uint32_t *d_tc_bucket = nullptr;  // row-bucket counters / cursors of the re-score ordering
    void *d_tc_sorted = nullptr;      // [256 * kTcKeptCap] uint2 {row, query << 16 | slot}
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
        if (d_tc_q8) cudaFree(d_tc_q8);
        if (d_tc_qmeta) cudaFree(d_tc_qmeta);
        if (h_tc_qmeta) cudaFreeHost(h_tc_qmeta);
        if (d_tc_coef) cudaFree(d_tc_coef);
        if (d_tc_kept_n) cudaFree(d_tc_kept_n);
        if (d_tc_kept) cudaFree(d_tc_kept);
        if (d_tc_keys) cudaFree(d_tc_keys);
        if (d_tc_out) cudaFree(d_tc_out);
        if (d_tc_redo) cudaFree(d_tc_redo);
        if (d_tc_bucket) cudaFree(d_tc_bucket);
        if (d_tc_sorted) cudaFree(d_tc_sorted);
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
        if (async_done) cudaEventDestroy(async_done);
        for (auto &pe : prof_events) {
            cudaEventDestroy(pe.first);
            cudaEventDestroy(pe.second);
        }
        if (stream) cudaStreamDestroy(stream);
    }
};

// One metadata field of the rows of a shard (filter_kernels.cuh).  Both arrays cover the
// shard's rows [0, init_rows) (zero = missing beyond what was ever set) and grow in place.
struct Column {
    GrowBuf tags_buf, vals_buf;
    uint8_t *d_tags = nullptr;
    uint64_t *d_vals = nullptr;
    uint64_t init_rows = 0;
};

// A computed row mask, shared between the cache and the searches using it.
struct MaskEntry {
    int device = 0;
    std::vector<uint8_t> key;   // program bytes + table words
    uint64_t epoch = 0;         // nm_index::mutation_epoch it was computed at
    uint32_t *d_mask = nullptr; // [words] u32, padded to whole row blocks
    size_t words = 0;
    size_t words_cap = 0;       // allocation sizes: entries come from a per-shard pool filled when
    size_t prog_cap = 0;        // metadata arrives and go back to it (no cudaMalloc while searching)
    void *d_prog = nullptr;     // FilterOpDev[] followed by the string tables
    cudaEvent_t ready = nullptr;
    ~MaskEntry() {
        cudaSetDevice(device);
        if (d_mask) cudaFree(d_mask);
        if (d_prog) cudaFree(d_prog);
        if (ready) cudaEventDestroy(ready);
    }
};

// What restricts a search to a subset of the rows: a host bitmask or a filter program.
struct MaskSpec {
    const uint64_t *host_mask = nullptr;
    const nm_filter_op *prog = nullptr;
    uint32_t n_ops = 0;
    const uint32_t *tables = nullptr;
    uint32_t n_table_words = 0;
    bool any() const { return host_mask || prog; }
};

struct Shard {
    int device = 0;
    int sm_count = 0;
    // The mirror and its int8 copy live in buffers that grow in place (nm_vmm.hpp): appends map
    // more physical chunks behind the existing rows, nothing is ever copied to grow.
    GrowBuf rows_buf, q8_buf, meta_buf, norms_buf;
    float *d_rows = nullptr;  // == rows_buf.ptr()
    uint64_t rows = 0;      // local rows
    uint64_t capacity = 0;  // local rows the mapped part of rows_buf holds
    uint64_t row_base = 0;  // global index of local row 0 (within this process)
    CUtensorMap tmap;
    bool tmap_valid = false;
    // int8 pre-filter copy (prefilter_kernels.cuh): [capacity8, pitch8] bytes + 16 B per row
    int8_t *d_q8 = nullptr;
    nm::RowMeta *d_meta = nullptr;
    float2 *d_norms = nullptr;  // [capacity8] per-row L2 norms for the batch pre-filter's bound
    uint32_t *d_q8_flag = nullptr;
    uint64_t q8_capacity = 0;
    uint64_t q8_rows = 0;  // rows [0, q8_rows) are quantised and current
    CUtensorMap tmap8;
    CUtensorMap tmap8_tc;  // same copy, [128 rows x 128 B] boxes (tensor-core batch pre-filter)
    bool tmap8_valid = false;
    // metadata columns (nm_index_column_set): typed per-row values next to the mirror, and the
    // row masks of recently used filters (valid until the next mutation)
    std::map<uint32_t, std::unique_ptr<Column>> columns;
    std::mutex mask_mu;
    std::vector<std::shared_ptr<MaskEntry>> mask_cache;
    std::vector<std::shared_ptr<MaskEntry>> mask_free;  // buffers of stale masks, for reuse
    cudaStream_t copy_stream = nullptr;
    float *staging[2] = {nullptr, nullptr};
    cudaEvent_t staging_done[2] = {nullptr, nullptr};
    std::mutex pool_mu;
    std::vector<std::unique_ptr<Workspace>> pool;
    // Workspaces bound to a caller stream (nm_search_device): work on one stream is ordered,
    // so the same scratch can be reused by consecutive asynchronous calls without a sync.
    std::vector<std::pair<cudaStream_t, std::unique_ptr<Workspace>>> stream_ws;
};

}  // namespace nmi

struct nm_index {
    uint32_t dim = 0;
    uint32_t pitch = 0;  // floats per row in device memory
    std::vector<std::unique_ptr<nmi::Shard>> shards;
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
    // nm_index_set_prefilter: 0 = never, 1 = int8 copy always (single queries through the dp4a
    // pre-filter, batches through the tensor-core pre-filter), 2 = auto (default): the copy is
    // built at the first eligible batch when it fits in free HBM and serves batches only
    std::atomic<int> prefilter{2};
    std::atomic<uint64_t> q8_auto_declined_rows{~0ull};  // auto build refused at this row count
    std::atomic<int> tensor_core{1};  // nm_index_set_tensor_core: batches of a pre-filtered index
                                      // go through the tcgen05 int8 GEMM pre-filter
    std::atomic<uint64_t> tc_queries{0}, tc_fallbacks{0}, tc_survivors{0};
    std::atomic<uint64_t> pf_queries{0}, pf_fallbacks{0}, pf_kept{0};
    std::atomic<int> batching{1};  // nm_index_set_batching: 0 forces one scan per query
    std::atomic<int> pipelining{0};  // nm_index_set_pipelining: async single-query scans overlap
    cudaStream_t xchg_async_stream = nullptr;  // last caller stream of an async collective search
    // Coalescing of concurrent single-query nm_search calls (nm_index_set_coalescing): while
    // one batch is on the GPU, calls from other host threads queue up and ride the next corpus
    // pass together (batched kernels).  No timers: an idle index serves a lone call at once.
    std::atomic<int> coalesce_max{64};
    struct PendingSearch {
        const float *query;
        uint32_t k;
        int metric;
        uint64_t *out_rows;
        float *out_scores;
        uint32_t *out_count;
        int rc = 0;
        std::string error;
        bool done = false;
    };
    std::mutex co_mu;
    std::condition_variable co_cv;
    std::vector<PendingSearch *> co_pending;
    bool co_leader = false;
    std::atomic<uint64_t> co_batches{0}, co_queries{0};
    std::atomic<uint64_t> filter_masks_built{0}, filter_mask_hits{0};
    std::atomic<uint64_t> mutation_epoch{0};  // bumped by everything that invalidates row masks
    double profiled_scan_ms = 0.0;  // guarded by mu (exclusive) in nm_index_stats
    uint64_t profiled_scans = 0;
    uint64_t total_rows() const {
        uint64_t n = 0;
        for (auto &s : shards) n += s->rows;
        return n;
    }
};

namespace nmi {

constexpr uint32_t kWsCounterWords = 8;
constexpr uint32_t kBatchMinQueries = 2;  // from 2 queries on, sharing the corpus pass pays

struct ResultLayout {
    size_t counts_off, rows_off, scores_off, total;
};
inline ResultLayout result_layout(uint32_t nq, uint32_t k) {
    ResultLayout l;
    l.counts_off = 0;
    l.rows_off = ((size_t)nq * 4 + 15) & ~size_t(15);
    l.scores_off = l.rows_off + (size_t)nq * k * 8;
    l.total = l.scores_off + (size_t)nq * k * 4;
    return l;
}
inline uint32_t pow2_ceil(uint32_t v) {
    uint32_t n = 2;
    while (n < v) n <<= 1;
    return n;
}

// ---- nm_core.cu ----
int build_tmap(nm_index *idx, Shard &sh);
int q8_refresh(nm_index *idx, Shard &sh, uint64_t first, uint64_t n, bool build = false);
// auto mode: build the int8 copy now if a batch of nq queries would use it and it fits
int q8_auto_prepare(nm_index *idx, uint32_t nq, uint32_t k);
bool tc_shape_ok(const nm_index *idx, uint64_t shard_rows, uint32_t nq, uint32_t k);
int encode_tmap_u8(CUtensorMap *out, void *base, uint64_t inner, uint64_t rows, uint64_t pitch_bytes,
                   uint32_t box_inner, uint32_t box_rows);
uint32_t q8_pitch(uint32_t dim);

// ---- nm_launch.cu: workspaces + every kernel launch ----
int ws_acquire(Shard &sh, std::unique_ptr<Workspace> &out);
// wait until every asynchronous nm_search_device call issued so far on this index has finished
// (mutations call it under the write lock, so they never tear an in-flight scan)
int wait_async_searches(nm_index *idx);
void ws_release(Shard &sh, std::unique_ptr<Workspace> &ws);
int ws_ensure(Workspace &ws, const Shard &sh, uint32_t dim, uint32_t nq, uint32_t k, bool need_query,
              bool need_result, bool need_hits, int gather_ranks);
int ws_ensure_prefilter(Workspace &ws, uint32_t nq);
uint32_t single_query_stages(uint32_t dim);
bool batch_eligible(const nm_index *idx, const Shard &sh, uint32_t nq, uint32_t k, int metric);
bool prefilter_usable(const nm_index *idx, const Shard &sh, uint32_t nq, uint32_t k, int metric,
                      bool masked);
int launch_scan(nm_index *idx, const Shard &sh, Workspace &ws, const float *d_query, uint32_t k,
                int metric, uint64_t row_base, uint64_t *out_rows, float *out_scores,
                uint32_t *out_count, nm::ShardHit *out_hits, cudaStream_t stream,
                const nm::PeerXchg *xchg = nullptr, const uint32_t *d_row_mask = nullptr,
                const uint32_t *d_gate = nullptr);
// d_gate (device, [nq], may be null): conditional execution — a query (or the pass of the batched
// kernels it belongs to) is only computed if its gate word is non-zero
int scan_queries(nm_index *idx, const Shard &sh, Workspace &ws, const float *d_queries, uint32_t nq,
                 uint32_t k, int metric, uint64_t row_base, uint64_t *out_rows, float *out_scores,
                 uint32_t *out_counts, nm::ShardHit *out_hits, cudaStream_t stream,
                 const uint32_t *d_gate = nullptr);
// after scan_queries_tc on the same stream: per-query redo flags, computed on the device
int tc_redo_flags(const Workspace &ws, uint32_t nq, uint32_t rows, uint32_t **d_redo, cudaStream_t stream);
int launch_prefiltered(nm_index *idx, const Shard &sh, Workspace &ws, const float *d_query,
                       uint32_t q, uint32_t k, int metric, uint64_t *out_rows, float *out_scores,
                       uint32_t *out_count, cudaStream_t stream);
bool tc_usable(const nm_index *idx, const Shard &sh, uint32_t nq, uint32_t k, int metric,
               bool masked);
// nq queries through the tensor-core pre-filter; h_flags[q] != 0 afterwards (once the stream has
// been waited for) means query q must be redone by the exact path.  *h_flags_out points into
// pinned memory owned by the workspace.
int scan_queries_tc(nm_index *idx, const Shard &sh, Workspace &ws, const float *d_queries,
                    uint32_t nq, uint32_t k, int metric, uint64_t row_base, uint64_t *out_rows,
                    float *out_scores, uint32_t *out_counts, cudaStream_t stream,
                    int *debug_dots = nullptr, const uint32_t *d_row_mask = nullptr);
int scan_queries_tc_hits_enqueue(nm_index *idx, const Shard &sh, Workspace &ws,
                                 const float *d_queries, uint32_t nq, uint32_t k, int metric,
                                 uint64_t row_base, nm::ShardHit *out_hits, cudaStream_t stream,
                                 const uint32_t *d_row_mask = nullptr);
int scan_queries_tc_hits_finish(nm_index *idx, const Shard &sh, Workspace &ws,
                                const float *d_queries, uint32_t nq, uint32_t k, int metric,
                                uint64_t row_base, nm::ShardHit *out_hits, cudaStream_t stream,
                                const uint32_t *d_row_mask = nullptr);
uint32_t tc_query_flags(const Workspace &ws, uint32_t q, uint32_t rows);
uint32_t tc_phases(const Workspace &ws);
uint32_t tc_survivors(const Workspace &ws);
int launch_merge_shards(nm_index *idx, const nm::ShardHit *d_gather, uint32_t nq, uint32_t k,
                        uint64_t *out_rows, float *out_scores, uint32_t *out_counts,
                        cudaStream_t stream);
int launch_exchange_empty(const nm::PeerXchg &x, uint32_t k, uint64_t *scratch, uint64_t *out_rows,
                          float *out_scores, uint32_t *out_count, cudaStream_t stream);
int launch_fill_synthetic(const Shard &sh, float *rows, uint64_t n, uint32_t dim, uint32_t pitch,
                          uint64_t seed, uint64_t global_row0, cudaStream_t stream);
int launch_quantize(const Shard &sh, const float *rows, uint32_t pitch, uint32_t dim, uint64_t first,
                    uint64_t n, int8_t *q8, uint32_t pitch8, nm::RowMeta *meta, float2 *norms,
                    uint32_t *flag, cudaStream_t stream);

int launch_filter_mask(const Shard &sh, const nm::FilterOpDev *d_ops, uint32_t n_ops, uint32_t max_depth,
                       uint64_t n_rows,
                       uint32_t *d_mask, uint64_t n_words, cudaStream_t stream);
int launch_column_move(uint8_t *tags, uint64_t *vals, uint64_t dst, uint64_t src, cudaStream_t stream);

// ---- nm_columns.cu: metadata columns, filter -> row mask ----
// (all but shard_mask are called with the index write lock held)
int columns_after_resize(nm_index *idx, Shard &sh);  // rows appended: new rows read as missing
void columns_drop(Shard &sh);                        // load / clear: rows replaced
int columns_swap_remove(nm_index *idx, Shard &dst, uint64_t dst_local, Shard &src, uint64_t src_local);
struct ColumnsSnapshot;                               // all columns of all shards, on the host
int columns_gather(nm_index *idx, std::shared_ptr<ColumnsSnapshot> *out);
int columns_scatter(nm_index *idx, const ColumnsSnapshot &snap);
int validate_filter_program(const nm_filter_op *prog, uint32_t n_ops, const uint32_t *tables,
                            uint32_t n_table_words);
// The device mask of `spec` for this shard on `stream` (shared lock held).  `first_row` = global
// index (within spec.host_mask) of the shard's row 0.  *hold keeps a cached mask alive.
int shard_mask(nm_index *idx, Shard &sh, Workspace &ws, const MaskSpec &spec, uint64_t first_row,
               cudaStream_t stream, const uint32_t **d_mask, std::shared_ptr<MaskEntry> *hold);

// ---- nm_comm.cu ----
nm::PeerXchg make_xchg(const nm_index *idx, uint32_t seq);
void setup_peer_exchange(nm_index *idx, cudaStream_t stream);
void teardown_peer_exchange(nm_index *idx);

}  // namespace nmi
