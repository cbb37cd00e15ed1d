// nm_columns.cu — the columnar metadata mirror and the filter -> row-mask step of filtered
// searches (kernel: filter_kernels.cuh).  Replaces the host loop that walked every row's
// metadata map per filtered query (reference: search_with_pre_filter,
// vector_engine/src/lib.rs:3514-3557 collects the matching keys first; evaluate_filter :3592).
#include "nm_internal.hpp"
#include "nm_trace.hpp"

using namespace nmi;

namespace nmi {

namespace {

int column_reserve(Shard &sh, Column &c, uint64_t rows) {
    if (rows > c.init_rows) {
        int rc = c.tags_buf.ensure(sh.device, rows, false);
        if (!rc) rc = c.vals_buf.ensure(sh.device, rows * sizeof(uint64_t), false);
        if (rc) return rc;
        c.d_tags = static_cast<uint8_t *>(c.tags_buf.ptr());
        c.d_vals = static_cast<uint64_t *>(c.vals_buf.ptr());
        // rows nobody has set yet read as "missing"
        CUDA_TRY(cudaMemsetAsync(c.d_tags + c.init_rows, 0, rows - c.init_rows, sh.copy_stream));
        CUDA_TRY(cudaMemsetAsync(c.d_vals + c.init_rows, 0, (rows - c.init_rows) * sizeof(uint64_t),
                                 sh.copy_stream));
        CUDA_TRY(cudaStreamSynchronize(sh.copy_stream));
        c.init_rows = rows;
    }
    return NM_OK;
}

Column &column_of(Shard &sh, uint32_t id) {
    auto &slot = sh.columns[id];
    if (!slot) slot.reset(new Column());
    return *slot;
}

}  // namespace

struct ColumnsSnapshot {
    std::map<uint32_t, std::pair<std::vector<uint8_t>, std::vector<uint64_t>>> cols;  // global rows
    uint64_t rows = 0;
};

namespace {
constexpr size_t kMaskCacheEntries = 8;   // masks kept per shard (distinct filters between mutations)
constexpr size_t kMaskPoolEntries = 12;   // + masks still held by searches in flight

size_t shard_mask_words_cap(const Shard &sh) {
    return (((size_t)std::max<uint64_t>(sh.capacity, sh.rows) + 255) / 256) * 8;
}

int mask_entry_alloc(const Shard &sh, size_t prog_bytes, std::shared_ptr<MaskEntry> *out) {
    std::shared_ptr<MaskEntry> e(new MaskEntry());
    e->device = sh.device;
    // sized for the shard's capacity and a full-size program: recycled across filters
    e->words_cap = shard_mask_words_cap(sh);
    e->prog_cap = std::max<size_t>(prog_bytes, nm::kFilterMaxOps * sizeof(nm::FilterOpDev) + 4096);
    CUDA_TRY(cudaMalloc(&e->d_mask, e->words_cap * 4));
    CUDA_TRY(cudaMalloc(&e->d_prog, e->prog_cap));
    CUDA_TRY(cudaEventCreateWithFlags(&e->ready, cudaEventDisableTiming));
    *out = e;
    return NM_OK;
}
}  // namespace

// Device allocation does not belong on the search path (a cudaMalloc next to a 30 GB mirror was
// measured at 0.2 - 100 ms): the mask buffers a shard's filters will need are allocated when its
// metadata arrives or its capacity changes, and recycled from then on.
int mask_pool_prepare(Shard &sh) {
    CUDA_TRY(cudaSetDevice(sh.device));
    const size_t cap = shard_mask_words_cap(sh);
    std::lock_guard<std::mutex> g(sh.mask_mu);
    for (auto it = sh.mask_free.begin(); it != sh.mask_free.end();)
        it = (*it)->words_cap < cap ? sh.mask_free.erase(it) : it + 1;
    for (auto it = sh.mask_cache.begin(); it != sh.mask_cache.end();)  // (stale as well: the shard grew)
        it = (*it)->words_cap < cap ? sh.mask_cache.erase(it) : it + 1;
    while (sh.mask_cache.size() + sh.mask_free.size() < kMaskPoolEntries) {
        std::shared_ptr<MaskEntry> e;
        if (int rc = mask_entry_alloc(sh, 0, &e)) return rc;
        sh.mask_free.push_back(e);
    }
    return NM_OK;
}

int columns_after_resize(nm_index *idx, Shard &sh) {
    (void)idx;
    if (!sh.columns.empty())
        if (int rc = mask_pool_prepare(sh)) return rc;
    for (auto &kv : sh.columns) {
        Column &c = *kv.second;
        if (c.init_rows > sh.rows) c.init_rows = sh.rows;  // rows were removed: re-zero on regrowth
        int rc = column_reserve(sh, c, sh.rows);
        if (rc) return rc;
    }
    return NM_OK;
}

void columns_drop(Shard &sh) {
    cudaSetDevice(sh.device);
    sh.columns.clear();
    std::lock_guard<std::mutex> g(sh.mask_mu);
    sh.mask_cache.clear();
}

// The row at (src, src_local) takes the slot (dst, dst_local): its column entries move with it.
int columns_swap_remove(nm_index *idx, Shard &dst, uint64_t dst_local, Shard &src, uint64_t src_local) {
    (void)idx;
    std::vector<uint32_t> ids;
    for (auto &kv : dst.columns) ids.push_back(kv.first);
    for (auto &kv : src.columns)
        if (!dst.columns.count(kv.first)) ids.push_back(kv.first);
    for (uint32_t id : ids) {
        auto sit = src.columns.find(id);
        const bool src_has = sit != src.columns.end() && src_local < sit->second->init_rows;
        if (&dst == &src) {
            if (!src_has) continue;
            Column &c = *sit->second;
            CUDA_TRY(cudaSetDevice(src.device));
            int rc = launch_column_move(c.d_tags, c.d_vals, dst_local, src_local, src.copy_stream);
            if (rc) return rc;
            CUDA_TRY(cudaStreamSynchronize(src.copy_stream));
            continue;
        }
        uint8_t tag = 0;
        uint64_t val = 0;
        if (src_has) {
            CUDA_TRY(cudaSetDevice(src.device));
            CUDA_TRY(cudaMemcpy(&tag, sit->second->d_tags + src_local, 1, cudaMemcpyDeviceToHost));
            CUDA_TRY(cudaMemcpy(&val, sit->second->d_vals + src_local, 8, cudaMemcpyDeviceToHost));
        }
        if (tag == nm::kTagMissing && !dst.columns.count(id)) continue;
        CUDA_TRY(cudaSetDevice(dst.device));
        Column &c = column_of(dst, id);
        int rc = column_reserve(dst, c, dst.rows);
        if (rc) return rc;
        CUDA_TRY(cudaMemcpy(c.d_tags + dst_local, &tag, 1, cudaMemcpyHostToDevice));
        CUDA_TRY(cudaMemcpy(c.d_vals + dst_local, &val, 8, cudaMemcpyHostToDevice));
    }
    return NM_OK;
}

int columns_gather(nm_index *idx, std::shared_ptr<ColumnsSnapshot> *out) {
    std::shared_ptr<ColumnsSnapshot> snap(new ColumnsSnapshot());
    snap->rows = idx->total_rows();
    for (auto &shp : idx->shards) {
        Shard &sh = *shp;
        CUDA_TRY(cudaSetDevice(sh.device));
        for (auto &kv : sh.columns) {
            auto &dst = snap->cols[kv.first];
            if (dst.first.empty()) {
                dst.first.assign(snap->rows, 0);
                dst.second.assign(snap->rows, 0);
            }
            const uint64_t n = std::min<uint64_t>(kv.second->init_rows, sh.rows);
            if (!n) continue;
            CUDA_TRY(cudaMemcpy(dst.first.data() + sh.row_base, kv.second->d_tags, n,
                                cudaMemcpyDeviceToHost));
            CUDA_TRY(cudaMemcpy(dst.second.data() + sh.row_base, kv.second->d_vals, n * 8,
                                cudaMemcpyDeviceToHost));
        }
    }
    *out = snap;
    return NM_OK;
}

int columns_scatter(nm_index *idx, const ColumnsSnapshot &snap) {
    for (auto &shp : idx->shards) {
        Shard &sh = *shp;
        CUDA_TRY(cudaSetDevice(sh.device));
        sh.columns.clear();
        {
            std::lock_guard<std::mutex> g(sh.mask_mu);
            sh.mask_cache.clear();
        }
        if (!sh.rows) continue;
        if (!snap.cols.empty())
            if (int rc = mask_pool_prepare(sh)) return rc;  // the shard's capacity may have changed
        for (auto &kv : snap.cols) {
            Column &c = column_of(sh, kv.first);
            int rc = column_reserve(sh, c, sh.rows);
            if (rc) return rc;
            CUDA_TRY(cudaMemcpy(c.d_tags, kv.second.first.data() + sh.row_base, sh.rows,
                                cudaMemcpyHostToDevice));
            CUDA_TRY(cudaMemcpy(c.d_vals, kv.second.second.data() + sh.row_base, sh.rows * 8,
                                cudaMemcpyHostToDevice));
        }
    }
    return NM_OK;
}

int validate_filter_program(const nm_filter_op *prog, uint32_t n_ops, const uint32_t *tables,
                            uint32_t n_table_words) {
    if (!prog || n_ops == 0) return fail(NM_ERR_INVALID_ARGUMENT, "empty filter program");
    if (n_ops > nm::kFilterMaxOps)
        return fail(NM_ERR_INVALID_ARGUMENT, "filter program has %u ops; limit is %u", n_ops,
                    nm::kFilterMaxOps);
    int depth = 0;
    for (uint32_t i = 0; i < n_ops; ++i) {
        const nm_filter_op &op = prog[i];
        switch (op.kind) {
        case NM_F_TRUE:
        case NM_F_FALSE:
        case NM_F_EXISTS: ++depth; break;
        case NM_F_AND:
        case NM_F_OR:
            if (depth < 2) return fail(NM_ERR_INVALID_ARGUMENT, "filter op %u: stack underflow", i);
            --depth;
            break;
        case NM_F_CMP:
            if (op.cmp > NM_C_GE) return fail(NM_ERR_INVALID_ARGUMENT, "filter op %u: bad comparison", i);
            if (op.lit_tag < NM_V_NULL || op.lit_tag > NM_V_FLOAT)
                return fail(NM_ERR_INVALID_ARGUMENT,
                            "filter op %u: literal must be NULL, BOOL, INT or FLOAT (strings go "
                            "through NM_F_STR_TABLE)", i);
            ++depth;
            break;
        case NM_F_STR_TABLE:
            if (op.table_bits &&
                (!tables || (uint64_t)op.table_off + (op.table_bits + 31u) / 32u > n_table_words))
                return fail(NM_ERR_INVALID_ARGUMENT, "filter op %u: table out of range", i);
            ++depth;
            break;
        default: return fail(NM_ERR_INVALID_ARGUMENT, "filter op %u: unknown kind %u", i, op.kind);
        }
        if (depth > 64) return fail(NM_ERR_INVALID_ARGUMENT, "filter program nests deeper than 64");
    }
    if (depth != 1) return fail(NM_ERR_INVALID_ARGUMENT, "filter program leaves %d values", depth);
    return NM_OK;
}

int shard_mask(nm_index *idx, Shard &sh, Workspace &ws, const MaskSpec &spec, uint64_t first_row,
               cudaStream_t stream, const uint32_t **d_mask, std::shared_ptr<MaskEntry> *hold) {
    *d_mask = nullptr;
    if (!spec.any() || sh.rows == 0) return NM_OK;
    const size_t words = (((size_t)sh.rows + 255) / 256) * 8;  // u32, whole row blocks
    if (spec.host_mask) {
        // slice [first_row, first_row + rows) of the caller's bitmask, re-based to bit 0
        if (ws.mask_cap < words) {
            if (ws.d_mask) CUDA_TRY(cudaFree(ws.d_mask));
            ws.mask_cap = 0;
            CUDA_TRY(cudaMalloc(&ws.d_mask, words * 4));
            ws.mask_cap = words;
        }
        const size_t w64 = ((size_t)sh.rows + 63) / 64;
        const uint64_t *src = spec.host_mask + first_row / 64;
        const unsigned sft = (unsigned)(first_row % 64);
        CUDA_TRY(cudaMemsetAsync(ws.d_mask, 0, words * 4, stream));
        if (sft == 0) {
            CUDA_TRY(cudaMemcpyAsync(ws.d_mask, src, w64 * 8, cudaMemcpyHostToDevice, stream));
        } else {
            // the last source word may be the caller's last word: never read past first_row + rows
            const size_t src_words = ((size_t)first_row % 64 + sh.rows + 63) / 64;
            std::vector<uint64_t> tmp(w64);
            for (size_t i = 0; i < w64; ++i) {
                uint64_t lo = src[i] >> sft;
                uint64_t hi = (i + 1 < src_words) ? (src[i + 1] << (64 - sft)) : 0ull;
                tmp[i] = lo | hi;
            }
            CUDA_TRY(cudaMemcpyAsync(ws.d_mask, tmp.data(), w64 * 8, cudaMemcpyHostToDevice, stream));
            CUDA_TRY(cudaStreamSynchronize(stream));  // tmp goes out of scope
        }
        idx->h2d_bytes += w64 * 8;
        // bits past the shard's last row must not rank rows of the next shard's slice
        *d_mask = ws.d_mask;
        return NM_OK;
    }
    // ---- filter program: cached per (program, tables) until the next mutation ----
    std::vector<uint8_t> key((size_t)spec.n_ops * sizeof(nm_filter_op) + (size_t)spec.n_table_words * 4);
    memcpy(key.data(), spec.prog, (size_t)spec.n_ops * sizeof(nm_filter_op));
    if (spec.n_table_words)
        memcpy(key.data() + (size_t)spec.n_ops * sizeof(nm_filter_op), spec.tables,
               (size_t)spec.n_table_words * 4);
    const uint64_t epoch = idx->mutation_epoch.load();
    const size_t ops_bytes = (size_t)spec.n_ops * sizeof(nm::FilterOpDev);
    const size_t tab_bytes = (size_t)spec.n_table_words * 4;
    std::shared_ptr<MaskEntry> e;
    {
        std::lock_guard<std::mutex> g(sh.mask_mu);
        for (auto it = sh.mask_cache.begin(); it != sh.mask_cache.end();) {
            if ((*it)->epoch != epoch) {
                // stale: keep its buffers for the next new filter unless a search still holds it
                // (buffers sized for a smaller capacity are of no use any more: let them go)
                if (it->use_count() == 1 && sh.mask_free.size() < kMaskPoolEntries &&
                    (*it)->words_cap >= shard_mask_words_cap(sh))
                    sh.mask_free.push_back(*it);
                it = sh.mask_cache.erase(it);
                continue;
            }
            if ((*it)->key == key) {
                *hold = *it;
                break;
            }
            ++it;
        }
        if (!*hold)
            for (auto it = sh.mask_free.begin(); it != sh.mask_free.end(); ++it)
                if ((*it)->words_cap >= words && (*it)->prog_cap >= ops_bytes + tab_bytes + 16) {
                    e = *it;
                    sh.mask_free.erase(it);
                    break;
                }
    }
    if (*hold) {
        idx->filter_mask_hits++;
        CUDA_TRY(cudaStreamWaitEvent(stream, (*hold)->ready, 0));
        *d_mask = (*hold)->d_mask;
        return NM_OK;
    }
    if (!e)  // pool exhausted (every entry held by a search in flight) or an oversized string table
        if (int rc = mask_entry_alloc(sh, ops_bytes + tab_bytes + 16, &e)) return rc;
    e->epoch = epoch;
    e->words = words;
    const uint32_t *d_tables = reinterpret_cast<const uint32_t *>(static_cast<uint8_t *>(e->d_prog) + ops_bytes);
    std::vector<nm::FilterOpDev> ops(spec.n_ops);
    for (uint32_t i = 0; i < spec.n_ops; ++i) {
        const nm_filter_op &s = spec.prog[i];
        nm::FilterOpDev &d = ops[i];
        memset(&d, 0, sizeof(d));
        d.kind = s.kind;
        d.cmp = s.cmp;
        d.lit_tag = s.lit_tag;
        d.lit = s.lit;
        if (s.kind == NM_F_EXISTS || s.kind == NM_F_CMP || s.kind == NM_F_STR_TABLE) {
            auto it = sh.columns.find(s.column);
            if (it != sh.columns.end() && it->second->init_rows >= sh.rows) {
                d.tags = it->second->d_tags;
                d.vals = it->second->d_vals;
            }
        }
        if (s.kind == NM_F_STR_TABLE) {
            d.table = d_tables + s.table_off;
            d.table_bits = s.table_bits;
        }
    }
    CUDA_TRY(cudaMemcpyAsync(e->d_prog, ops.data(), ops_bytes, cudaMemcpyHostToDevice, stream));
    if (tab_bytes)
        CUDA_TRY(cudaMemcpyAsync(static_cast<uint8_t *>(e->d_prog) + ops_bytes, spec.tables, tab_bytes,
                                 cudaMemcpyHostToDevice, stream));
    // (pageable sources: both copies have been staged when the calls return)
    uint32_t depth = 0, max_depth = 0;  // validated: never underflows, ends at 1
    for (uint32_t i = 0; i < spec.n_ops; ++i) {
        if (spec.prog[i].kind == NM_F_AND || spec.prog[i].kind == NM_F_OR) --depth;
        else max_depth = std::max(max_depth, ++depth);
    }
    int rc = launch_filter_mask(sh, static_cast<const nm::FilterOpDev *>(e->d_prog), spec.n_ops, max_depth, sh.rows,
                                e->d_mask, words, stream);
    if (rc) return rc;
    CUDA_TRY(cudaEventRecord(e->ready, stream));
    idx->filter_masks_built++;
    idx->h2d_bytes += ops_bytes + tab_bytes;
    e->key.swap(key);
    {
        std::lock_guard<std::mutex> g(sh.mask_mu);
        if (sh.mask_cache.size() >= kMaskCacheEntries) {
            if (sh.mask_cache.front().use_count() == 1 && sh.mask_free.size() < kMaskPoolEntries &&
                sh.mask_cache.front()->words_cap >= shard_mask_words_cap(sh))
                sh.mask_free.push_back(sh.mask_cache.front());
            sh.mask_cache.erase(sh.mask_cache.begin());
        }
        sh.mask_cache.push_back(e);
    }
    *hold = e;
    *d_mask = e->d_mask;
    return NM_OK;
}

}  // namespace nmi

extern "C" {

int nm_index_column_set(nm_index *idx, uint32_t column, uint64_t first_row, uint64_t n,
                        const uint8_t *tags, const uint64_t *values) {
    NM_TRACE("nm_index_column_set");
    if (!idx) return fail(NM_ERR_INVALID_ARGUMENT, "null index");
    if (n == 0) return NM_OK;
    if (!tags || !values) return fail(NM_ERR_INVALID_ARGUMENT, "null column data");
    for (uint64_t i = 0; i < n; ++i)
        if (tags[i] > NM_V_STRING) return fail(NM_ERR_INVALID_ARGUMENT, "row %llu: unknown value tag %u",
                                               (unsigned long long)(first_row + i), tags[i]);
    std::unique_lock<std::shared_mutex> g(idx->mu);
    if (int wrc = wait_async_searches(idx)) return wrc;
    if (first_row + n < first_row || first_row + n > idx->total_rows())
        return fail(NM_ERR_INVALID_ARGUMENT, "rows [%llu, %llu) out of range", (unsigned long long)first_row,
                    (unsigned long long)(first_row + n));
    for (auto &shp : idx->shards) {
        Shard &sh = *shp;
        const uint64_t lo = std::max(first_row, sh.row_base), hi = std::min(first_row + n, sh.row_base + sh.rows);
        if (lo >= hi) continue;
        CUDA_TRY(cudaSetDevice(sh.device));
        Column &c = column_of(sh, column);
        if (c.init_rows > sh.rows) c.init_rows = sh.rows;
        if (int rc = column_reserve(sh, c, sh.rows)) return rc;  // rows never set read as "missing"
        if (int rc = mask_pool_prepare(sh)) return rc;           // filters will follow: no cudaMalloc then
        CUDA_TRY(cudaMemcpyAsync(c.d_tags + (lo - sh.row_base), tags + (lo - first_row), hi - lo,
                                 cudaMemcpyHostToDevice, sh.copy_stream));
        CUDA_TRY(cudaMemcpyAsync(c.d_vals + (lo - sh.row_base), values + (lo - first_row), (hi - lo) * 8,
                                 cudaMemcpyHostToDevice, sh.copy_stream));
        CUDA_TRY(cudaStreamSynchronize(sh.copy_stream));
        idx->h2d_bytes += (hi - lo) * 9;
    }
    idx->mutation_epoch++;
    return NM_OK;
}

int nm_index_filter_mask(nm_index *idx, const nm_filter_op *program, uint32_t n_ops,
                         const uint32_t *tables, uint32_t n_table_words, uint64_t *out_mask,
                         uint64_t *out_eligible) {
    if (!idx) return fail(NM_ERR_INVALID_ARGUMENT, "null index");
    int rc = validate_filter_program(program, n_ops, tables, n_table_words);
    if (rc) return rc;
    std::shared_lock<std::shared_mutex> g(idx->mu);
    const uint64_t total = idx->total_rows();
    if (out_mask) memset(out_mask, 0, ((total + 63) / 64) * 8);
    uint64_t eligible = 0;
    MaskSpec spec;
    spec.prog = program;
    spec.n_ops = n_ops;
    spec.tables = tables;
    spec.n_table_words = n_table_words;
    for (auto &shp : idx->shards) {
        Shard &sh = *shp;
        if (!sh.rows) continue;
        CUDA_TRY(cudaSetDevice(sh.device));
        std::unique_ptr<Workspace> ws;
        rc = ws_acquire(sh, ws);
        if (rc) return rc;
        struct Releaser {
            Shard &s;
            std::unique_ptr<Workspace> &w;
            ~Releaser() { ws_release(s, w); }
        } rel{sh, ws};
        const uint32_t *d_mask = nullptr;
        std::shared_ptr<MaskEntry> hold;
        rc = shard_mask(idx, sh, *ws, spec, sh.row_base, ws->stream, &d_mask, &hold);
        if (rc) return rc;
        const size_t w32 = ((size_t)sh.rows + 31) / 32;
        std::vector<uint32_t> h(w32);
        CUDA_TRY(cudaMemcpyAsync(h.data(), d_mask, w32 * 4, cudaMemcpyDeviceToHost, ws->stream));
        CUDA_TRY(cudaStreamSynchronize(ws->stream));
        for (uint64_t r = 0; r < sh.rows; ++r)
            if ((h[r >> 5] >> (r & 31)) & 1u) {
                ++eligible;
                if (out_mask) {
                    const uint64_t gr = sh.row_base + r;
                    out_mask[gr >> 6] |= 1ull << (gr & 63);
                }
            }
    }
    if (out_eligible) *out_eligible = eligible;
    return NM_OK;
}

}  // extern "C"
