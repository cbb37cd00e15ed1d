// vector_engine.cpp — see vector_engine.hpp.  Host-side key tables + device-mirror upkeep;
// every scan goes to the GPU through nm_search.
#include "vector_engine.hpp"

#include <algorithm>
#include <cmath>
#include <cstring>

#include "../../include/neumann_b200.h"

namespace neumann {

// --------------------------------------------------------------------------------------
// errors
// --------------------------------------------------------------------------------------
int VectorError::status() const {
    switch (kind) {
    case ErrorKind::NotFound: return NM_ERR_NOT_FOUND;
    case ErrorKind::DimensionMismatch: return NM_ERR_DIMENSION_MISMATCH;
    case ErrorKind::EmptyVector: return NM_ERR_EMPTY_VECTOR;
    case ErrorKind::InvalidTopK: return NM_ERR_INVALID_TOP_K;
    case ErrorKind::StorageError: return NM_ERR_STORAGE;
    case ErrorKind::ConfigurationError: return NM_ERR_CONFIGURATION;
    case ErrorKind::CollectionExists: return NM_ERR_COLLECTION_EXISTS;
    case ErrorKind::CollectionNotFound: return NM_ERR_COLLECTION_NOT_FOUND;
    case ErrorKind::SearchTimeout: return NM_ERR_SEARCH_TIMEOUT;
    default: return NM_ERR_INVALID_ARGUMENT;
    }
}

// Display strings follow `impl Display for VectorError` (vector_engine/src/lib.rs:151-182).
std::string VectorError::to_string() const {
    switch (kind) {
    case ErrorKind::NotFound: return "Embedding not found: " + message;
    case ErrorKind::DimensionMismatch:
        return "Dimension mismatch: expected " + std::to_string(expected) + ", got " +
               std::to_string(got);
    case ErrorKind::EmptyVector: return "Empty vector provided";
    case ErrorKind::InvalidTopK: return "Invalid top_k value (must be > 0)";
    case ErrorKind::StorageError: return "Storage error: " + message;
    case ErrorKind::ConfigurationError: return "Configuration error: " + message;
    case ErrorKind::CollectionExists: return "Collection already exists: " + message;
    case ErrorKind::CollectionNotFound: return "Collection not found: " + message;
    case ErrorKind::SearchTimeout:
        return "search timeout: " + operation + " exceeded " + std::to_string(timeout_ms) + "ms";
    case ErrorKind::BatchValidationError:
        return "Batch validation error at index " + std::to_string(index) + ": " + message;
    case ErrorKind::BatchOperationError:
        return "Batch operation error at index " + std::to_string(index) + ": " + message;
    default: return "Invalid argument: " + message;
    }
}

namespace {
VectorError err(ErrorKind k, std::string msg = {}) {
    VectorError e;
    e.kind = k;
    e.message = std::move(msg);
    return e;
}
VectorError dim_mismatch(size_t expected, size_t got) {
    VectorError e;
    e.kind = ErrorKind::DimensionMismatch;
    e.expected = expected;
    e.got = got;
    return e;
}
VectorError storage_from_nm(int code) {
    VectorError e;
    e.kind = ErrorKind::StorageError;  // CUDA / NCCL failures surface as StorageError(msg)
    e.message = std::string(nm_last_error()) + " (nm_status " + std::to_string(code) + ")";
    return e;
}
}  // namespace

// --------------------------------------------------------------------------------------
// simd (tensor_store/src/hnsw.rs:168-229): 8 lane accumulators, separate mul and add, lanes
// folded left to right from 0.0, scalar tail.  Built with -ffp-contract=off.
// --------------------------------------------------------------------------------------
namespace simd {
float dot_product(const float *a, const float *b, size_t n) {
    const size_t chunks = n / 8, rem = n % 8;
    float l[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (size_t c = 0; c < chunks; ++c)
        for (int j = 0; j < 8; ++j) {
            float p = a[c * 8 + j] * b[c * 8 + j];
            l[j] = l[j] + p;
        }
    float r = 0.0f;
    for (int j = 0; j < 8; ++j) r = r + l[j];
    for (size_t i = 0; i < rem; ++i) {
        float p = a[chunks * 8 + i] * b[chunks * 8 + i];
        r = r + p;
    }
    return r;
}
float sum_of_squares(const float *v, size_t n) { return dot_product(v, v, n); }
float magnitude(const float *v, size_t n) { return std::sqrt(sum_of_squares(v, n)); }
}  // namespace simd

// --------------------------------------------------------------------------------------
// config
// --------------------------------------------------------------------------------------
Result<Unit> VectorEngineConfig::validate() const {
    if (sparse_threshold < 0.0f || sparse_threshold > 1.0f || std::isnan(sparse_threshold))
        return err(ErrorKind::ConfigurationError, "sparse_threshold must be between 0.0 and 1.0");
    if (parallel_threshold == 0)
        return err(ErrorKind::ConfigurationError, "parallel_threshold must be greater than 0");
    if (max_dimension && *max_dimension == 0)
        return err(ErrorKind::ConfigurationError, "max_dimension must be greater than 0");
    if (max_keys_per_scan && *max_keys_per_scan == 0)
        return err(ErrorKind::ConfigurationError, "max_keys_per_scan must be greater than 0");
    if (batch_parallel_threshold == 0)
        return err(ErrorKind::ConfigurationError,
                   "batch_parallel_threshold must be greater than 0");
    return Unit{};
}

// --------------------------------------------------------------------------------------
// storage: one Space per key namespace (`emb:` or `coll:{name}:emb:`), rows bucketed by
// dimension because the scan only ever compares equal-length vectors (lib.rs:2126-2128).
// --------------------------------------------------------------------------------------
struct VectorEngine::Bucket {
    uint32_t dim = 0;
    std::vector<std::string> keys;  // row -> key (mirror order)
    std::vector<float> rows;        // host copy, row-major
    std::vector<Metadata> meta;     // row -> metadata fields (empty map when none were stored)
    nm_index *mirror = nullptr;     // device mirror, created at the first search
    uint64_t synced_rows = 0;       // rows [0, synced_rows) are on the device
    std::mutex sync_mu;             // serialises lazy appends issued by concurrent searches
    // Columnar copy of `meta` on the device (nm_index_column_set): one column per field name,
    // strings dictionary-encoded per column.  Filled lazily by the first pre-filtered search and
    // kept up to date incrementally; filters are then evaluated on the device.
    struct ColumnInfo {
        uint32_t id = 0;
        std::unordered_map<std::string, uint32_t> dict;
        std::vector<std::string> strings;
    };
    std::unordered_map<std::string, ColumnInfo> columns;
    uint64_t cols_synced_rows = 0;      // rows [0, cols_synced_rows) have their columns on the device
    std::vector<uint64_t> meta_dirty;   // ... except these (metadata replaced / row moved since)
    ~Bucket() {
        if (mirror) nm_index_destroy(mirror);
    }
};

struct VectorEngine::Space {
    mutable std::shared_mutex mu;
    std::unordered_map<std::string, std::pair<uint32_t, uint64_t>> where;  // key -> (dim, row)
    std::map<uint32_t, std::unique_ptr<Bucket>> buckets;
};

VectorEngine::VectorEngine() : default_space_(new Space()), entity_space_(new Space()) {}
VectorEngine::VectorEngine(VectorEngineConfig config)
    : config_(std::move(config)), default_space_(new Space()), entity_space_(new Space()) {}
VectorEngine::~VectorEngine() = default;

Result<std::unique_ptr<VectorEngine>> VectorEngine::with_config(VectorEngineConfig config) {
    auto v = config.validate();
    if (v.is_err()) return v.error();
    return std::unique_ptr<VectorEngine>(new VectorEngine(std::move(config)));
}

// lib.rs:1876-1885
bool VectorEngine::should_use_sparse_with_threshold(const std::vector<float> &v, float threshold) {
    if (v.empty()) return false;
    size_t nnz = 0;
    for (float x : v)
        if (std::fabs(x) > 1e-6f) ++nnz;
    float zero_ratio = 1.0f - ((float)nnz / (float)v.size());
    return zero_ratio >= threshold;
}

bool VectorEngine::should_use_sparse(const std::vector<float> &v) const {
    return should_use_sparse_with_threshold(v, config_.sparse_threshold);
}

Result<Unit> VectorEngine::store_in_space(Space &sp, const std::string &key,
                                          std::vector<float> vector, const Metadata *metadata) {
    // Sparse storage round-trips through SparseVector::from_dense / to_dense
    // (tensor_store/src/sparse_vector.rs:212-229, 400-406): entries == 0.0 are dropped and come
    // back as +0.0, so -0.0 loses its sign; everything else (incl. NaN) is preserved.
    if (should_use_sparse(vector))
        for (float &x : vector)
            if (x == 0.0f) x = 0.0f;
    const uint32_t dim = (uint32_t)vector.size();
    std::unique_lock<std::shared_mutex> g(sp.mu);
    auto it = sp.where.find(key);
    if (it != sp.where.end() && it->second.first != dim) {
        g.unlock();
        auto d = delete_in_space(sp, key);
        if (d.is_err()) return d.error();
        g.lock();
        it = sp.where.end();
    }
    auto &slot = sp.buckets[dim];
    if (!slot) {
        slot.reset(new Bucket());
        slot->dim = dim;
    }
    Bucket &b = *slot;
    if (it != sp.where.end()) {
        uint64_t row = it->second.second;
        std::memcpy(&b.rows[row * dim], vector.data(), (size_t)dim * 4);
        // store.put replaces the whole TensorData: metadata of the old value is gone
        b.meta[row] = metadata ? *metadata : Metadata{};
        if (row < b.cols_synced_rows) b.meta_dirty.push_back(row);
        if (b.mirror && row < b.synced_rows) {
            int rc = nm_index_update(b.mirror, row, vector.data());
            if (rc) return storage_from_nm(rc);
        }
        return Unit{};
    }
    uint64_t row = b.keys.size();
    b.keys.push_back(key);
    b.meta.push_back(metadata ? *metadata : Metadata{});
    b.rows.insert(b.rows.end(), vector.begin(), vector.end());
    sp.where[key] = {dim, row};
    return Unit{};  // the device append is deferred to the next search (batched)
}

Result<Unit> VectorEngine::delete_in_space(Space &sp, const std::string &key) {
    std::unique_lock<std::shared_mutex> g(sp.mu);
    auto it = sp.where.find(key);
    if (it == sp.where.end()) return err(ErrorKind::NotFound, key);
    const uint32_t dim = it->second.first;
    const uint64_t row = it->second.second;
    Bucket &b = *sp.buckets[dim];
    const uint64_t last = b.keys.size() - 1;
    if (b.mirror && row < b.synced_rows) {
        // bring the device up to date so that "last row" means the same on both sides
        if (b.synced_rows < b.keys.size()) {
            int rc = nm_index_append(b.mirror, &b.rows[b.synced_rows * dim],
                                     b.keys.size() - b.synced_rows);
            if (rc) return storage_from_nm(rc);
            b.synced_rows = b.keys.size();
        }
        uint64_t moved = 0;
        int rc = nm_index_swap_remove(b.mirror, row, &moved);
        if (rc) return storage_from_nm(rc);
        b.synced_rows -= 1;
        // the device moved the last row's column entries along; they are only right if that row's
        // columns had been pushed already — re-push the slot at the next filtered search
        if (row < b.cols_synced_rows && row != last) b.meta_dirty.push_back(row);
        b.cols_synced_rows = std::min(b.cols_synced_rows, b.synced_rows);
    }
    if (row != last) {
        std::memcpy(&b.rows[row * dim], &b.rows[last * dim], (size_t)dim * 4);
        b.keys[row] = std::move(b.keys[last]);
        b.meta[row] = std::move(b.meta[last]);
        sp.where[b.keys[row]] = {dim, row};
    }
    b.keys.pop_back();
    b.meta.pop_back();
    b.rows.resize(b.keys.size() * (size_t)dim);
    sp.where.erase(key);
    if (b.keys.empty()) sp.buckets.erase(dim);
    return Unit{};
}

Result<std::vector<float>> VectorEngine::get_in_space(const Space &sp,
                                                      const std::string &key) const {
    std::shared_lock<std::shared_mutex> g(sp.mu);
    auto it = sp.where.find(key);
    if (it == sp.where.end()) return err(ErrorKind::NotFound, key);
    const Bucket &b = *sp.buckets.at(it->second.first);
    const float *p = &b.rows[it->second.second * (size_t)b.dim];
    return std::vector<float>(p, p + b.dim);
}

// ---- device-side filters: columnar metadata + FilterCondition -> postfix program ----------
namespace {

// compare_tensor_value_to_filter restricted to two strings (lib.rs:3658-3684)
bool string_cmp_holds(const std::string &a, const std::string &b, FilterCondition::Op op) {
    const int c = a.compare(b);
    switch (op) {
    case FilterCondition::Op::Eq: return c == 0;
    case FilterCondition::Op::Ne: return c != 0;
    case FilterCondition::Op::Lt: return c < 0;
    case FilterCondition::Op::Le: return c <= 0;
    case FilterCondition::Op::Gt: return c > 0;
    default: return c >= 0;
    }
}

struct FilterProgram {
    std::vector<nm_filter_op> ops;
    std::vector<uint32_t> tables;
    bool ok = true;  // false: too large for the device evaluator
    void push(nm_filter_op op) { ops.push_back(op); }
    void constant(bool v) {
        nm_filter_op op{};
        op.kind = v ? NM_F_TRUE : NM_F_FALSE;
        push(op);
    }
};

// Appends the postfix code of `f`.  Everything that needs the bytes of a string is evaluated
// here once per DISTINCT string of the column into a bit table.
template <class BucketT>
void compile_filter(const BucketT &b, const FilterCondition &f, FilterProgram &out) {
    using Op = FilterCondition::Op;
    using T = MetadataValue::Type;
    if (f.op == Op::True) return out.constant(true);
    if (f.op == Op::And || f.op == Op::Or) {
        compile_filter(b, *f.lhs, out);
        compile_filter(b, *f.rhs, out);
        nm_filter_op op{};
        op.kind = f.op == Op::And ? NM_F_AND : NM_F_OR;
        return out.push(op);
    }
    auto cit = b.columns.find(f.field);
    if (cit == b.columns.end()) return out.constant(false);  // no row has the field
    const auto &col = cit->second;
    auto table_leaf = [&](auto pred) {
        nm_filter_op op{};
        op.kind = NM_F_STR_TABLE;
        op.column = col.id;
        op.table_off = (uint32_t)out.tables.size();
        op.table_bits = (uint32_t)col.strings.size();
        out.tables.resize(out.tables.size() + (col.strings.size() + 31) / 32, 0u);
        for (size_t c = 0; c < col.strings.size(); ++c)
            if (pred(col.strings[c])) out.tables[op.table_off + c / 32] |= 1u << (c % 32);
        out.push(op);
    };
    auto cmp_leaf = [&](const FilterValue &v, Op o) {
        if (v.type == T::String) return table_leaf([&](const std::string &s) { return string_cmp_holds(s, v.s, o); });
        nm_filter_op op{};
        op.kind = NM_F_CMP;
        op.column = col.id;
        op.cmp = o == Op::Eq ? NM_C_EQ : o == Op::Ne ? NM_C_NE : o == Op::Lt ? NM_C_LT
                 : o == Op::Le ? NM_C_LE : o == Op::Gt ? NM_C_GT : NM_C_GE;
        switch (v.type) {
        case T::Null: op.lit_tag = NM_V_NULL; break;
        case T::Bool: op.lit_tag = NM_V_BOOL; op.lit = v.b ? 1 : 0; break;
        case T::Int: op.lit_tag = NM_V_INT; op.lit = (uint64_t)v.i; break;
        default: op.lit_tag = NM_V_FLOAT; std::memcpy(&op.lit, &v.f, 8); break;
        }
        out.push(op);
    };
    switch (f.op) {
    case Op::Exists: {
        nm_filter_op op{};
        op.kind = NM_F_EXISTS;
        op.column = col.id;
        return out.push(op);
    }
    case Op::Contains:
        return table_leaf([&](const std::string &s) { return s.find(f.value.s) != std::string::npos; });
    case Op::StartsWith:
        return table_leaf([&](const std::string &s) { return s.compare(0, f.value.s.size(), f.value.s) == 0; });
    case Op::In: {
        if (f.values.empty()) return out.constant(false);
        size_t pushed = 0;
        bool any_string = false;
        for (auto &v : f.values) any_string |= v.type == T::String;
        if (any_string) {
            table_leaf([&](const std::string &s) {
                for (auto &v : f.values)
                    if (v.type == T::String && v.s == s) return true;
                return false;
            });
            ++pushed;
        }
        for (auto &v : f.values) {
            if (v.type == T::String) continue;
            cmp_leaf(v, Op::Eq);
            if (++pushed > 1) {
                nm_filter_op op{};
                op.kind = NM_F_OR;
                out.push(op);
            }
        }
        return;
    }
    default: return cmp_leaf(f.value, f.op);
    }
}

// Column entries of rows [first, first + n): one pass over the rows' metadata maps per 1M-row
// chunk, `sink(column id, first row, rows, tags, vals)` once per column that occurs in the chunk
// (every known column when `all_columns`: rows being re-pushed must overwrite absent fields with
// "missing").  Registers new field names and new distinct strings on the way.
template <class BucketT, class Sink>
int stage_columns(BucketT &b, uint64_t first, uint64_t n, bool all_columns, Sink sink) {
    constexpr uint64_t kChunk = 1u << 20;
    struct Staged {
        std::vector<uint8_t> tags;
        std::vector<uint64_t> vals;
    };
    for (uint64_t c0 = first; c0 < first + n; c0 += kChunk) {
        const uint64_t m = std::min(kChunk, first + n - c0);
        std::unordered_map<uint32_t, Staged> staged;
        if (all_columns)
            for (auto &ckv : b.columns) {
                Staged &st = staged[ckv.second.id];
                st.tags.assign(m, NM_V_MISSING);
                st.vals.assign(m, 0);
            }
        for (uint64_t i = 0; i < m; ++i)
            for (auto &kv : b.meta[c0 + i]) {
                auto ins = b.columns.try_emplace(kv.first);
                auto &col = ins.first->second;
                if (ins.second) col.id = (uint32_t)b.columns.size();  // ids 1, 2, ...
                Staged &st = staged[col.id];
                if (st.tags.empty()) {
                    st.tags.assign(m, NM_V_MISSING);
                    st.vals.assign(m, 0);
                }
                const MetadataValue &v = kv.second;
                switch (v.type) {
                case MetadataValue::Type::Null: st.tags[i] = NM_V_NULL; break;
                case MetadataValue::Type::Bool: st.tags[i] = NM_V_BOOL; st.vals[i] = v.b ? 1 : 0; break;
                case MetadataValue::Type::Int: st.tags[i] = NM_V_INT; st.vals[i] = (uint64_t)v.i; break;
                case MetadataValue::Type::Float:
                    st.tags[i] = NM_V_FLOAT;
                    std::memcpy(&st.vals[i], &v.f, 8);
                    break;
                case MetadataValue::Type::String: {
                    auto d = col.dict.try_emplace(v.s, (uint32_t)col.strings.size());
                    if (d.second) col.strings.push_back(v.s);
                    st.tags[i] = NM_V_STRING;
                    st.vals[i] = d.first->second;
                    break;
                }
                }
            }
        for (auto &kv : staged) {
            int rc = sink(kv.first, c0, m, kv.second.tags, kv.second.vals);
            if (rc) return rc;
        }
    }
    return 0;
}

// Push the column entries of rows [first, first + n) to the device.
template <class BucketT>
int push_columns(BucketT &b, uint64_t first, uint64_t n) {
    const bool repush = first < b.cols_synced_rows;
    return stage_columns(b, first, n, repush,
                         [&](uint32_t col, uint64_t row0, uint64_t m, const std::vector<uint8_t> &tags,
                             const std::vector<uint64_t> &vals) {
                             return nm_index_column_set(b.mirror, col, row0, m, tags.data(), vals.data());
                         });
}

}  // namespace

// The seam of SURVEY 8b: "store.scan -> search_* -> sort_by -> truncate" becomes one nm_search.
Result<std::vector<SearchResult>> VectorEngine::scan_space(
    const Space &sp, const std::vector<float> &query, size_t top_k, DistanceMetric metric,
    const char *operation, std::chrono::steady_clock::time_point start,
    const FilterCondition *pre_filter) const {
    auto expired = [&]() {
        if (!config_.search_timeout) return false;
        return std::chrono::steady_clock::now() - start >= *config_.search_timeout;
    };
    auto timeout_err = [&]() {
        VectorError e;
        e.kind = ErrorKind::SearchTimeout;
        e.operation = operation;
        e.timeout_ms = (uint64_t)config_.search_timeout->count();
        return e;
    };
    std::shared_lock<std::shared_mutex> g(sp.mu);
    auto bit = sp.buckets.find((uint32_t)query.size());
    // first deadline check: after the reference's store.scan (lib.rs:2005-2010)
    if (expired()) return timeout_err();
    if (bit == sp.buckets.end() || bit->second->keys.empty())
        return std::vector<SearchResult>{};  // rows of other dimensions are skipped
    Bucket &b = *bit->second;
    // "filter first, then search the subset" (lib.rs:3514-3557): the filter is compiled to a
    // postfix program and evaluated ON THE DEVICE over the columnar copy of the metadata
    // (nm_search_filtered); only a program too large for the device evaluator falls back to a
    // host-built bitmask.
    FilterProgram prog;
    std::vector<uint64_t> mask;
    {
        std::lock_guard<std::mutex> sg(b.sync_mu);
        if (!b.mirror) {
            const int *devs = config_.devices.empty() ? nullptr : config_.devices.data();
            int rc = nm_index_create(b.dim, devs, (int)config_.devices.size(), &b.mirror);
            if (rc) return storage_from_nm(rc);
            if (config_.device_prefilter) {
                rc = nm_index_set_prefilter(b.mirror, 1);
                if (rc) return storage_from_nm(rc);
            }
        }
        if (b.synced_rows < b.keys.size()) {
            int rc = nm_index_append(b.mirror, &b.rows[b.synced_rows * (size_t)b.dim],
                                     b.keys.size() - b.synced_rows);
            if (rc) return storage_from_nm(rc);
            b.synced_rows = b.keys.size();
        }
        if (pre_filter) {
            // metadata columns: rows not pushed yet, then rows whose metadata changed since
            int rc = push_columns(b, b.cols_synced_rows, b.keys.size() - b.cols_synced_rows);
            if (rc) return storage_from_nm(rc);
            std::sort(b.meta_dirty.begin(), b.meta_dirty.end());
            b.meta_dirty.erase(std::unique(b.meta_dirty.begin(), b.meta_dirty.end()), b.meta_dirty.end());
            for (uint64_t r : b.meta_dirty)
                if (r < b.cols_synced_rows) {
                    rc = push_columns(b, r, 1);
                    if (rc) return storage_from_nm(rc);
                }
            b.meta_dirty.clear();
            b.cols_synced_rows = b.keys.size();
            compile_filter(b, *pre_filter, prog);
            int depth = 0, max_depth = 0;
            for (auto &op : prog.ops) {
                depth += (op.kind == NM_F_AND || op.kind == NM_F_OR) ? -1 : 1;
                max_depth = std::max(max_depth, depth);
            }
            prog.ok = prog.ops.size() <= 128 && max_depth <= 64;
        }
    }
    if (pre_filter && !prog.ok) {
        mask.assign((b.keys.size() + 63) / 64, 0ull);
        size_t eligible = 0;
        for (size_t r = 0; r < b.keys.size(); ++r)
            if (evaluate_filter(b.meta[r], *pre_filter)) {
                mask[r >> 6] |= 1ull << (r & 63);
                ++eligible;
            }
        if (eligible == 0) return std::vector<SearchResult>{};
    }
    // The device path serves k <= NM_TOPK_FAST_MAX per call; clamp to the row count first
    // (truncate(top_k) on fewer rows returns them all, lib.rs:2034).
    size_t k = std::min<size_t>(top_k, b.keys.size());
    std::vector<uint64_t> rows(k);
    std::vector<float> scores(k);
    uint32_t count = 0;
    int rc;
    if (pre_filter && prog.ok) {
        rc = nm_search_filtered(b.mirror, query.data(), 1, (uint32_t)k, (int)metric, prog.ops.data(),
                                (uint32_t)prog.ops.size(), prog.tables.empty() ? nullptr : prog.tables.data(),
                                (uint32_t)prog.tables.size(), rows.data(), scores.data(), &count);
    } else if (pre_filter) {
        rc = nm_search_masked(b.mirror, query.data(), 1, (uint32_t)k, (int)metric, mask.data(),
                              rows.data(), scores.data(), &count);
    } else {
        rc = nm_search(b.mirror, query.data(), 1, (uint32_t)k, (int)metric, rows.data(),
                       scores.data(), &count);
    }
    if (rc) return storage_from_nm(rc);
    // second deadline check: after scoring (lib.rs:2019-2024)
    if (expired()) return timeout_err();
    std::vector<SearchResult> out;
    out.reserve(count);
    for (uint32_t i = 0; i < count; ++i) out.push_back(SearchResult{b.keys[rows[i]], scores[i]});
    return out;
}

Result<std::vector<std::vector<SearchResult>>> VectorEngine::scan_space_batch(
    const Space &sp, const float *queries, size_t nq, size_t dim, size_t top_k,
    DistanceMetric metric, const char *operation,
    std::chrono::steady_clock::time_point start) const {
    auto expired = [&]() {
        if (!config_.search_timeout) return false;
        return std::chrono::steady_clock::now() - start >= *config_.search_timeout;
    };
    auto timeout_err = [&]() {
        VectorError e;
        e.kind = ErrorKind::SearchTimeout;
        e.operation = operation;
        e.timeout_ms = (uint64_t)config_.search_timeout->count();
        return e;
    };
    std::vector<std::vector<SearchResult>> out(nq);
    std::shared_lock<std::shared_mutex> g(sp.mu);
    auto bit = sp.buckets.find((uint32_t)dim);
    if (expired()) return timeout_err();
    if (bit == sp.buckets.end() || bit->second->keys.empty()) return out;
    Bucket &b = *bit->second;
    {
        std::lock_guard<std::mutex> sg(b.sync_mu);
        if (!b.mirror) {
            const int *devs = config_.devices.empty() ? nullptr : config_.devices.data();
            int rc = nm_index_create(b.dim, devs, (int)config_.devices.size(), &b.mirror);
            if (rc) return storage_from_nm(rc);
            if (config_.device_prefilter) {
                rc = nm_index_set_prefilter(b.mirror, 1);
                if (rc) return storage_from_nm(rc);
            }
        }
        if (b.synced_rows < b.keys.size()) {
            int rc = nm_index_append(b.mirror, &b.rows[b.synced_rows * (size_t)b.dim],
                                     b.keys.size() - b.synced_rows);
            if (rc) return storage_from_nm(rc);
            b.synced_rows = b.keys.size();
        }
    }
    const size_t k = std::min<size_t>(top_k, b.keys.size());
    std::vector<uint64_t> rows(nq * k);
    std::vector<float> scores(nq * k);
    std::vector<uint32_t> counts(nq, 0);
    int rc = nm_search(b.mirror, queries, (uint32_t)nq, (uint32_t)k, (int)metric, rows.data(),
                       scores.data(), counts.data());
    if (rc) return storage_from_nm(rc);
    if (expired()) return timeout_err();
    for (size_t q = 0; q < nq; ++q) {
        out[q].reserve(counts[q]);
        for (uint32_t i = 0; i < counts[q]; ++i)
            out[q].push_back(SearchResult{b.keys[rows[q * k + i]], scores[q * k + i]});
    }
    return out;
}

// --------------------------------------------------------------------------------------
// public API
// --------------------------------------------------------------------------------------
Result<Unit> VectorEngine::store_embedding(const std::string &key, std::vector<float> vector) {
    if (vector.empty()) return err(ErrorKind::EmptyVector);
    if (config_.max_dimension && vector.size() > *config_.max_dimension)
        return dim_mismatch(*config_.max_dimension, vector.size());
    return store_in_space(*default_space_, key, std::move(vector));
}

Result<std::vector<float>> VectorEngine::get_embedding(const std::string &key) const {
    return get_in_space(*default_space_, key);
}

Result<Unit> VectorEngine::delete_embedding(const std::string &key) {
    return delete_in_space(*default_space_, key);
}

bool VectorEngine::exists(const std::string &key) const {
    std::shared_lock<std::shared_mutex> g(default_space_->mu);
    return default_space_->where.count(key) != 0;
}

size_t VectorEngine::count() const {
    std::shared_lock<std::shared_mutex> g(default_space_->mu);
    return default_space_->where.size();
}

std::optional<size_t> VectorEngine::dimension() const {
    std::shared_lock<std::shared_mutex> g(default_space_->mu);
    if (default_space_->buckets.empty()) return std::nullopt;
    return (size_t)default_space_->buckets.begin()->first;
}

Result<VectorEngine::BatchResult> VectorEngine::batch_store_embeddings(
    const std::vector<EmbeddingInput> &inputs) {
    BatchResult out;
    for (size_t i = 0; i < inputs.size(); ++i)
        if (inputs[i].vector.empty()) {
            VectorError e = err(ErrorKind::BatchValidationError, "Empty vector provided");
            e.index = i;
            return e;
        }
    out.stored_keys.reserve(inputs.size());
    for (size_t i = 0; i < inputs.size(); ++i) {
        auto r = store_embedding(inputs[i].key, inputs[i].vector);
        if (r.is_err()) {
            VectorError e = err(ErrorKind::BatchOperationError, r.error().to_string());
            e.index = i;
            return e;
        }
        out.stored_keys.push_back(inputs[i].key);
    }
    out.count = out.stored_keys.size();
    return out;
}

Result<std::vector<SearchResult>> VectorEngine::search_similar(const std::vector<float> &query,
                                                               size_t top_k) const {
    auto start = std::chrono::steady_clock::now();
    if (query.empty()) return err(ErrorKind::EmptyVector);
    if (top_k == 0) return err(ErrorKind::InvalidTopK);
    if (config_.max_dimension && query.size() > *config_.max_dimension)
        return dim_mismatch(*config_.max_dimension, query.size());
    float qmag = simd::magnitude(query.data(), query.size());
    if (qmag == 0.0f) return std::vector<SearchResult>{};  // lib.rs:1970-1974
    return scan_space(*default_space_, query, top_k, DistanceMetric::Cosine, "search_similar",
                      start);
}

Result<std::vector<SearchResult>> VectorEngine::search_similar_with_metric(
    const std::vector<float> &query, size_t top_k, DistanceMetric metric) const {
    auto start = std::chrono::steady_clock::now();
    if (query.empty()) return err(ErrorKind::EmptyVector);
    if (top_k == 0) return err(ErrorKind::InvalidTopK);
    float qmag = simd::magnitude(query.data(), query.size());
    if (qmag == 0.0f && metric != DistanceMetric::Euclidean)  // lib.rs:2066
        return std::vector<SearchResult>{};
    return scan_space(*default_space_, query, top_k, metric, "search_similar_with_metric", start);
}

Result<std::vector<std::vector<SearchResult>>> VectorEngine::search_similar_batch(
    const std::vector<std::vector<float>> &queries, size_t top_k, DistanceMetric metric) const {
    auto start = std::chrono::steady_clock::now();
    // the checks of search_similar_with_metric, in call order (lib.rs:2049-2066)
    for (auto &q : queries) {
        if (q.empty()) return err(ErrorKind::EmptyVector);
        if (top_k == 0) return err(ErrorKind::InvalidTopK);
    }
    std::vector<std::vector<SearchResult>> out(queries.size());
    // queries of one dimension share a device call; zero-magnitude queries short-circuit
    std::map<size_t, std::vector<size_t>> by_dim;
    for (size_t i = 0; i < queries.size(); ++i) {
        const auto &q = queries[i];
        if (metric != DistanceMetric::Euclidean && simd::magnitude(q.data(), q.size()) == 0.0f)
            continue;
        by_dim[q.size()].push_back(i);
    }
    for (auto &kv : by_dim) {
        const size_t dim = kv.first, nq = kv.second.size();
        std::vector<float> flat(nq * dim);
        for (size_t j = 0; j < nq; ++j)
            std::copy(queries[kv.second[j]].begin(), queries[kv.second[j]].end(),
                      flat.begin() + j * dim);
        auto r = scan_space_batch(*default_space_, flat.data(), nq, dim, top_k, metric,
                                  "search_similar_batch", start);
        if (r.is_err()) return r.error();
        for (size_t j = 0; j < nq; ++j) out[kv.second[j]] = std::move(r.value()[j]);
    }
    return out;
}

Result<float> VectorEngine::compute_similarity(const std::vector<float> &a,
                                               const std::vector<float> &b) {
    if (a.empty() || b.empty()) return err(ErrorKind::EmptyVector);
    if (a.size() != b.size()) return dim_mismatch(a.size(), b.size());
    float am = simd::magnitude(a.data(), a.size());
    if (am == 0.0f) return 0.0f;
    float dot = simd::dot_product(a.data(), b.data(), a.size());
    float bm = simd::magnitude(b.data(), b.size());
    if (am == 0.0f || bm == 0.0f) return 0.0f;
    float den = am * bm;
    return dot / den;
}

// ---- collections ----
struct VectorEngine::CollectionEntry {
    std::optional<VectorCollectionConfig> config;
    std::shared_ptr<Space> space{new Space()};
};

std::shared_ptr<VectorEngine::Space> VectorEngine::collection_space(const std::string &name) {
    std::unique_lock<std::shared_mutex> g(collections_mu_);
    auto &e = collections_[name];
    if (!e) e.reset(new CollectionEntry());
    return e->space;
}

std::shared_ptr<const VectorEngine::Space> VectorEngine::find_collection_space(
    const std::string &name) const {
    std::shared_lock<std::shared_mutex> g(collections_mu_);
    auto it = collections_.find(name);
    if (it == collections_.end()) return nullptr;
    return it->second->space;
}

Result<Unit> VectorEngine::create_collection(const std::string &name,
                                             VectorCollectionConfig config) {
    std::unique_lock<std::shared_mutex> g(collections_mu_);
    auto &e = collections_[name];
    if (e && e->config) return err(ErrorKind::CollectionExists, name);
    if (!e) e.reset(new CollectionEntry());
    e->config = config;
    return Unit{};
}

Result<Unit> VectorEngine::delete_collection(const std::string &name) {
    std::unique_lock<std::shared_mutex> g(collections_mu_);
    auto it = collections_.find(name);
    if (it == collections_.end() || !it->second->config)
        return err(ErrorKind::CollectionNotFound, name);
    // drops the map's reference; operations still running on the collection hold their own, the
    // last one out destroys the rows and the device mirror
    collections_.erase(it);
    return Unit{};
}

bool VectorEngine::collection_exists(const std::string &name) const {
    std::shared_lock<std::shared_mutex> g(collections_mu_);
    auto it = collections_.find(name);
    return it != collections_.end() && it->second->config.has_value();
}

std::vector<std::string> VectorEngine::list_collections() const {
    std::shared_lock<std::shared_mutex> g(collections_mu_);
    std::vector<std::string> out;
    for (auto &kv : collections_)
        if (kv.second->config) out.push_back(kv.first);
    return out;
}

Result<Unit> VectorEngine::store_in_collection(const std::string &collection,
                                               const std::string &key,
                                               std::vector<float> vector) {
    if (vector.empty()) return err(ErrorKind::EmptyVector);
    {
        std::shared_lock<std::shared_mutex> g(collections_mu_);
        auto it = collections_.find(collection);
        if (it != collections_.end() && it->second->config && it->second->config->dimension &&
            vector.size() != *it->second->config->dimension)
            return dim_mismatch(*it->second->config->dimension, vector.size());
    }
    if (config_.max_dimension && vector.size() > *config_.max_dimension)
        return dim_mismatch(*config_.max_dimension, vector.size());
    std::shared_ptr<Space> sp = collection_space(collection);
    return store_in_space(*sp, key, std::move(vector));
}

Result<std::vector<float>> VectorEngine::get_from_collection(const std::string &collection,
                                                             const std::string &key) const {
    std::shared_ptr<const Space> sp = find_collection_space(collection);
    if (!sp) return err(ErrorKind::NotFound, collection + ":" + key);
    auto r = get_in_space(*sp, key);
    if (r.is_err()) return err(ErrorKind::NotFound, collection + ":" + key);
    return r;
}

Result<Unit> VectorEngine::delete_from_collection(const std::string &collection,
                                                  const std::string &key) {
    std::shared_ptr<const Space> sp = find_collection_space(collection);
    if (!sp) return err(ErrorKind::NotFound, collection + ":" + key);
    auto r = delete_in_space(*const_cast<Space *>(sp.get()), key);
    if (r.is_err()) return err(ErrorKind::NotFound, collection + ":" + key);
    return r;
}

size_t VectorEngine::collection_count(const std::string &collection) const {
    std::shared_ptr<const Space> sp = find_collection_space(collection);
    if (!sp) return 0;
    std::shared_lock<std::shared_mutex> g(sp->mu);
    return sp->where.size();
}

// lib.rs:1585-1689
Result<std::vector<SearchResult>> VectorEngine::search_in_collection(
    const std::string &collection, const std::vector<float> &query, size_t top_k) const {
    auto start = std::chrono::steady_clock::now();
    if (query.empty()) return err(ErrorKind::EmptyVector);
    if (top_k == 0) return err(ErrorKind::InvalidTopK);
    DistanceMetric metric = DistanceMetric::Cosine;
    std::shared_ptr<const Space> sp;  // keeps the collection alive across a concurrent delete
    {
        std::shared_lock<std::shared_mutex> g(collections_mu_);
        auto it = collections_.find(collection);
        if (it != collections_.end()) {
            if (it->second->config) {
                if (it->second->config->dimension &&
                    query.size() != *it->second->config->dimension)
                    return dim_mismatch(*it->second->config->dimension, query.size());
                metric = it->second->config->distance_metric;
            }
            sp = it->second->space;
        }
    }
    float qmag = simd::magnitude(query.data(), query.size());
    if (qmag == 0.0f && metric == DistanceMetric::Cosine)  // lib.rs:1618-1620 (cosine only)
        return std::vector<SearchResult>{};
    if (!sp) return std::vector<SearchResult>{};
    return scan_space(*sp, query, top_k, metric, "search_in_collection", start);
}

// ---- metadata + filtered search ----
Result<Unit> VectorEngine::store_embedding_with_metadata(const std::string &key,
                                                         std::vector<float> vector,
                                                         Metadata metadata) {
    if (vector.empty()) return err(ErrorKind::EmptyVector);
    if (config_.max_dimension && vector.size() > *config_.max_dimension)
        return dim_mismatch(*config_.max_dimension, vector.size());
    return store_in_space(*default_space_, key, std::move(vector), &metadata);
}

Result<Metadata> VectorEngine::get_metadata(const std::string &key) const {
    const Space &sp = *default_space_;
    std::shared_lock<std::shared_mutex> g(sp.mu);
    auto it = sp.where.find(key);
    if (it == sp.where.end()) return err(ErrorKind::NotFound, key);
    return sp.buckets.at(it->second.first)->meta[it->second.second];
}

Result<Metadata> VectorEngine::get_collection_metadata(const std::string &collection,
                                                       const std::string &key) const {
    std::shared_ptr<const Space> sp = find_collection_space(collection);
    if (!sp) return err(ErrorKind::NotFound, key);
    std::shared_lock<std::shared_mutex> g(sp->mu);
    auto it = sp->where.find(key);
    if (it == sp->where.end()) return err(ErrorKind::NotFound, key);
    return sp->buckets.at(it->second.first)->meta[it->second.second];
}

Result<Unit> VectorEngine::update_metadata(const std::string &key, const Metadata &metadata) {
    Space &sp = *default_space_;
    std::unique_lock<std::shared_mutex> g(sp.mu);
    auto it = sp.where.find(key);
    if (it == sp.where.end()) return err(ErrorKind::NotFound, key);
    Bucket &b = *sp.buckets.at(it->second.first);
    const uint64_t row = it->second.second;
    for (auto &kv : metadata) b.meta[row][kv.first] = kv.second;
    if (row < b.cols_synced_rows) b.meta_dirty.push_back(row);
    return Unit{};
}

Result<Unit> VectorEngine::remove_metadata_field(const std::string &key, const std::string &field) {
    Space &sp = *default_space_;
    std::unique_lock<std::shared_mutex> g(sp.mu);
    auto it = sp.where.find(key);
    if (it == sp.where.end()) return err(ErrorKind::NotFound, key);
    Bucket &b = *sp.buckets.at(it->second.first);
    const uint64_t row = it->second.second;
    if (b.meta[row].erase(field) && row < b.cols_synced_rows) b.meta_dirty.push_back(row);
    return Unit{};
}

bool VectorEngine::has_metadata_field(const std::string &key, const std::string &field) const {
    const Space &sp = *default_space_;
    std::shared_lock<std::shared_mutex> g(sp.mu);
    auto it = sp.where.find(key);
    if (it == sp.where.end()) return false;
    return sp.buckets.at(it->second.first)->meta[it->second.second].count(field) != 0;
}

Result<std::optional<MetadataValue>> VectorEngine::get_metadata_field(const std::string &key,
                                                                      const std::string &field) const {
    const Space &sp = *default_space_;
    std::shared_lock<std::shared_mutex> g(sp.mu);
    auto it = sp.where.find(key);
    if (it == sp.where.end()) return err(ErrorKind::NotFound, key);
    const Metadata &m = sp.buckets.at(it->second.first)->meta[it->second.second];
    auto f = m.find(field);
    if (f == m.end()) return std::optional<MetadataValue>{};
    return std::optional<MetadataValue>{f->second};
}

// lib.rs:2312-2354.  The reference lists keys in HashSet order; here: bucket by bucket, row order.
std::vector<std::string> VectorEngine::list_keys() const { return list_keys_bounded(); }

std::vector<std::string> VectorEngine::list_keys_bounded() const {
    const size_t limit = config_.max_keys_per_scan.value_or(SIZE_MAX);
    std::vector<std::string> out;
    std::shared_lock<std::shared_mutex> g(default_space_->mu);
    for (auto &kv : default_space_->buckets)
        for (auto &k : kv.second->keys) {
            if (out.size() >= limit) return out;
            out.push_back(k);
        }
    return out;
}

VectorEngine::PagedResult<std::string> VectorEngine::list_keys_paginated(Pagination pagination) const {
    const size_t max_scan = config_.max_keys_per_scan.value_or(SIZE_MAX);
    const size_t want = pagination.limit.value_or(max_scan);
    const size_t fetch_limit =
        std::min(pagination.skip > SIZE_MAX - want ? SIZE_MAX : pagination.skip + want, max_scan);
    PagedResult<std::string> out;
    size_t total = 0;
    {
        std::shared_lock<std::shared_mutex> g(default_space_->mu);
        size_t seen = 0;
        for (auto &kv : default_space_->buckets) {
            for (auto &k : kv.second->keys) {
                if (seen >= fetch_limit) break;
                if (seen >= pagination.skip && out.items.size() < pagination.limit.value_or(SIZE_MAX))
                    out.items.push_back(k);
                ++seen;
            }
        }
        total = default_space_->where.size();
    }
    if (pagination.count_total) {
        out.total_count = total;
        out.has_more = (pagination.skip > SIZE_MAX - out.items.size() ? SIZE_MAX
                                                                        : pagination.skip + out.items.size()) < total;
    } else {
        out.has_more = out.items.size() == pagination.limit.value_or(0);
    }
    return out;
}

Result<size_t> VectorEngine::clear() {
    // up to max_keys_per_scan embeddings per call (lib.rs:2340-2354): call again until 0 comes back
    {
        Space &sp = *default_space_;
        std::unique_lock<std::shared_mutex> g(sp.mu);
        const size_t total = sp.where.size();
        if (total <= config_.max_keys_per_scan.value_or(SIZE_MAX)) {
            sp.buckets.clear();  // everything goes: drop the buckets and their device mirrors wholesale
            sp.where.clear();
            return total;
        }
    }
    std::vector<std::string> keys = list_keys_bounded();
    for (auto &k : keys) {
        auto r = delete_in_space(*default_space_, k);
        if (r.is_err() && r.error().kind != ErrorKind::NotFound) return r.error();
    }
    return keys.size();
}

Result<size_t> VectorEngine::batch_delete_embeddings(const std::vector<std::string> &keys) {
    size_t deleted = 0;
    for (auto &k : keys)
        if (delete_in_space(*default_space_, k).is_ok()) ++deleted;  // missing keys are skipped
    return deleted;
}

namespace {
// lib.rs:2994-3020 / 3033-3058, shared by the two paginated searches
template <class Paged, class Pag>
Paged paginate_hits(std::vector<SearchResult> results, const Pag &pg) {
    Paged out;
    if (pg.count_total) out.total_count = results.size();
    for (size_t i = pg.skip; i < results.size(); ++i) {
        if (pg.limit && out.items.size() >= *pg.limit) break;
        out.items.push_back(std::move(results[i]));
    }
    out.has_more = false;
    if (pg.limit && out.total_count) {
        size_t end = pg.skip + out.items.size();
        if (end < pg.skip) end = SIZE_MAX;  // saturating_add
        out.has_more = end < *out.total_count;
    }
    return out;
}
size_t paginated_fetch(size_t top_k, size_t skip, const std::optional<size_t> &limit) {
    size_t need = skip + limit.value_or(top_k);
    if (need < skip) need = SIZE_MAX;  // saturating_add
    return std::min(need, top_k);
}
}  // namespace

Result<VectorEngine::PagedResult<SearchResult>> VectorEngine::search_similar_paginated(
    const std::vector<float> &query, size_t top_k, Pagination pagination) const {
    auto r = search_similar(query, paginated_fetch(top_k, pagination.skip, pagination.limit));
    if (r.is_err()) return r.error();
    return paginate_hits<PagedResult<SearchResult>>(std::move(r.value()), pagination);
}

Result<VectorEngine::PagedResult<SearchResult>> VectorEngine::search_entities_paginated(
    const std::vector<float> &query, size_t top_k, Pagination pagination) const {
    auto r = search_entities(query, paginated_fetch(top_k, pagination.skip, pagination.limit));
    if (r.is_err()) return r.error();
    return paginate_hits<PagedResult<SearchResult>>(std::move(r.value()), pagination);
}

bool VectorEngine::exists_in_collection(const std::string &collection, const std::string &key) const {
    std::shared_ptr<const Space> sp = find_collection_space(collection);
    if (!sp) return false;
    std::shared_lock<std::shared_mutex> g(sp->mu);
    return sp->where.count(key) != 0;
}

std::vector<std::string> VectorEngine::list_collection_keys(const std::string &collection) const {
    std::vector<std::string> out;
    std::shared_ptr<const Space> sp = find_collection_space(collection);
    if (!sp) return out;
    std::shared_lock<std::shared_mutex> g(sp->mu);
    for (auto &kv : sp->buckets)
        for (auto &k : kv.second->keys) out.push_back(k);
    return out;
}

std::optional<VectorCollectionConfig> VectorEngine::get_collection_config(const std::string &name) const {
    std::shared_lock<std::shared_mutex> g(collections_mu_);
    auto it = collections_.find(name);
    if (it == collections_.end()) return std::nullopt;
    return it->second->config;
}

Result<Unit> VectorEngine::store_in_collection_with_metadata(const std::string &collection,
                                                             const std::string &key,
                                                             std::vector<float> vector,
                                                             Metadata metadata) {
    if (vector.empty()) return err(ErrorKind::EmptyVector);
    {
        std::shared_lock<std::shared_mutex> g(collections_mu_);
        auto it = collections_.find(collection);
        if (it != collections_.end() && it->second->config && it->second->config->dimension &&
            vector.size() != *it->second->config->dimension)
            return dim_mismatch(*it->second->config->dimension, vector.size());
    }
    if (config_.max_dimension && vector.size() > *config_.max_dimension)
        return dim_mismatch(*config_.max_dimension, vector.size());
    std::shared_ptr<Space> sp = collection_space(collection);
    return store_in_space(*sp, key, std::move(vector), &metadata);
}

namespace {
// the reference samples the first min(100, n) keys of an (unordered) key listing
template <class SpaceT>
float sample_selectivity(const SpaceT &sp, const FilterCondition &f) {
    size_t seen = 0, hit = 0;
    for (auto &kv : sp.buckets)
        for (size_t r = 0; r < kv.second->keys.size() && seen < 100; ++r, ++seen)
            if (evaluate_filter(kv.second->meta[r], f)) ++hit;
    return seen ? (float)hit / (float)seen : -1.0f;
}
}  // namespace

Result<std::vector<SearchResult>> VectorEngine::filtered_in_space(
    const Space *sp, const std::vector<float> &query, size_t top_k, DistanceMetric post_metric,
    const FilterCondition &filter, const FilteredSearchConfig &cfg, const char *operation,
    std::chrono::steady_clock::time_point start) const {
    FilterStrategy strategy = cfg.strategy;
    if (strategy == FilterStrategy::Auto) {
        // choose_filter_strategy (lib.rs:3485-3511): True -> post; empty -> post; else sample
        strategy = FilterStrategy::PostFilter;
        if (filter.op != FilterCondition::Op::True && sp) {
            std::shared_lock<std::shared_mutex> g(sp->mu);
            float sel = sample_selectivity(*sp, filter);
            if (sel >= 0.0f && sel < cfg.selectivity_threshold) strategy = FilterStrategy::PreFilter;
        }
    }
    if (config_.search_timeout && std::chrono::steady_clock::now() - start >= *config_.search_timeout) {
        VectorError e;
        e.kind = ErrorKind::SearchTimeout;
        e.operation = operation;
        e.timeout_ms = (uint64_t)config_.search_timeout->count();
        return e;
    }
    if (!sp) return std::vector<SearchResult>{};
    if (strategy == FilterStrategy::PreFilter) {
        // always cosine, zero query -> empty (lib.rs:3520-3523, 1775-1790)
        if (simd::magnitude(query.data(), query.size()) == 0.0f) return std::vector<SearchResult>{};
        return scan_space(*sp, query, top_k, DistanceMetric::Cosine, operation, start, &filter);
    }
    // post-filter: oversample, then filter, then take top_k (lib.rs:3560-3578, 1792-1812)
    size_t oversample_k = top_k * cfg.oversample_factor;
    if (cfg.oversample_factor != 0 && oversample_k / cfg.oversample_factor != top_k)
        oversample_k = SIZE_MAX;  // saturating_mul
    oversample_k = std::max(oversample_k, top_k);
    if (post_metric == DistanceMetric::Cosine &&
        simd::magnitude(query.data(), query.size()) == 0.0f)
        return std::vector<SearchResult>{};
    auto cand = scan_space(*sp, query, oversample_k, post_metric, operation, start);
    if (cand.is_err()) return cand;
    std::vector<SearchResult> out;
    std::shared_lock<std::shared_mutex> g(sp->mu);
    for (auto &r : cand.value()) {
        auto it = sp->where.find(r.key);
        if (it == sp->where.end()) continue;
        if (evaluate_filter(sp->buckets.at(it->second.first)->meta[it->second.second], filter)) {
            out.push_back(r);
            if (out.size() == top_k) break;
        }
    }
    return out;
}

Result<std::vector<SearchResult>> VectorEngine::search_similar_filtered(
    const std::vector<float> &query, size_t top_k, const FilterCondition &filter,
    std::optional<FilteredSearchConfig> config) const {
    auto start = std::chrono::steady_clock::now();
    if (query.empty()) return err(ErrorKind::EmptyVector);
    if (top_k == 0) return err(ErrorKind::InvalidTopK);
    if (config_.max_dimension && query.size() > *config_.max_dimension)
        return dim_mismatch(*config_.max_dimension, query.size());
    return filtered_in_space(default_space_.get(), query, top_k, DistanceMetric::Cosine, filter,
                             config.value_or(FilteredSearchConfig{}), "search_similar_filtered",
                             start);
}

Result<std::vector<SearchResult>> VectorEngine::search_filtered_in_collection(
    const std::string &collection, const std::vector<float> &query, size_t top_k,
    const FilterCondition &filter, std::optional<FilteredSearchConfig> config) const {
    auto start = std::chrono::steady_clock::now();
    if (query.empty()) return err(ErrorKind::EmptyVector);
    if (top_k == 0) return err(ErrorKind::InvalidTopK);
    DistanceMetric metric = DistanceMetric::Cosine;
    std::shared_ptr<const Space> sp;  // keeps the collection alive across a concurrent delete
    {
        std::shared_lock<std::shared_mutex> g(collections_mu_);
        auto it = collections_.find(collection);
        if (it != collections_.end()) {
            if (it->second->config) {
                if (it->second->config->dimension &&
                    query.size() != *it->second->config->dimension)
                    return dim_mismatch(*it->second->config->dimension, query.size());
                metric = it->second->config->distance_metric;
            }
            sp = it->second->space;
        }
    }
    // zero query -> empty for every metric on this path (lib.rs:1729-1732)
    if (simd::magnitude(query.data(), query.size()) == 0.0f) return std::vector<SearchResult>{};
    return filtered_in_space(sp.get(), query, top_k, metric, filter,
                             config.value_or(FilteredSearchConfig{}),
                             "search_filtered_in_collection", start);
}

float VectorEngine::estimate_filter_selectivity(const FilterCondition &filter) const {
    std::shared_lock<std::shared_mutex> g(default_space_->mu);
    float s = sample_selectivity(*default_space_, filter);
    return s < 0.0f ? 0.0f : s;
}

size_t VectorEngine::count_matching(const FilterCondition &filter) const {
    return list_keys_matching(filter).size();
}

std::vector<std::string> VectorEngine::list_keys_matching(const FilterCondition &filter) const {
    std::vector<std::string> out;
    std::shared_lock<std::shared_mutex> g(default_space_->mu);
    for (auto &kv : default_space_->buckets)
        for (size_t r = 0; r < kv.second->keys.size(); ++r)
            if (evaluate_filter(kv.second->meta[r], filter)) out.push_back(kv.second->keys[r]);
    return out;
}

// ---- unified entity mode (lib.rs:3060-3219) ----
Result<Unit> VectorEngine::set_entity_embedding(const std::string &entity_key,
                                                std::vector<float> vector) {
    if (vector.empty()) return err(ErrorKind::EmptyVector);
    if (config_.max_dimension && vector.size() > *config_.max_dimension)
        return dim_mismatch(*config_.max_dimension, vector.size());
    return store_in_space(*entity_space_, entity_key, std::move(vector));
}

Result<std::vector<float>> VectorEngine::get_entity_embedding(const std::string &entity_key) const {
    return get_in_space(*entity_space_, entity_key);
}

bool VectorEngine::entity_has_embedding(const std::string &entity_key) const {
    std::shared_lock<std::shared_mutex> g(entity_space_->mu);
    return entity_space_->where.count(entity_key) != 0;
}

Result<Unit> VectorEngine::remove_entity_embedding(const std::string &entity_key) {
    return delete_in_space(*entity_space_, entity_key);
}

std::vector<std::string> VectorEngine::scan_entities_with_embeddings() const {
    const size_t limit = config_.max_keys_per_scan.value_or(SIZE_MAX);
    std::vector<std::string> out;
    std::shared_lock<std::shared_mutex> g(entity_space_->mu);
    for (auto &kv : entity_space_->buckets)
        for (auto &k : kv.second->keys) {
            if (out.size() >= limit) return out;
            out.push_back(k);
        }
    return out;
}

size_t VectorEngine::count_entities_with_embeddings() const {
    return scan_entities_with_embeddings().size();
}

Result<std::vector<SearchResult>> VectorEngine::search_entities(const std::vector<float> &query,
                                                                size_t top_k) const {
    auto start = std::chrono::steady_clock::now();
    if (query.empty()) return err(ErrorKind::EmptyVector);
    if (top_k == 0) return err(ErrorKind::InvalidTopK);
    if (config_.max_dimension && query.size() > *config_.max_dimension)
        return dim_mismatch(*config_.max_dimension, query.size());
    if (simd::magnitude(query.data(), query.size()) == 0.0f) return std::vector<SearchResult>{};
    return scan_space(*entity_space_, query, top_k, DistanceMetric::Cosine, "search_entities",
                      start);
}

// neumann_server/src/service/points.rs:449-485
Result<std::vector<VectorEngine::ScoredPoint>> VectorEngine::query_points(
    const std::string &collection, const std::vector<float> &vector, size_t limit, size_t offset,
    std::optional<float> score_threshold, bool with_vector) const {
    limit = std::max<size_t>(limit, 1);
    size_t want = limit + offset;
    if (want < limit) want = SIZE_MAX;  // saturating_add
    auto items = search_in_collection(collection, vector, want);
    if (items.is_err()) return items.error();
    std::vector<ScoredPoint> out;
    size_t taken = 0;
    for (size_t i = offset; i < items.value().size() && taken < limit; ++i, ++taken) {
        const SearchResult &it = items.value()[i];
        if (score_threshold && it.score < *score_threshold) continue;
        ScoredPoint p{it.key, it.score, {}};
        if (with_vector) {
            auto v = get_from_collection(collection, it.key);
            if (v.is_ok()) p.vector = v.value();
        }
        out.push_back(std::move(p));
    }
    return out;
}

std::vector<VectorEngine::MirrorInfo> VectorEngine::mirror_info() const {
    std::vector<MirrorInfo> out;
    std::shared_lock<std::shared_mutex> g(default_space_->mu);
    for (auto &kv : default_space_->buckets)
        out.push_back(MirrorInfo{kv.first, (uint64_t)kv.second->keys.size(),
                                 kv.second->mirror ? nm_index_rows(kv.second->mirror) : 0});
    return out;
}

// Host-only view of the device-side filter path (tests): the columns the rows of dimension `dim`
// of the default space would be pushed as, the postfix program `filter` compiles to, and what
// evaluate_filter says per row.  JSON, no device involved.
std::string VectorEngine::debug_filter_program(uint32_t dim, const FilterCondition &filter) const {
    Space &sp = *default_space_;
    std::unique_lock<std::shared_mutex> g(sp.mu);
    std::string out = "{";
    auto bit = sp.buckets.find(dim);
    if (bit == sp.buckets.end()) return "{\"rows\": 0}";
    Bucket &b = *bit->second;
    std::lock_guard<std::mutex> sg(b.sync_mu);
    const uint64_t n = b.keys.size();
    out += "\"rows\": " + std::to_string(n) + ", \"columns\": {";
    bool first_col = true;
    stage_columns(b, 0, n, true,
                  [&](uint32_t col, uint64_t row0, uint64_t m, const std::vector<uint8_t> &tags,
                      const std::vector<uint64_t> &vals) {
                      (void)row0;  // n <= one chunk in tests
                      out += std::string(first_col ? "" : ", ") + "\"" + std::to_string(col) + "\": {\"tags\": [";
                      first_col = false;
                      for (uint64_t i = 0; i < m; ++i) out += (i ? "," : "") + std::to_string(tags[i]);
                      out += "], \"vals\": [";
                      for (uint64_t i = 0; i < m; ++i) out += (i ? "," : "") + std::to_string(vals[i]);
                      out += "]}";
                      return 0;
                  });
    out += "}, \"ops\": [";
    FilterProgram prog;
    compile_filter(b, filter, prog);
    for (size_t i = 0; i < prog.ops.size(); ++i) {
        const nm_filter_op &op = prog.ops[i];
        out += std::string(i ? ", " : "") + "[" + std::to_string(op.kind) + "," + std::to_string(op.cmp) + "," +
               std::to_string(op.lit_tag) + "," + std::to_string(op.column) + "," + std::to_string(op.lit) +
               "," + std::to_string(op.table_off) + "," + std::to_string(op.table_bits) + "]";
    }
    out += "], \"tables\": [";
    for (size_t i = 0; i < prog.tables.size(); ++i) out += (i ? "," : "") + std::to_string(prog.tables[i]);
    out += "], \"host\": [";
    for (uint64_t r = 0; r < n; ++r) out += std::string(r ? "," : "") + (evaluate_filter(b.meta[r], filter) ? "1" : "0");
    out += "]}";
    return out;
}

}  // namespace neumann
