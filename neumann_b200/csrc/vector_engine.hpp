// vector_engine.hpp — C++ host mirror of the reference's `vector_engine` crate, restricted to
// the SIMILAR hot path and the store operations that feed it.  Same names, argument meaning
// and error behaviour as the Rust API (vector_engine/src/lib.rs); the scan itself is delegated
// to the device through the nm_* C ABI (include/neumann_b200.h).  There is no CPU scan here:
// without a CUDA device the search methods return StorageError.
//
// Not mirrored (out of scope): HNSW / IVF / PQ wrappers and persistence.
// tests/cpp/reference_suite.cpp replays the reference's own unit tests against this class.
#pragma once
#include <chrono>
#include <cstdint>
#include <map>
#include <memory>
#include <mutex>
#include <optional>
#include <shared_mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "filter.hpp"

struct nm_index;

namespace neumann {

// vector_engine/src/lib.rs:281-289
enum class DistanceMetric : int { Cosine = 0, Euclidean = 1, DotProduct = 2 };

// vector_engine/src/lib.rs:252-258
struct SearchResult {
    std::string key;
    float score;
};

// vector_engine/src/lib.rs:102-149 (the variants this path can produce)
enum class ErrorKind : int {
    NotFound,
    DimensionMismatch,
    EmptyVector,
    InvalidTopK,
    StorageError,
    ConfigurationError,
    CollectionExists,
    CollectionNotFound,
    SearchTimeout,
    InvalidArgument,
    BatchValidationError,  // { index, cause } — cause in `message`
    BatchOperationError,   // { index, cause }
};

struct VectorError {
    ErrorKind kind = ErrorKind::StorageError;
    std::string message;       // NotFound(key) / StorageError(msg) / ConfigurationError(msg) / ...
    size_t expected = 0, got = 0;  // DimensionMismatch
    size_t index = 0;          // BatchValidationError / BatchOperationError
    std::string operation;     // SearchTimeout
    uint64_t timeout_ms = 0;   // SearchTimeout
    int status() const;        // nm_status code
    std::string to_string() const;
};

template <class T>
class Result {
  public:
    Result(T v) : val_(std::move(v)) {}
    Result(VectorError e) : err_(std::move(e)) {}
    bool is_ok() const { return val_.has_value(); }
    bool is_err() const { return !val_.has_value(); }
    T &value() { return *val_; }
    const T &value() const { return *val_; }
    const VectorError &error() const { return *err_; }

  private:
    std::optional<T> val_;
    std::optional<VectorError> err_;
};
struct Unit {};

// vector_engine/src/lib.rs:626-663.  `devices` is the only addition: which GPUs hold the mirror.
struct VectorEngineConfig {
    std::optional<size_t> default_dimension;
    float sparse_threshold = 0.5f;
    size_t parallel_threshold = 5000;  // kept for API parity; the device scan has no such switch
    DistanceMetric default_metric = DistanceMetric::Cosine;
    std::optional<size_t> max_dimension;
    std::optional<size_t> max_keys_per_scan;
    size_t batch_parallel_threshold = 100;
    std::optional<std::chrono::milliseconds> search_timeout;
    std::vector<int> devices;  // empty = current device
    // Device-side addition.  false (default): nm_index_set_prefilter mode 2 — batches and
    // coalesced concurrent callers use the tensor-core pre-filter (the int8 copy is built by the
    // first eligible batch when it fits), single queries the f32 scan.  true: mode 1 — the copy is
    // built eagerly and single queries use the dp4a pre-filter too.  Results are bit-identical
    // either way; the copy costs +1 byte per element.
    bool device_prefilter = false;
    Result<Unit> validate() const;  // lib.rs:771-826
    // presets, lib.rs:666-700 (the index-file limits of the reference belong to persistence)
    static VectorEngineConfig high_throughput() {
        VectorEngineConfig c;
        c.parallel_threshold = 1000;
        return c;
    }
    VectorEngineConfig with_search_timeout(std::chrono::nanoseconds t) const {  // lib.rs:1000
        VectorEngineConfig c = *this;
        // milliseconds, like Deadline::timeout_ms: a sub-millisecond timeout (the 1 ns of the
        // reference's tests) is 0 ms = expired at the first check
        c.search_timeout = std::chrono::duration_cast<std::chrono::milliseconds>(t);
        return c;
    }
    static VectorEngineConfig low_memory() {
        VectorEngineConfig c;
        c.sparse_threshold = 0.3f;
        c.max_dimension = 4096;
        c.max_keys_per_scan = 10000;
        c.search_timeout = std::chrono::milliseconds(30000);
        return c;
    }
};

// vector_engine/src/lib.rs:455-499
struct VectorCollectionConfig {
    std::optional<size_t> dimension;
    DistanceMetric distance_metric = DistanceMetric::Cosine;
    bool auto_index = false;          // kept for API parity (HNSW is out of scope)
    size_t auto_index_threshold = 1000;
    VectorCollectionConfig with_dimension(size_t dim) const {
        VectorCollectionConfig c = *this;
        c.dimension = dim;
        return c;
    }
    VectorCollectionConfig with_metric(DistanceMetric m) const {
        VectorCollectionConfig c = *this;
        c.distance_metric = m;
        return c;
    }
    VectorCollectionConfig with_auto_index(size_t threshold) const {
        VectorCollectionConfig c = *this;
        c.auto_index = true;
        c.auto_index_threshold = threshold;
        return c;
    }
};

// The reference's f32 helpers this layer needs on the host (zero-query short-circuit,
// compute_similarity).  Bit-identical restatement of tensor_store/src/hnsw.rs:168-229.
namespace simd {
float dot_product(const float *a, const float *b, size_t n);
float sum_of_squares(const float *v, size_t n);
float magnitude(const float *v, size_t n);
}  // namespace simd

class VectorEngine {
  public:
    VectorEngine();
    explicit VectorEngine(VectorEngineConfig config);
    static Result<std::unique_ptr<VectorEngine>> with_config(VectorEngineConfig config);
    ~VectorEngine();
    VectorEngine(const VectorEngine &) = delete;
    VectorEngine &operator=(const VectorEngine &) = delete;

    const VectorEngineConfig &config() const { return config_; }

    // lib.rs:1840-1868 / 1896-1925 / 1929-1940
    Result<Unit> store_embedding(const std::string &key, std::vector<float> vector);
    Result<std::vector<float>> get_embedding(const std::string &key) const;
    Result<Unit> delete_embedding(const std::string &key);
    bool exists(const std::string &key) const;
    size_t count() const;
    std::optional<size_t> dimension() const;
    // lib.rs:2312-2354, 2924-2940: key listing, clear and batch delete (respecting
    // max_keys_per_scan like the reference); deletes maintain the device mirror by swap-remove.
    std::vector<std::string> list_keys() const;
    std::vector<std::string> list_keys_bounded() const;
    Result<size_t> clear();
    Result<size_t> batch_delete_embeddings(const std::vector<std::string> &keys);
    // lib.rs:1005-1046, 2858-2913: every input is validated first (BatchValidationError { index }),
    // then stored in order; a store that fails reports BatchOperationError { index, cause }.
    struct EmbeddingInput {
        std::string key;
        std::vector<float> vector;
    };
    struct BatchResult {
        std::vector<std::string> stored_keys;
        size_t count = 0;
    };
    Result<BatchResult> batch_store_embeddings(const std::vector<EmbeddingInput> &inputs);

    // lib.rs:1950-2037, 2049-2101
    Result<std::vector<SearchResult>> search_similar(const std::vector<float> &query,
                                                     size_t top_k) const;
    Result<std::vector<SearchResult>> search_similar_with_metric(const std::vector<float> &query,
                                                                 size_t top_k,
                                                                 DistanceMetric metric) const;
    // Batch form of search_similar_with_metric (no reference counterpart: there a batch is N
    // calls, e.g. the gRPC PointsService loop, neumann_server/src/service/points.rs:410-480).
    // Element i is exactly what search_similar_with_metric(queries[i], top_k, metric) returns;
    // an invalid query fails the whole call with the error that call would give.  Queries of
    // one dimension share device passes (batched kernels / tensor-core pre-filter).
    Result<std::vector<std::vector<SearchResult>>> search_similar_batch(
        const std::vector<std::vector<float>> &queries, size_t top_k, DistanceMetric metric) const;
    // metadata + filtered search: lib.rs:2930-3010 (store/get metadata), :3429-3557
    // (search_similar_filtered, pre/post filter), :1698-1829 (search_filtered_in_collection),
    // :3690-3735 (selectivity / count / list).  Pre-filter = device scan under a row bitmask.
    Result<Unit> store_embedding_with_metadata(const std::string &key, std::vector<float> vector,
                                               Metadata metadata);
    Result<Metadata> get_metadata(const std::string &key) const;
    // lib.rs:3346-3420: merge fields into / remove a field from an embedding's metadata.  The
    // device-side metadata columns follow at the next filtered search.
    Result<Unit> update_metadata(const std::string &key, const Metadata &metadata);
    Result<Unit> remove_metadata_field(const std::string &key, const std::string &field);
    bool has_metadata_field(const std::string &key, const std::string &field) const;
    Result<std::optional<MetadataValue>> get_metadata_field(const std::string &key,
                                                            const std::string &field) const;
    Result<std::vector<SearchResult>> search_similar_filtered(
        const std::vector<float> &query, size_t top_k, const FilterCondition &filter,
        std::optional<FilteredSearchConfig> config = std::nullopt) const;
    Result<Unit> store_in_collection_with_metadata(const std::string &collection,
                                                   const std::string &key, std::vector<float> vector,
                                                   Metadata metadata);
    Result<std::vector<SearchResult>> search_filtered_in_collection(
        const std::string &collection, const std::vector<float> &query, size_t top_k,
        const FilterCondition &filter, std::optional<FilteredSearchConfig> config = std::nullopt) const;
    float estimate_filter_selectivity(const FilterCondition &filter) const;
    size_t count_matching(const FilterCondition &filter) const;
    std::vector<std::string> list_keys_matching(const FilterCondition &filter) const;

    // unified entity mode: the `_embedding` field of arbitrary entity keys (lib.rs:3060-3219);
    // search_entities is what tensor_unified's find_similar_connected calls
    // (tensor_unified/src/lib.rs:913-914)
    Result<Unit> set_entity_embedding(const std::string &entity_key, std::vector<float> vector);
    Result<std::vector<float>> get_entity_embedding(const std::string &entity_key) const;
    bool entity_has_embedding(const std::string &entity_key) const;
    Result<Unit> remove_entity_embedding(const std::string &entity_key);
    // lib.rs:3224-3237 (at most max_keys_per_scan entities are visited, as there)
    std::vector<std::string> scan_entities_with_embeddings() const;
    size_t count_entities_with_embeddings() const;
    // lib.rs:1052-1121, 2988-3058: pagination over the ranked hits
    struct Pagination {
        size_t skip = 0;
        std::optional<size_t> limit;
        bool count_total = false;
        static Pagination with(size_t skip, size_t limit) {  // Pagination::new
            Pagination p;
            p.skip = skip;
            p.limit = limit;
            return p;
        }
        static Pagination skip_only(size_t skip) {
            Pagination p;
            p.skip = skip;
            return p;
        }
        Pagination with_total() const {
            Pagination p = *this;
            p.count_total = true;
            return p;
        }
    };
    template <class T>
    struct PagedResult {
        std::vector<T> items;
        std::optional<size_t> total_count;
        bool has_more = false;
        static PagedResult empty() {
            PagedResult r;
            r.total_count = 0;
            return r;
        }
    };
    PagedResult<std::string> list_keys_paginated(Pagination pagination) const;  // lib.rs:2946-2981
    Result<PagedResult<SearchResult>> search_similar_paginated(const std::vector<float> &query,
                                                              size_t top_k, Pagination pagination) const;
    Result<PagedResult<SearchResult>> search_entities_paginated(const std::vector<float> &query,
                                                               size_t top_k, Pagination pagination) const;
    Result<std::vector<SearchResult>> search_entities(const std::vector<float> &query,
                                                      size_t top_k) const;

    // gRPC PointsService::query post-processing (neumann_server/src/service/points.rs:449-485):
    // search limit+offset, skip offset, take limit, drop hits below score_threshold.
    struct ScoredPoint {
        std::string id;
        float score;
        std::vector<float> vector;  // filled when with_vector
    };
    Result<std::vector<ScoredPoint>> query_points(const std::string &collection,
                                                  const std::vector<float> &vector, size_t limit,
                                                  size_t offset, std::optional<float> score_threshold,
                                                  bool with_vector) const;

    // lib.rs:2277-2295
    static Result<float> compute_similarity(const std::vector<float> &a,
                                            const std::vector<float> &b);

    // collections: lib.rs:1371-1470, 1475-1560, 1585-1689
    Result<Unit> create_collection(const std::string &name, VectorCollectionConfig config);
    Result<Unit> delete_collection(const std::string &name);
    bool collection_exists(const std::string &name) const;
    std::vector<std::string> list_collections() const;
    Result<Unit> store_in_collection(const std::string &collection, const std::string &key,
                                     std::vector<float> vector);
    Result<std::vector<float>> get_from_collection(const std::string &collection,
                                                   const std::string &key) const;
    Result<Unit> delete_from_collection(const std::string &collection, const std::string &key);
    size_t collection_count(const std::string &collection) const;
    bool exists_in_collection(const std::string &collection, const std::string &key) const;   // lib.rs:1537
    std::vector<std::string> list_collection_keys(const std::string &collection) const;       // lib.rs:1543
    std::optional<VectorCollectionConfig> get_collection_config(const std::string &name) const;  // lib.rs:1412
    Result<Metadata> get_collection_metadata(const std::string &collection,
                                             const std::string &key) const;                   // lib.rs:1557
    Result<std::vector<SearchResult>> search_in_collection(const std::string &collection,
                                                           const std::vector<float> &query,
                                                           size_t top_k) const;

    // Device-mirror introspection (tests / metrics).
    struct MirrorInfo {
        uint32_t dim;
        uint64_t host_rows, device_rows;
    };
    std::vector<MirrorInfo> mirror_info() const;
    // Tests: columns + compiled postfix program + host evaluation for the rows of one dimension
    // of the default space, as JSON (no device involved).
    std::string debug_filter_program(uint32_t dim, const FilterCondition &filter) const;

  private:
    struct Bucket;
    struct Space;
    VectorEngineConfig config_;
    std::unique_ptr<Space> default_space_;
    std::unique_ptr<Space> entity_space_;
    mutable std::shared_mutex collections_mu_;
    // A collection's rows may exist without a config (store_in_collection does not require
    // create_collection, lib.rs:1445-1500); `config` is set by create_collection only.
    struct CollectionEntry;
    std::map<std::string, std::unique_ptr<CollectionEntry>> collections_;
    // Collection spaces are handed out as shared_ptr copies taken under collections_mu_: a
    // concurrent delete_collection only drops the map's reference, the last running operation
    // destroys the rows and the device mirrors (the reference's VectorEngine is Send + Sync and
    // its delete_collection is safe against concurrent searches, lib.rs:1127-1131).
    std::shared_ptr<Space> collection_space(const std::string &name);  // creates on demand
    std::shared_ptr<const Space> find_collection_space(const std::string &name) const;

    bool should_use_sparse(const std::vector<float> &v) const;  // lib.rs:1871-1886
  public:
    static bool should_use_sparse_with_threshold(const std::vector<float> &v, float threshold);  // lib.rs:1876
  private:
    Result<Unit> store_in_space(Space &sp, const std::string &key, std::vector<float> vector,
                                const Metadata *metadata = nullptr);
    Result<std::vector<SearchResult>> filtered_in_space(
        const Space *sp, const std::vector<float> &query, size_t top_k, DistanceMetric post_metric,
        const FilterCondition &filter, const FilteredSearchConfig &cfg, const char *operation,
        std::chrono::steady_clock::time_point start) const;
    Result<Unit> delete_in_space(Space &sp, const std::string &key);
    Result<std::vector<float>> get_in_space(const Space &sp, const std::string &key) const;
    Result<std::vector<SearchResult>> scan_space(const Space &sp, const std::vector<float> &query,
                                                 size_t top_k, DistanceMetric metric,
                                                 const char *operation,
                                                 std::chrono::steady_clock::time_point start,
                                                 const FilterCondition *pre_filter = nullptr) const;
    // nq queries of one dimension (row-major) over one space: one nm_search call
    Result<std::vector<std::vector<SearchResult>>> scan_space_batch(
        const Space &sp, const float *queries, size_t nq, size_t dim, size_t top_k,
        DistanceMetric metric, const char *operation,
        std::chrono::steady_clock::time_point start) const;
};

}  // namespace neumann
