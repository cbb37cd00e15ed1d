// reference_suite.cpp — the reference's OWN unit tests for the SIMILAR path, replayed one by one
// against the C++ host mirror (neumann_b200/csrc/vector_engine.hpp) and, through it, the C ABI
// and the CUDA kernels.  Every TEST carries the name of the Rust test it restates and its
// location in vector_engine/src/lib.rs (Shadylukin/Neumann @ aae3d465); the assertions are the
// reference's assertions.  Tests of subsystems outside SURVEY §8 (HNSW / IVF / PQ wrappers,
// persistence, TensorStore plumbing) are not here.
//
// The last sections do the same for the SIMILAR / EMBED tests of query_router/src/lib.rs against
// neumann::QueryRouter (similar_router.hpp) and for the PointsService::query cases of
// neumann_server/tests/grpc_vector_points.rs against VectorEngine::query_points.
//
// Test infrastructure: built and run by tests/test_cpp_reference_suite.py
//   g++ -std=c++17 tests/cpp/reference_suite.cpp -Ineumann_b200/csrc -Iinclude -Lneumann_b200 -lneumann_b200
//   ./reference_suite            every test (needs a CUDA device: searches run on the GPU)
//   ./reference_suite --host     only the tests that never search (no device needed)
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <functional>
#include <string>
#include <thread>
#include <vector>

#include "similar_router.hpp"
#include "vector_engine.hpp"

using namespace neumann;

namespace {

struct TestCase {
    const char *name;
    const char *src;
    bool needs_device;
    std::function<void()> fn;
};
std::vector<TestCase> &registry() {
    static std::vector<TestCase> r;
    return r;
}
struct Registrar {
    Registrar(const char *name, const char *src, bool dev, std::function<void()> fn) {
        registry().push_back(TestCase{name, src, dev, std::move(fn)});
    }
};
int g_failures = 0;
const char *g_current = "";

#define TEST_IMPL(name, src, dev)                                   \
    static void test_##name();                                      \
    static Registrar reg_##name(#name, src, dev, test_##name);      \
    static void test_##name()
#define TEST_HOST(name, src) TEST_IMPL(name, src, false)   // never reaches the device
#define TEST_GPU(name, src) TEST_IMPL(name, src, true)     // runs a scan on the device

#define REQUIRE(cond)                                                                        \
    do {                                                                                     \
        if (!(cond)) {                                                                       \
            std::fprintf(stderr, "FAIL %s (%s:%d): %s\n", g_current, __FILE__, __LINE__, #cond); \
            ++g_failures;                                                                    \
            return;                                                                          \
        }                                                                                    \
    } while (0)
#define REQUIRE_OK(res) REQUIRE((res).is_ok())
#define REQUIRE_ERR(res, k) REQUIRE((res).is_err() && (res).error().kind == ErrorKind::k)

using Vec = std::vector<float>;

// lib.rs:4029-4038
Vec create_test_vector(size_t dim, size_t seed) {
    Vec v(dim);
    for (size_t i = 0; i < dim; ++i) {
        const float x = (float)(seed * 31 + i * 17);
        v[i] = std::sin(x * 0.0001f) * ((float)(seed + i) * 0.001f);
    }
    return v;
}
// lib.rs:4040-4047
Vec normalize(const Vec &v) {
    float s = 0.0f;
    for (float x : v) s += x * x;
    const float mag = std::sqrt(s);
    if (mag == 0.0f) return v;
    Vec o(v);
    for (float &x : o) x /= mag;
    return o;
}
std::string key_of(const char *prefix, size_t i) { return std::string(prefix) + std::to_string(i); }
bool has_key(const std::vector<SearchResult> &r, const std::string &k) {
    return std::any_of(r.begin(), r.end(), [&](const SearchResult &x) { return x.key == k; });
}
Metadata meta(std::initializer_list<std::pair<const std::string, MetadataValue>> l) { return Metadata(l); }
MetadataValue S(const char *s) { return MetadataValue::string(s); }
MetadataValue I(int64_t v) { return MetadataValue::integer(v); }
MetadataValue F(double v) { return MetadataValue::real(v); }
MetadataValue B(bool v) { return MetadataValue::boolean(v); }
using Op = FilterCondition::Op;

// ---------------------------------------------------------------------------------------------
// basic store / get / delete (lib.rs:4050-4118, 4355-4399, 4475-4483)
// ---------------------------------------------------------------------------------------------
TEST_HOST(store_and_retrieve_embedding, "lib.rs:4050") {
    VectorEngine engine;
    Vec vector{1.0f, 2.0f, 3.0f};
    REQUIRE_OK(engine.store_embedding("test", vector));
    auto got = engine.get_embedding("test");
    REQUIRE_OK(got);
    REQUIRE(got.value() == vector);
}
TEST_HOST(store_overwrites_existing, "lib.rs:4061") {
    VectorEngine engine;
    REQUIRE_OK(engine.store_embedding("key", {1.0f, 2.0f}));
    REQUIRE_OK(engine.store_embedding("key", {3.0f, 4.0f}));
    REQUIRE(engine.get_embedding("key").value() == (Vec{3.0f, 4.0f}));
}
TEST_HOST(delete_embedding, "lib.rs:4072") {
    VectorEngine engine;
    REQUIRE_OK(engine.store_embedding("key", {1.0f, 2.0f}));
    REQUIRE(engine.exists("key"));
    REQUIRE_OK(engine.delete_embedding("key"));
    REQUIRE(!engine.exists("key"));
}
TEST_HOST(delete_nonexistent_returns_error, "lib.rs:4083") {
    VectorEngine engine;
    REQUIRE_ERR(engine.delete_embedding("nonexistent"), NotFound);
}
TEST_HOST(get_nonexistent_returns_error, "lib.rs:4091") {
    VectorEngine engine;
    REQUIRE_ERR(engine.get_embedding("nonexistent"), NotFound);
}
TEST_HOST(empty_vector_returns_error, "lib.rs:4099") {
    VectorEngine engine;
    REQUIRE_ERR(engine.store_embedding("key", {}), EmptyVector);
}
TEST_HOST(count_embeddings, "lib.rs:4107") {
    VectorEngine engine;
    REQUIRE(engine.count() == 0);
    REQUIRE_OK(engine.store_embedding("a", {1.0f}));
    REQUIRE_OK(engine.store_embedding("b", {2.0f}));
    REQUIRE_OK(engine.store_embedding("c", {3.0f}));
    REQUIRE(engine.count() == 3);
}
TEST_HOST(list_keys, "lib.rs:4355") {
    VectorEngine engine;
    REQUIRE_OK(engine.store_embedding("alpha", {1.0f}));
    REQUIRE_OK(engine.store_embedding("beta", {2.0f}));
    REQUIRE_OK(engine.store_embedding("gamma", {3.0f}));
    auto keys = engine.list_keys();
    std::sort(keys.begin(), keys.end());
    REQUIRE(keys == (std::vector<std::string>{"alpha", "beta", "gamma"}));
}
TEST_HOST(clear_all, "lib.rs:4369") {
    VectorEngine engine;
    REQUIRE_OK(engine.store_embedding("a", {1.0f}));
    REQUIRE_OK(engine.store_embedding("b", {2.0f}));
    REQUIRE(engine.count() == 2);
    auto cleared = engine.clear();
    REQUIRE_OK(cleared);
    REQUIRE(cleared.value() == 2);
    REQUIRE(engine.count() == 0);
}
TEST_HOST(dimension, "lib.rs:4384") {
    VectorEngine engine;
    REQUIRE(!engine.dimension().has_value());
    REQUIRE_OK(engine.store_embedding("test", {1.0f, 2.0f, 3.0f}));
    REQUIRE(engine.dimension() == std::optional<size_t>(3));
}
TEST_HOST(default_trait, "lib.rs:4395") {
    VectorEngine engine;
    REQUIRE(engine.count() == 0);
}
TEST_HOST(error_display, "lib.rs:4408") {
    VectorError e;
    e.kind = ErrorKind::NotFound;
    e.message = "test";
    REQUIRE(e.to_string() == "Embedding not found: test");
    e = VectorError{};
    e.kind = ErrorKind::DimensionMismatch;
    e.expected = 3;
    e.got = 5;
    REQUIRE(e.to_string() == "Dimension mismatch: expected 3, got 5");
    e = VectorError{};
    e.kind = ErrorKind::EmptyVector;
    REQUIRE(e.to_string() == "Empty vector provided");
    e.kind = ErrorKind::InvalidTopK;
    REQUIRE(e.to_string() == "Invalid top_k value (must be > 0)");
    e.kind = ErrorKind::StorageError;
    e.message = "test";
    REQUIRE(e.to_string() == "Storage error: test");
}
TEST_HOST(exists_check, "lib.rs:4475") {
    VectorEngine engine;
    REQUIRE(!engine.exists("key"));
    REQUIRE_OK(engine.store_embedding("key", {1.0f}));
    REQUIRE(engine.exists("key"));
}

// ---------------------------------------------------------------------------------------------
// search_similar + compute_similarity (lib.rs:4120-4352, 4458-4473)
// ---------------------------------------------------------------------------------------------
TEST_GPU(search_similar_basic, "lib.rs:4120") {
    VectorEngine engine;
    REQUIRE_OK(engine.store_embedding("a", {1.0f, 0.0f, 0.0f}));
    REQUIRE_OK(engine.store_embedding("b", {0.0f, 1.0f, 0.0f}));
    REQUIRE_OK(engine.store_embedding("c", {1.0f, 1.0f, 0.0f}));
    auto results = engine.search_similar({1.0f, 0.0f, 0.0f}, 3);
    REQUIRE_OK(results);
    REQUIRE(results.value().size() == 3);
    REQUIRE(results.value()[0].key == "a");
    REQUIRE(std::fabs(results.value()[0].score - 1.0f) < 1e-6f);
}
TEST_GPU(search_similar_top_k, "lib.rs:4138") {
    VectorEngine engine;
    for (int i = 0; i < 10; ++i) REQUIRE_OK(engine.store_embedding(key_of("v", i), {(float)i, 0.0f}));
    auto results = engine.search_similar({5.0f, 0.0f}, 3);
    REQUIRE_OK(results);
    REQUIRE(results.value().size() == 3);
}
TEST_GPU(search_similar_fewer_than_k, "lib.rs:4154") {
    VectorEngine engine;
    REQUIRE_OK(engine.store_embedding("a", {1.0f, 0.0f}));
    REQUIRE_OK(engine.store_embedding("b", {0.0f, 1.0f}));
    auto results = engine.search_similar({1.0f, 0.0f}, 10);
    REQUIRE_OK(results);
    REQUIRE(results.value().size() == 2);
}
TEST_HOST(search_similar_empty_query_error, "lib.rs:4167") {
    VectorEngine engine;
    REQUIRE_ERR(engine.search_similar({}, 5), EmptyVector);
}
TEST_HOST(search_similar_zero_top_k_error, "lib.rs:4175") {
    VectorEngine engine;
    REQUIRE_ERR(engine.search_similar({1.0f, 0.0f}, 0), InvalidTopK);
}
TEST_HOST(cosine_similarity_identical_vectors, "lib.rs:4183") {
    Vec v{1.0f, 2.0f, 3.0f};
    REQUIRE(std::fabs(VectorEngine::compute_similarity(v, v).value() - 1.0f) < 1e-6f);
}
TEST_HOST(cosine_similarity_orthogonal_vectors, "lib.rs:4190") {
    REQUIRE(std::fabs(VectorEngine::compute_similarity({1.0f, 0.0f}, {0.0f, 1.0f}).value()) < 1e-6f);
}
TEST_HOST(cosine_similarity_opposite_vectors, "lib.rs:4198") {
    REQUIRE(std::fabs(VectorEngine::compute_similarity({1.0f, 0.0f}, {-1.0f, 0.0f}).value() + 1.0f) < 1e-6f);
}
TEST_HOST(cosine_similarity_normalized_vectors, "lib.rs:4206") {
    const float score = VectorEngine::compute_similarity(normalize({1.0f, 0.0f}), normalize({1.0f, 1.0f})).value();
    REQUIRE(std::fabs(score - std::sqrt(2.0f) / 2.0f) < 1e-6f);
}
TEST_HOST(cosine_similarity_dimension_mismatch, "lib.rs:4216") {
    REQUIRE_ERR(VectorEngine::compute_similarity({1.0f, 2.0f}, {1.0f, 2.0f, 3.0f}), DimensionMismatch);
}
TEST_HOST(cosine_similarity_zero_vector, "lib.rs:4224") {
    REQUIRE(VectorEngine::compute_similarity({0.0f, 0.0f}, {1.0f, 0.0f}).value() == 0.0f);
}
TEST_HOST(cosine_similarity_both_zero_vectors, "lib.rs:4232") {
    const float score = VectorEngine::compute_similarity({0.0f, 0.0f}, {0.0f, 0.0f}).value();
    REQUIRE(score == 0.0f);
    REQUIRE(!std::isnan(score));
}
TEST_GPU(search_skips_dimension_mismatch, "lib.rs:4242") {
    VectorEngine engine;
    REQUIRE_OK(engine.store_embedding("2d", {1.0f, 0.0f}));
    REQUIRE_OK(engine.store_embedding("3d", {1.0f, 0.0f, 0.0f}));
    auto results = engine.search_similar({1.0f, 0.0f}, 10);
    REQUIRE_OK(results);
    REQUIRE(results.value().size() == 1);
    REQUIRE(results.value()[0].key == "2d");
}
TEST_GPU(store_10000_vectors_search, "lib.rs:4256") {
    VectorEngine engine;
    const size_t dim = 128;
    for (size_t i = 0; i < 10000; ++i) REQUIRE_OK(engine.store_embedding(key_of("v", i), create_test_vector(dim, i)));
    REQUIRE(engine.count() == 10000);
    auto results = engine.search_similar(create_test_vector(dim, 5000), 5);
    REQUIRE_OK(results);
    REQUIRE(results.value().size() == 5);
    REQUIRE(results.value()[0].key == "v5000");
    REQUIRE(std::fabs(results.value()[0].score - 1.0f) < 1e-5f);
}
TEST_GPU(high_dimensional_768, "lib.rs:4279") {
    VectorEngine engine;
    for (size_t i = 0; i < 100; ++i) REQUIRE_OK(engine.store_embedding(key_of("v", i), create_test_vector(768, i)));
    auto results = engine.search_similar(create_test_vector(768, 50), 3);
    REQUIRE_OK(results);
    REQUIRE(results.value().size() == 3);
    REQUIRE(results.value()[0].key == "v50");
}
TEST_GPU(high_dimensional_1536, "lib.rs:4297") {
    VectorEngine engine;
    for (size_t i = 0; i < 100; ++i) REQUIRE_OK(engine.store_embedding(key_of("v", i), create_test_vector(1536, i)));
    auto results = engine.search_similar(create_test_vector(1536, 75), 5);
    REQUIRE_OK(results);
    REQUIRE(results.value().size() == 5);
    REQUIRE(results.value()[0].key == "v75");
}
TEST_GPU(similarity_scores_mathematically_correct, "lib.rs:4315") {
    VectorEngine engine;
    REQUIRE_OK(engine.store_embedding("unit_x", normalize({1.0f, 0.0f, 0.0f})));
    REQUIRE_OK(engine.store_embedding("unit_y", normalize({0.0f, 1.0f, 0.0f})));
    REQUIRE_OK(engine.store_embedding("unit_z", normalize({0.0f, 0.0f, 1.0f})));
    REQUIRE_OK(engine.store_embedding("diag_xy", normalize({1.0f, 1.0f, 0.0f})));
    REQUIRE_OK(engine.store_embedding("neg_x", normalize({-1.0f, 0.0f, 0.0f})));
    auto results = engine.search_similar(normalize({1.0f, 0.0f, 0.0f}), 5);
    REQUIRE_OK(results);
    REQUIRE(results.value().size() == 5);
    for (const auto &r : results.value()) {
        if (r.key == "unit_x") REQUIRE(std::fabs(r.score - 1.0f) < 1e-6f);
        else if (r.key == "unit_y" || r.key == "unit_z") REQUIRE(std::fabs(r.score) < 1e-6f);
        else if (r.key == "diag_xy") REQUIRE(std::fabs(r.score - std::sqrt(2.0f) / 2.0f) < 1e-6f);
        else if (r.key == "neg_x") REQUIRE(std::fabs(r.score + 1.0f) < 1e-6f);
        else REQUIRE(!"unexpected key");
    }
}
TEST_HOST(search_zero_query_vector, "lib.rs:4458") {
    VectorEngine engine;
    REQUIRE_OK(engine.store_embedding("a", {1.0f, 0.0f}));
    auto results = engine.search_similar({0.0f, 0.0f}, 5);  // short-circuits before the scan
    REQUIRE_OK(results);
    REQUIRE(results.value().empty());
}
TEST_HOST(search_no_embeddings, "lib.rs:4468") {
    VectorEngine engine;
    auto results = engine.search_similar({1.0f, 0.0f}, 5);
    REQUIRE_OK(results);
    REQUIRE(results.value().empty());
}

// ---------------------------------------------------------------------------------------------
// unified entity embeddings (lib.rs:4692-4865)
// ---------------------------------------------------------------------------------------------
TEST_HOST(entity_embedding_store_and_retrieve, "lib.rs:4692") {
    VectorEngine engine;
    Vec vector{1.0f, 2.0f, 3.0f};
    REQUIRE_OK(engine.set_entity_embedding("user:1", vector));
    REQUIRE(engine.get_entity_embedding("user:1").value() == vector);
}
TEST_HOST(entity_has_embedding_check, "lib.rs:4726") {
    VectorEngine engine;
    REQUIRE(!engine.entity_has_embedding("user:1"));
    REQUIRE_OK(engine.set_entity_embedding("user:1", {1.0f, 2.0f}));
    REQUIRE(engine.entity_has_embedding("user:1"));
}
TEST_HOST(entity_embedding_remove, "lib.rs:4738") {
    VectorEngine engine;
    REQUIRE_OK(engine.set_entity_embedding("user:1", {1.0f, 2.0f}));
    REQUIRE(engine.entity_has_embedding("user:1"));
    REQUIRE_OK(engine.remove_entity_embedding("user:1"));
    REQUIRE(!engine.entity_has_embedding("user:1"));
}
TEST_HOST(entity_embedding_remove_nonexistent_error, "lib.rs:4751") {
    VectorEngine engine;
    REQUIRE_ERR(engine.remove_entity_embedding("user:999"), NotFound);
}
TEST_HOST(entity_embedding_get_nonexistent_error, "lib.rs:4758") {
    VectorEngine engine;
    REQUIRE_ERR(engine.get_entity_embedding("user:999"), NotFound);
}
TEST_HOST(entity_embedding_empty_vector_error, "lib.rs:4765") {
    VectorEngine engine;
    REQUIRE_ERR(engine.set_entity_embedding("user:1", {}), EmptyVector);
}
TEST_GPU(search_entities_basic, "lib.rs:4772") {
    VectorEngine engine;
    REQUIRE_OK(engine.set_entity_embedding("user:1", {1.0f, 0.0f, 0.0f}));
    REQUIRE_OK(engine.set_entity_embedding("user:2", {0.0f, 1.0f, 0.0f}));
    REQUIRE_OK(engine.set_entity_embedding("user:3", {1.0f, 1.0f, 0.0f}));
    auto results = engine.search_entities({1.0f, 0.0f, 0.0f}, 3);
    REQUIRE_OK(results);
    REQUIRE(results.value().size() == 3);
    REQUIRE(results.value()[0].key == "user:1");
    REQUIRE(std::fabs(results.value()[0].score - 1.0f) < 1e-6f);
}
TEST_GPU(search_entities_filters_non_embeddings, "lib.rs:4793") {
    // the reference puts a second entity without an `_embedding` field into the TensorStore; here
    // an entity without an embedding is simply absent from the entity space — and an embedding
    // stored under the `emb:` namespace must not come back from search_entities either
    VectorEngine engine;
    REQUIRE_OK(engine.set_entity_embedding("user:1", {1.0f, 0.0f}));
    REQUIRE_OK(engine.store_embedding("user:2", {1.0f, 0.0f}));
    auto results = engine.search_entities({1.0f, 0.0f}, 10);
    REQUIRE_OK(results);
    REQUIRE(results.value().size() == 1);
    REQUIRE(results.value()[0].key == "user:1");
}
TEST_HOST(scan_entities_with_embeddings, "lib.rs:4815") {
    VectorEngine engine;
    REQUIRE_OK(engine.set_entity_embedding("user:1", {1.0f, 2.0f}));
    REQUIRE_OK(engine.set_entity_embedding("user:2", {3.0f, 4.0f}));
    REQUIRE(engine.scan_entities_with_embeddings().size() == 2);
}
TEST_HOST(count_entities_with_embeddings, "lib.rs:4830") {
    VectorEngine engine;
    REQUIRE(engine.count_entities_with_embeddings() == 0);
    REQUIRE_OK(engine.set_entity_embedding("user:1", {1.0f}));
    REQUIRE_OK(engine.set_entity_embedding("user:2", {2.0f}));
    REQUIRE(engine.count_entities_with_embeddings() == 2);
}
TEST_HOST(search_entities_empty_query_error, "lib.rs:4842") {
    VectorEngine engine;
    REQUIRE_ERR(engine.search_entities({}, 5), EmptyVector);
}
TEST_HOST(search_entities_zero_top_k_error, "lib.rs:4849") {
    VectorEngine engine;
    REQUIRE_ERR(engine.search_entities({1.0f}, 0), InvalidTopK);
}
TEST_HOST(search_entities_zero_query_returns_empty, "lib.rs:4856") {
    VectorEngine engine;
    REQUIRE_OK(engine.set_entity_embedding("user:1", {1.0f, 0.0f}));
    auto results = engine.search_entities({0.0f, 0.0f}, 5);
    REQUIRE_OK(results);
    REQUIRE(results.value().empty());
}

// ---------------------------------------------------------------------------------------------
// metrics (lib.rs:4867-4997)
// ---------------------------------------------------------------------------------------------
TEST_HOST(distance_metric_default, "lib.rs:4867") {
    REQUIRE(VectorEngineConfig{}.default_metric == DistanceMetric::Cosine);
    REQUIRE(VectorCollectionConfig{}.distance_metric == DistanceMetric::Cosine);
}
TEST_GPU(search_with_metric_cosine, "lib.rs:4872") {
    VectorEngine engine;
    REQUIRE_OK(engine.store_embedding("a", {1.0f, 0.0f}));
    REQUIRE_OK(engine.store_embedding("b", {0.707f, 0.707f}));
    REQUIRE_OK(engine.store_embedding("c", {0.0f, 1.0f}));
    auto results = engine.search_similar_with_metric({1.0f, 0.0f}, 3, DistanceMetric::Cosine);
    REQUIRE_OK(results);
    REQUIRE(results.value().size() == 3);
    REQUIRE(results.value()[0].key == "a");
    REQUIRE(std::fabs(results.value()[0].score - 1.0f) < 0.01f);
}
TEST_GPU(search_with_metric_dot_product, "lib.rs:4889") {
    VectorEngine engine;
    REQUIRE_OK(engine.store_embedding("a", {1.0f, 0.0f}));
    REQUIRE_OK(engine.store_embedding("b", {2.0f, 0.0f}));
    REQUIRE_OK(engine.store_embedding("c", {0.5f, 0.0f}));
    auto results = engine.search_similar_with_metric({1.0f, 0.0f}, 3, DistanceMetric::DotProduct);
    REQUIRE_OK(results);
    REQUIRE(results.value().size() == 3);
    REQUIRE(results.value()[0].key == "b");
    REQUIRE(std::fabs(results.value()[0].score - 2.0f) < 0.01f);
}
TEST_GPU(search_with_metric_euclidean, "lib.rs:4906") {
    VectorEngine engine;
    REQUIRE_OK(engine.store_embedding("a", {1.0f, 0.0f}));
    REQUIRE_OK(engine.store_embedding("b", {2.0f, 0.0f}));
    REQUIRE_OK(engine.store_embedding("c", {10.0f, 0.0f}));
    auto results = engine.search_similar_with_metric({1.0f, 0.0f}, 3, DistanceMetric::Euclidean);
    REQUIRE_OK(results);
    REQUIRE(results.value().size() == 3);
    REQUIRE(results.value()[0].key == "a");
    REQUIRE(std::fabs(results.value()[0].score - 1.0f) < 0.01f);
    REQUIRE(results.value()[1].key == "b");
    REQUIRE(std::fabs(results.value()[1].score - 0.5f) < 0.01f);
}
TEST_HOST(search_with_metric_empty_query_error, "lib.rs:4927") {
    VectorEngine engine;
    REQUIRE_ERR(engine.search_similar_with_metric({}, 5, DistanceMetric::Cosine), EmptyVector);
}
TEST_HOST(search_with_metric_zero_top_k_error, "lib.rs:4934") {
    VectorEngine engine;
    REQUIRE_ERR(engine.search_similar_with_metric({1.0f}, 0, DistanceMetric::Cosine), InvalidTopK);
}
TEST_HOST(search_with_metric_zero_query, "lib.rs:4941") {
    VectorEngine engine;
    REQUIRE_OK(engine.store_embedding("a", {1.0f, 0.0f}));
    auto results = engine.search_similar_with_metric({0.0f, 0.0f}, 5, DistanceMetric::Cosine);
    REQUIRE_OK(results);
    REQUIRE(results.value().empty());
}
TEST_GPU(search_with_metric_zero_query_euclidean, "lib.rs:4952") {
    VectorEngine engine;
    REQUIRE_OK(engine.store_embedding("origin", {0.0f, 0.0f}));
    REQUIRE_OK(engine.store_embedding("unit", {1.0f, 0.0f}));
    REQUIRE_OK(engine.store_embedding("far", {10.0f, 0.0f}));
    auto results = engine.search_similar_with_metric({0.0f, 0.0f}, 3, DistanceMetric::Euclidean);
    REQUIRE_OK(results);
    REQUIRE(results.value().size() == 3);
    REQUIRE(results.value()[0].key == "origin");
    REQUIRE(std::fabs(results.value()[0].score - 1.0f) < 0.01f);
    REQUIRE(results.value()[1].key == "unit");
    REQUIRE(std::fabs(results.value()[1].score - 0.5f) < 0.01f);
}
// euclidean_distance is private in the reference; its value is observable as d = 1/score - 1
static float euclid_via_search(const Vec &a, const Vec &b, bool *ok) {
    VectorEngine engine;
    *ok = engine.store_embedding("b", b).is_ok();
    auto r = engine.search_similar_with_metric(a, 1, DistanceMetric::Euclidean);
    *ok = *ok && r.is_ok() && r.value().size() == 1;
    return *ok ? 1.0f / r.value()[0].score - 1.0f : NAN;
}
TEST_GPU(euclidean_distance_identical, "lib.rs:4973") {
    bool ok;
    const float d = euclid_via_search({1.0f, 2.0f, 3.0f}, {1.0f, 2.0f, 3.0f}, &ok);
    REQUIRE(ok && std::fabs(d) < 1e-6f);
}
TEST_GPU(euclidean_distance_unit, "lib.rs:4980") {
    bool ok;
    const float d = euclid_via_search({0.0f, 0.0f}, {1.0f, 0.0f}, &ok);
    REQUIRE(ok && std::fabs(d - 1.0f) < 1e-6f);
}
TEST_GPU(euclidean_distance_pythagoras, "lib.rs:4988") {
    bool ok;
    const float d = euclid_via_search({0.0f, 0.0f}, {3.0f, 4.0f}, &ok);
    REQUIRE(ok && std::fabs(d - 5.0f) < 1e-5f);
}

// ---------------------------------------------------------------------------------------------
// sparse storage (lib.rs:4999-5145, 5759)
// ---------------------------------------------------------------------------------------------
TEST_HOST(sparse_vector_storage_and_retrieval, "lib.rs:4999") {
    VectorEngine engine;
    Vec sparse(100, 0.0f);
    sparse[0] = 1.0f;
    sparse[50] = 2.0f;
    sparse[99] = 3.0f;
    REQUIRE_OK(engine.store_embedding("sparse", sparse));
    auto got = engine.get_embedding("sparse");
    REQUIRE_OK(got);
    REQUIRE(got.value().size() == sparse.size());
    REQUIRE(got.value()[0] == 1.0f && got.value()[50] == 2.0f && got.value()[99] == 3.0f);
}
TEST_GPU(sparse_vector_search, "lib.rs:5018") {
    VectorEngine engine;
    Vec v1(100, 0.0f), v2(100, 0.0f), v3(100, 0.0f), query(100, 0.0f);
    v1[0] = 1.0f;
    v2[0] = 0.707f;
    v2[1] = 0.707f;
    v3[1] = 1.0f;
    query[0] = 1.0f;
    REQUIRE_OK(engine.store_embedding("v1", v1));
    REQUIRE_OK(engine.store_embedding("v2", v2));
    REQUIRE_OK(engine.store_embedding("v3", v3));
    auto results = engine.search_similar(query, 3);
    REQUIRE_OK(results);
    REQUIRE(results.value().size() == 3);
    REQUIRE(results.value()[0].key == "v1");
    REQUIRE(std::fabs(results.value()[0].score - 1.0f) < 0.01f);
}
TEST_HOST(sparse_entity_embedding, "lib.rs:5052") {
    VectorEngine engine;
    Vec sparse(100, 0.0f);
    sparse[10] = 5.0f;
    sparse[20] = -3.0f;
    REQUIRE_OK(engine.set_entity_embedding("entity:1", sparse));
    auto got = engine.get_entity_embedding("entity:1");
    REQUIRE_OK(got);
    REQUIRE(got.value().size() == 100);
    REQUIRE(got.value()[10] == 5.0f && got.value()[20] == -3.0f && got.value()[0] == 0.0f);
}
TEST_HOST(sparse_detection_threshold, "lib.rs:5071") {
    Vec half(100), dense(100), very(100);
    for (int i = 0; i < 100; ++i) {
        half[i] = i < 50 ? 0.0f : 1.0f;
        dense[i] = i < 40 ? 0.0f : 1.0f;
        very[i] = i < 3 ? 1.0f : 0.0f;
    }
    REQUIRE(VectorEngine::should_use_sparse_with_threshold(half, 0.5f));
    REQUIRE(!VectorEngine::should_use_sparse_with_threshold(dense, 0.5f));
    REQUIRE(VectorEngine::should_use_sparse_with_threshold(very, 0.5f));
}
TEST_GPU(sparse_search_with_metric, "lib.rs:5095") {
    VectorEngine engine;
    Vec v1(100, 0.0f), v2(100, 0.0f), query(100, 0.0f);
    v1[0] = 1.0f;
    v2[0] = 2.0f;
    query[0] = 1.0f;
    REQUIRE_OK(engine.store_embedding("v1", v1));
    REQUIRE_OK(engine.store_embedding("v2", v2));
    auto results = engine.search_similar_with_metric(query, 2, DistanceMetric::Euclidean);
    REQUIRE_OK(results);
    REQUIRE(results.value().size() == 2);
    REQUIRE(results.value()[0].key == "v1");
}
TEST_GPU(search_entities_with_sparse, "lib.rs:5121") {
    VectorEngine engine;
    Vec e1(100, 0.0f), e2(100, 0.0f), query(100, 0.0f);
    e1[0] = 1.0f;
    e2[1] = 1.0f;
    query[0] = 1.0f;
    REQUIRE_OK(engine.set_entity_embedding("user:1", e1));
    REQUIRE_OK(engine.set_entity_embedding("user:2", e2));
    auto results = engine.search_entities(query, 2);
    REQUIRE_OK(results);
    REQUIRE(results.value().size() == 2);
    REQUIRE(results.value()[0].key == "user:1");
    REQUIRE(std::fabs(results.value()[0].score - 1.0f) < 0.01f);
}
TEST_HOST(sparse_with_custom_threshold, "lib.rs:5759") {
    VectorEngineConfig config;
    config.sparse_threshold = 0.8f;
    auto engine = VectorEngine::with_config(config);
    REQUIRE_OK(engine);
    Vec mostly(100), very(100);
    for (int i = 0; i < 100; ++i) {
        mostly[i] = i < 30 ? 1.0f : 0.0f;
        very[i] = i < 10 ? 1.0f : 0.0f;
    }
    const float thr = engine.value()->config().sparse_threshold;
    REQUIRE(!VectorEngine::should_use_sparse_with_threshold(mostly, thr));
    REQUIRE(VectorEngine::should_use_sparse_with_threshold(very, thr));
}
TEST_HOST(sparse_detection_empty_vector, "lib.rs:6173") {
    REQUIRE(!VectorEngine::should_use_sparse_with_threshold({}, 0.5f));
}

// ---------------------------------------------------------------------------------------------
// configuration (lib.rs:5147-5213, 6070-6290)
// ---------------------------------------------------------------------------------------------
TEST_HOST(config_default, "lib.rs:5147") {
    VectorEngineConfig config;
    REQUIRE(!config.default_dimension.has_value());
    REQUIRE(std::fabs(config.sparse_threshold - 0.5f) < 1e-6f);
    REQUIRE(config.parallel_threshold == 5000);
    REQUIRE(config.default_metric == DistanceMetric::Cosine);
}
TEST_HOST(config_high_throughput, "lib.rs:5156") {
    REQUIRE(VectorEngineConfig::high_throughput().parallel_threshold == 1000);
}
TEST_HOST(config_low_memory, "lib.rs:5162") {
    REQUIRE(std::fabs(VectorEngineConfig::low_memory().sparse_threshold - 0.3f) < 1e-6f);
}
TEST_HOST(config_validate_valid, "lib.rs:5168") { REQUIRE_OK(VectorEngineConfig{}.validate()); }
TEST_HOST(config_validate_invalid_sparse_threshold, "lib.rs:5174") {
    VectorEngineConfig config;
    config.sparse_threshold = 1.5f;
    REQUIRE_ERR(config.validate(), ConfigurationError);
}
TEST_HOST(config_validate_invalid_parallel_threshold, "lib.rs:5186") {
    VectorEngineConfig config;
    config.parallel_threshold = 0;
    REQUIRE_ERR(config.validate(), ConfigurationError);
}
TEST_HOST(engine_with_config, "lib.rs:5198") {
    auto engine = VectorEngine::with_config(VectorEngineConfig::high_throughput());
    REQUIRE_OK(engine);
    REQUIRE(engine.value()->config().parallel_threshold == 1000);
}
TEST_HOST(config_validate_negative_sparse_threshold, "lib.rs:6070") {
    VectorEngineConfig config;
    config.sparse_threshold = -0.1f;
    REQUIRE_ERR(config.validate(), ConfigurationError);
}
TEST_HOST(config_presets_are_valid, "lib.rs:6083") {
    REQUIRE_OK(VectorEngineConfig{}.validate());
    REQUIRE_OK(VectorEngineConfig::high_throughput().validate());
    REQUIRE_OK(VectorEngineConfig::low_memory().validate());
}
TEST_HOST(config_validate_invalid_max_dimension_zero, "lib.rs:6213") {
    VectorEngineConfig config;
    config.max_dimension = 0;
    auto r = config.validate();
    REQUIRE_ERR(r, ConfigurationError);
    REQUIRE(r.error().message.find("max_dimension") != std::string::npos);
}
TEST_HOST(config_validate_invalid_max_keys_per_scan_zero, "lib.rs:6226") {
    VectorEngineConfig config;
    config.max_keys_per_scan = 0;
    auto r = config.validate();
    REQUIRE_ERR(r, ConfigurationError);
    REQUIRE(r.error().message.find("max_keys_per_scan") != std::string::npos);
}
TEST_HOST(config_validate_invalid_batch_parallel_threshold_zero, "lib.rs:6239") {
    VectorEngineConfig config;
    config.batch_parallel_threshold = 0;
    auto r = config.validate();
    REQUIRE_ERR(r, ConfigurationError);
    REQUIRE(r.error().message.find("batch_parallel_threshold") != std::string::npos);
}
TEST_HOST(with_config_validates_and_returns_error, "lib.rs:6278") {
    VectorEngineConfig config;
    config.max_dimension = 0;
    REQUIRE_ERR(VectorEngine::with_config(config), ConfigurationError);
}
TEST_HOST(with_config_valid_succeeds, "lib.rs:6300") {
    VectorEngineConfig config;
    config.max_dimension = 1024;
    config.max_keys_per_scan = 1000;
    REQUIRE_OK(VectorEngine::with_config(config));
}
TEST_HOST(list_keys_bounded_respects_limit, "lib.rs:6310") {
    VectorEngineConfig config;
    config.max_keys_per_scan = 3;
    auto engine = VectorEngine::with_config(config);
    REQUIRE_OK(engine);
    for (int i = 0; i < 10; ++i) REQUIRE_OK(engine.value()->store_embedding(key_of("v", i), {(float)i}));
    REQUIRE(engine.value()->list_keys_bounded().size() == 3);
}
TEST_HOST(list_keys_bounded_no_limit, "lib.rs:6324") {
    VectorEngine engine;
    for (int i = 0; i < 10; ++i) REQUIRE_OK(engine.store_embedding(key_of("v", i), {(float)i}));
    REQUIRE(engine.list_keys_bounded().size() == 10);
}
TEST_HOST(search_similar_rejects_oversized_dimension, "lib.rs:6336") {
    VectorEngineConfig config;
    config.max_dimension = 10;
    auto engine = VectorEngine::with_config(config);
    REQUIRE_OK(engine);
    auto r = engine.value()->search_similar(Vec(20, 0.0f), 5);
    REQUIRE_ERR(r, DimensionMismatch);
    REQUIRE(r.error().expected == 10 && r.error().got == 20);
}

// ---------------------------------------------------------------------------------------------
// batch operations (lib.rs:5215-5300, 5776-5800) and error displays (lib.rs:5409-5437)
// ---------------------------------------------------------------------------------------------
using EI = VectorEngine::EmbeddingInput;
TEST_HOST(batch_store_embeddings_basic, "lib.rs:5215") {
    VectorEngine engine;
    auto result = engine.batch_store_embeddings({EI{"a", {1.0f, 0.0f}}, EI{"b", {0.0f, 1.0f}}, EI{"c", {1.0f, 1.0f}}});
    REQUIRE_OK(result);
    REQUIRE(result.value().count == 3);
    REQUIRE(result.value().stored_keys == (std::vector<std::string>{"a", "b", "c"}));
    REQUIRE(engine.count() == 3);
}
TEST_HOST(batch_store_embeddings_empty, "lib.rs:5230") {
    VectorEngine engine;
    auto result = engine.batch_store_embeddings({});
    REQUIRE_OK(result);
    REQUIRE(result.value().count == 0 && result.value().stored_keys.empty());
}
TEST_HOST(batch_store_embeddings_validation_error, "lib.rs:5238") {
    VectorEngine engine;
    auto result = engine.batch_store_embeddings({EI{"a", {1.0f, 0.0f}}, EI{"b", {}}});
    REQUIRE_ERR(result, BatchValidationError);
    REQUIRE(result.error().index == 1);
    REQUIRE(engine.count() == 0);  // validation comes before any store
}
TEST_HOST(batch_delete_embeddings_basic, "lib.rs:5253") {
    VectorEngine engine;
    REQUIRE_OK(engine.store_embedding("a", {1.0f}));
    REQUIRE_OK(engine.store_embedding("b", {2.0f}));
    REQUIRE_OK(engine.store_embedding("c", {3.0f}));
    auto count = engine.batch_delete_embeddings({"a", "b"});
    REQUIRE_OK(count);
    REQUIRE(count.value() == 2);
    REQUIRE(engine.count() == 1);
    REQUIRE(engine.exists("c"));
}
TEST_HOST(batch_delete_embeddings_empty, "lib.rs:5268") {
    VectorEngine engine;
    REQUIRE(engine.batch_delete_embeddings({}).value() == 0);
}
TEST_HOST(batch_delete_embeddings_nonexistent, "lib.rs:5275") {
    VectorEngine engine;
    REQUIRE_OK(engine.store_embedding("a", {1.0f}));
    REQUIRE(engine.batch_delete_embeddings({"a", "nonexistent"}).value() == 1);
}
TEST_HOST(batch_store_large_batch, "lib.rs:5776") {
    VectorEngine engine;
    std::vector<EI> inputs;
    for (int i = 0; i < 150; ++i) inputs.push_back(EI{key_of("v", i), {(float)i, 0.0f}});
    auto result = engine.batch_store_embeddings(inputs);
    REQUIRE_OK(result);
    REQUIRE(result.value().count == 150);
    REQUIRE(engine.count() == 150);
}
TEST_HOST(batch_delete_large_batch, "lib.rs:5788") {
    VectorEngine engine;
    std::vector<std::string> keys;
    for (int i = 0; i < 150; ++i) {
        REQUIRE_OK(engine.store_embedding(key_of("v", i), {(float)i}));
        keys.push_back(key_of("v", i));
    }
    REQUIRE(engine.batch_delete_embeddings(keys).value() == 150);
    REQUIRE(engine.count() == 0);
}
TEST_HOST(error_batch_validation_display, "lib.rs:5409") {
    VectorError e;
    e.kind = ErrorKind::BatchValidationError;
    e.index = 5;
    e.message = "test error";
    REQUIRE(e.to_string() == "Batch validation error at index 5: test error");
}
TEST_HOST(error_batch_operation_display, "lib.rs:5421") {
    VectorError e;
    e.kind = ErrorKind::BatchOperationError;
    e.index = 3;
    e.message = "op failed";
    REQUIRE(e.to_string() == "Batch operation error at index 3: op failed");
}
TEST_HOST(error_configuration_display, "lib.rs:5433") {
    VectorError e;
    e.kind = ErrorKind::ConfigurationError;
    e.message = "bad config";
    REQUIRE(e.to_string() == "Configuration error: bad config");
}

// ---------------------------------------------------------------------------------------------
// pagination (lib.rs:5302-5407, 5803-5826, 6153-6170)
// ---------------------------------------------------------------------------------------------
using Pg = VectorEngine::Pagination;
TEST_HOST(pagination_new, "lib.rs:5302") {
    const Pg p = Pg::with(10, 20);
    REQUIRE(p.skip == 10 && p.limit == std::optional<size_t>(20) && !p.count_total);
}
TEST_HOST(pagination_with_total, "lib.rs:5310") { REQUIRE(Pg::with(0, 10).with_total().count_total); }
TEST_HOST(pagination_skip_only, "lib.rs:5316") {
    const Pg p = Pg::skip_only(5);
    REQUIRE(p.skip == 5 && !p.limit.has_value());
}
TEST_HOST(pagination_default, "lib.rs:6163") {
    const Pg p;
    REQUIRE(p.skip == 0 && !p.limit.has_value() && !p.count_total);
}
static void store_ten(VectorEngine &engine) {
    char buf[16];
    for (int i = 0; i < 10; ++i) {
        std::snprintf(buf, sizeof buf, "v%02d", i);
        engine.store_embedding(buf, {(float)i});
    }
}
TEST_HOST(list_keys_paginated_basic, "lib.rs:5323") {
    VectorEngine engine;
    store_ten(engine);
    auto result = engine.list_keys_paginated(Pg::with(0, 3));
    REQUIRE(result.items.size() == 3);
    REQUIRE(result.has_more);
    REQUIRE(!result.total_count.has_value());
}
TEST_HOST(list_keys_paginated_with_total, "lib.rs:5338") {
    VectorEngine engine;
    store_ten(engine);
    auto result = engine.list_keys_paginated(Pg::with(0, 5).with_total());
    REQUIRE(result.items.size() == 5);
    REQUIRE(result.total_count == std::optional<size_t>(10));
    REQUIRE(result.has_more);
}
TEST_HOST(list_keys_paginated_skip, "lib.rs:5353") {
    VectorEngine engine;
    store_ten(engine);
    auto result = engine.list_keys_paginated(Pg::with(8, 5).with_total());
    REQUIRE(result.items.size() == 2);
    REQUIRE(!result.has_more);
}
TEST_GPU(search_similar_paginated_basic, "lib.rs:5367") {
    VectorEngine engine;
    for (int i = 0; i < 10; ++i) REQUIRE_OK(engine.store_embedding(key_of("v", i), {(float)i, 0.0f}));
    auto result = engine.search_similar_paginated({5.0f, 0.0f}, 10, Pg::with(0, 3).with_total());
    REQUIRE_OK(result);
    REQUIRE(result.value().items.size() == 3);
}
TEST_GPU(search_entities_paginated_basic, "lib.rs:5383") {
    VectorEngine engine;
    for (int i = 0; i < 5; ++i) REQUIRE_OK(engine.set_entity_embedding(key_of("user:", i), {(float)i, 0.0f}));
    auto result = engine.search_entities_paginated({2.0f, 0.0f}, 5, Pg::with(0, 2).with_total());
    REQUIRE_OK(result);
    REQUIRE(result.value().items.size() == 2);
}
TEST_HOST(paged_result_empty, "lib.rs:5399") {
    auto result = VectorEngine::PagedResult<std::string>::empty();
    REQUIRE(result.items.empty());
    REQUIRE(result.total_count == std::optional<size_t>(0));
    REQUIRE(!result.has_more);
}
TEST_HOST(pagination_empty_result, "lib.rs:5803") {
    VectorEngine engine;
    auto result = engine.list_keys_paginated(Pg::with(0, 10).with_total());
    REQUIRE(result.items.empty());
    REQUIRE(result.total_count == std::optional<size_t>(0));
    REQUIRE(!result.has_more);
}
TEST_HOST(pagination_skip_past_end, "lib.rs:5812") {
    VectorEngine engine;
    for (int i = 0; i < 5; ++i) REQUIRE_OK(engine.store_embedding(key_of("v", i), {(float)i}));
    auto result = engine.list_keys_paginated(Pg::with(10, 5).with_total());
    REQUIRE(result.items.empty());
    REQUIRE(result.total_count == std::optional<size_t>(5));
    REQUIRE(!result.has_more);
}

// ---------------------------------------------------------------------------------------------
// concurrency (lib.rs:5541-5745)
// ---------------------------------------------------------------------------------------------
TEST_HOST(test_concurrent_store_embedding_same_key, "lib.rs:5541") {
    VectorEngine engine;
    std::atomic<int> success{0};
    std::vector<std::thread> ts;
    for (int i = 0; i < 10; ++i)
        ts.emplace_back([&, i] {
            if (engine.store_embedding("contested", {(float)i, (float)i}).is_ok()) ++success;
        });
    for (auto &t : ts) t.join();
    REQUIRE(success == 10);  // last write wins
    REQUIRE(engine.exists("contested"));
}
TEST_HOST(test_concurrent_delete_embedding_same_key, "lib.rs:5575") {
    VectorEngine engine;
    REQUIRE_OK(engine.store_embedding("to_delete", {1.0f, 2.0f}));
    std::atomic<int> success{0}, not_found{0}, other{0};
    std::vector<std::thread> ts;
    for (int i = 0; i < 10; ++i)
        ts.emplace_back([&] {
            auto r = engine.delete_embedding("to_delete");
            if (r.is_ok()) ++success;
            else if (r.error().kind == ErrorKind::NotFound) ++not_found;
            else ++other;
        });
    for (auto &t : ts) t.join();
    REQUIRE(success == 1 && not_found == 9 && other == 0);
}
TEST_GPU(test_concurrent_search_similar, "lib.rs:5615") {
    VectorEngine engine;
    for (int i = 0; i < 100; ++i) REQUIRE_OK(engine.store_embedding(key_of("v", i), {(float)i, 0.0f}));
    REQUIRE(engine.count() == 100);
    std::atomic<int> success{0};
    std::vector<std::thread> ts;
    for (int t = 0; t < 20; ++t)
        ts.emplace_back([&, t] {
            for (int j = 0; j < 5; ++j) {
                auto r = engine.search_similar({(float)t, 0.0f}, 5);
                if (r.is_ok() && !r.value().empty()) ++success;
            }
        });
    for (auto &t : ts) t.join();
    // the reference asks for > 50 of 100 ("some may transiently see partial state"); thread 0's
    // query is the zero vector and returns nothing by design
    REQUIRE(success == 95);
}
TEST_GPU(test_concurrent_store_and_search, "lib.rs:5669") {
    VectorEngine engine;
    for (int i = 0; i < 50; ++i) REQUIRE_OK(engine.store_embedding(key_of("init", i), {(float)i, 0.0f}));
    std::atomic<int> failures{0};
    std::vector<std::thread> ts;
    for (int t = 0; t < 20; ++t)
        ts.emplace_back([&, t] {
            if (t % 2 == 0) {
                for (int i = 0; i < 10; ++i)
                    if (engine.store_embedding("t" + std::to_string(t) + "_v" + std::to_string(i), {(float)t, (float)i}).is_err())
                        ++failures;
            } else {
                for (int i = 0; i < 10; ++i)
                    if (engine.search_similar({(float)t, 0.0f}, 5).is_err()) ++failures;
            }
        });
    for (auto &t : ts) t.join();
    REQUIRE(failures == 0);
    REQUIRE(engine.count() == 150);
}
TEST_HOST(test_concurrent_batch_operations, "lib.rs:5713") {
    VectorEngine engine;
    std::vector<size_t> counts(5, 0);
    std::vector<std::thread> ts;
    for (int t = 0; t < 5; ++t)
        ts.emplace_back([&, t] {
            std::vector<EI> inputs;
            for (int i = 0; i < 20; ++i)
                inputs.push_back(EI{"t" + std::to_string(t) + "_b" + std::to_string(i), {(float)t, (float)i}});
            auto r = engine.batch_store_embeddings(inputs);
            counts[t] = r.is_ok() ? r.value().count : 0;
        });
    for (auto &t : ts) t.join();
    for (size_t c : counts) REQUIRE(c == 20);
    REQUIRE(engine.count() == 100);
}

// ---------------------------------------------------------------------------------------------
// numeric and dimension edge cases (lib.rs:5949-6060)
// ---------------------------------------------------------------------------------------------
TEST_GPU(store_and_search_with_very_small_values, "lib.rs:5949") {
    VectorEngine engine;
    Vec tiny{1e-18f, 1e-18f, 1e-18f};
    REQUIRE_OK(engine.store_embedding("tiny", tiny));
    auto results = engine.search_similar(tiny, 1);
    REQUIRE_OK(results);
    REQUIRE(results.value().size() == 1);
    REQUIRE(results.value()[0].key == "tiny");
}
TEST_GPU(store_and_search_with_large_values, "lib.rs:5963") {
    VectorEngine engine;
    Vec large{1e30f, 1e30f, 1e30f};
    REQUIRE_OK(engine.store_embedding("large", large));
    auto results = engine.search_similar(large, 1);
    REQUIRE_OK(results);
    REQUIRE(results.value().size() == 1);
}
TEST_GPU(search_handles_denormalized_floats, "lib.rs:5974") {
    VectorEngine engine;
    Vec denorm{1.17549435e-38f / 2.0f, 1.0f, 0.0f};
    REQUIRE_OK(engine.store_embedding("denorm", denorm));
    auto results = engine.search_similar(denorm, 1);
    REQUIRE_OK(results);
    REQUIRE(results.value().size() == 1);
}
TEST_GPU(zero_vector_with_euclidean_metric, "lib.rs:5985") {
    VectorEngine engine;
    REQUIRE_OK(engine.store_embedding("a", {1.0f, 0.0f}));
    REQUIRE_OK(engine.store_embedding("b", {0.0f, 0.0f}));
    auto results = engine.search_similar_with_metric({0.0f, 0.0f}, 2, DistanceMetric::Euclidean);
    REQUIRE_OK(results);
    REQUIRE(results.value().size() == 2);
    REQUIRE(results.value()[0].key == "b");
}
TEST_GPU(single_dimension_vector, "lib.rs:6004") {
    VectorEngine engine;
    REQUIRE_OK(engine.store_embedding("a", {1.0f}));
    REQUIRE_OK(engine.store_embedding("b", {2.0f}));
    REQUIRE_OK(engine.store_embedding("c", {-1.0f}));
    auto results = engine.search_similar({1.0f}, 3);
    REQUIRE_OK(results);
    REQUIRE(results.value().size() == 3);
    REQUIRE(results.value()[0].key == "a" || results.value()[0].key == "b");
    REQUIRE(results.value()[2].key == "c");
}
TEST_GPU(high_dimension_4096, "lib.rs:6024") {
    VectorEngine engine;
    Vec v1(4096), v2(4096);
    for (int i = 0; i < 4096; ++i) {
        v1[i] = std::sin((float)i * 0.001f);
        v2[i] = std::sin((float)i * 0.002f);
    }
    REQUIRE_OK(engine.store_embedding("v1", v1));
    REQUIRE_OK(engine.store_embedding("v2", v2));
    auto results = engine.search_similar(v1, 2);
    REQUIRE_OK(results);
    REQUIRE(results.value().size() == 2);
    REQUIRE(results.value()[0].key == "v1");
}
TEST_GPU(mismatched_dimensions_silently_skipped, "lib.rs:6040") {
    VectorEngine engine;
    REQUIRE_OK(engine.store_embedding("dim2", {1.0f, 0.0f}));
    REQUIRE_OK(engine.store_embedding("dim3", {1.0f, 0.0f, 0.0f}));
    REQUIRE_OK(engine.store_embedding("dim4", {1.0f, 0.0f, 0.0f, 0.0f}));
    auto results = engine.search_similar({1.0f, 0.0f}, 10);
    REQUIRE_OK(results);
    REQUIRE(results.value().size() == 1);
    REQUIRE(results.value()[0].key == "dim2");
}

// ---------------------------------------------------------------------------------------------
// max_dimension on the other entry points (lib.rs:6352-6390), parallel path (lib.rs:6583)
// ---------------------------------------------------------------------------------------------
TEST_HOST(set_entity_embedding_rejects_oversized_dimension, "lib.rs:6352") {
    VectorEngineConfig config;
    config.max_dimension = 5;
    auto engine = VectorEngine::with_config(config);
    REQUIRE_OK(engine);
    auto r = engine.value()->set_entity_embedding("entity:1", Vec(10, 0.0f));
    REQUIRE_ERR(r, DimensionMismatch);
    REQUIRE(r.error().expected == 5 && r.error().got == 10);
}
TEST_HOST(search_entities_rejects_oversized_dimension, "lib.rs:6372") {
    VectorEngineConfig config;
    config.max_dimension = 5;
    auto engine = VectorEngine::with_config(config);
    REQUIRE_OK(engine);
    auto r = engine.value()->search_entities(Vec(10, 0.0f), 5);
    REQUIRE_ERR(r, DimensionMismatch);
    REQUIRE(r.error().expected == 5 && r.error().got == 10);
}
TEST_GPU(search_with_metric_parallel_path, "lib.rs:6583") {
    VectorEngineConfig config;
    config.parallel_threshold = 5;
    auto engine = VectorEngine::with_config(config);
    REQUIRE_OK(engine);
    for (int i = 0; i < 10; ++i) REQUIRE_OK(engine.value()->store_embedding(key_of("vec_", i), {(float)i, 0.0f, 0.0f}));
    auto results = engine.value()->search_similar_with_metric({5.0f, 0.0f, 0.0f}, 3, DistanceMetric::Euclidean);
    REQUIRE_OK(results);
    REQUIRE(results.value().size() == 3);
}

// ---------------------------------------------------------------------------------------------
// metadata storage (lib.rs:6622-6965)
// ---------------------------------------------------------------------------------------------
TEST_HOST(store_embedding_with_metadata_basic, "lib.rs:6622") {
    VectorEngine engine;
    REQUIRE_OK(engine.store_embedding_with_metadata("product1", {0.1f, 0.2f, 0.3f},
                                                    meta({{"category", S("electronics")}, {"price", F(299.99)}})));
    REQUIRE(engine.get_embedding("product1").value().size() == 3);
    auto m = engine.get_metadata("product1");
    REQUIRE_OK(m);
    REQUIRE(m.value().size() == 2 && m.value().count("category") && m.value().count("price"));
}
TEST_HOST(store_embedding_with_metadata_empty_metadata, "lib.rs:6650") {
    VectorEngine engine;
    REQUIRE_OK(engine.store_embedding_with_metadata("key", {1.0f, 2.0f}, {}));
    REQUIRE(engine.get_embedding("key").value() == (Vec{1.0f, 2.0f}));
    REQUIRE(engine.get_metadata("key").value().empty());
}
TEST_HOST(store_embedding_with_metadata_empty_vector_error, "lib.rs:6666") {
    VectorEngine engine;
    REQUIRE_ERR(engine.store_embedding_with_metadata("key", {}, {}), EmptyVector);
}
TEST_HOST(store_embedding_with_metadata_dimension_limit, "lib.rs:6675") {
    VectorEngineConfig config;
    config.max_dimension = 5;
    auto engine = VectorEngine::with_config(config);
    REQUIRE_OK(engine);
    auto r = engine.value()->store_embedding_with_metadata("key", Vec(10, 0.0f), {});
    REQUIRE_ERR(r, DimensionMismatch);
    REQUIRE(r.error().expected == 5 && r.error().got == 10);
}
TEST_HOST(get_metadata_nonexistent_key, "lib.rs:6693") {
    VectorEngine engine;
    REQUIRE_ERR(engine.get_metadata("nonexistent"), NotFound);
}
TEST_HOST(update_metadata_basic, "lib.rs:6700") {
    VectorEngine engine;
    REQUIRE_OK(engine.store_embedding_with_metadata("item", {1.0f, 2.0f}, meta({{"color", S("red")}})));
    REQUIRE_OK(engine.update_metadata("item", meta({{"size", S("large")}, {"color", S("blue")}})));
    auto m = engine.get_metadata("item");
    REQUIRE_OK(m);
    REQUIRE(m.value().size() == 2);
    REQUIRE(m.value().at("color").type == MetadataValue::Type::String && m.value().at("color").s == "blue");
    REQUIRE(m.value().count("size") == 1);
}
TEST_HOST(update_metadata_nonexistent_key, "lib.rs:6739") {
    VectorEngine engine;
    REQUIRE_ERR(engine.update_metadata("nonexistent", {}), NotFound);
}
TEST_HOST(remove_metadata_field_basic, "lib.rs:6763") {
    VectorEngine engine;
    REQUIRE_OK(engine.store_embedding_with_metadata("key", {1.0f}, meta({{"a", I(1)}, {"b", I(2)}})));
    REQUIRE_OK(engine.remove_metadata_field("key", "a"));
    auto m = engine.get_metadata("key");
    REQUIRE_OK(m);
    REQUIRE(!m.value().count("a") && m.value().count("b"));
}
TEST_HOST(remove_metadata_field_nonexistent_key, "lib.rs:6781") {
    VectorEngine engine;
    REQUIRE_ERR(engine.remove_metadata_field("nonexistent", "field"), NotFound);
}
TEST_HOST(has_metadata_field_true, "lib.rs:6788") {
    VectorEngine engine;
    REQUIRE_OK(engine.store_embedding_with_metadata("key", {1.0f}, meta({{"field1", I(42)}})));
    REQUIRE(engine.has_metadata_field("key", "field1"));
}
TEST_HOST(has_metadata_field_false, "lib.rs:6804") {
    VectorEngine engine;
    REQUIRE_OK(engine.store_embedding("key", {1.0f}));
    REQUIRE(!engine.has_metadata_field("key", "nonexistent_field"));
}
TEST_HOST(has_metadata_field_nonexistent_key, "lib.rs:6812") {
    VectorEngine engine;
    REQUIRE(!engine.has_metadata_field("nonexistent", "field"));
}
TEST_HOST(get_metadata_field_basic, "lib.rs:6818") {
    VectorEngine engine;
    REQUIRE_OK(engine.store_embedding_with_metadata("key", {1.0f}, meta({{"score", F(0.95)}})));
    auto v = engine.get_metadata_field("key", "score");
    REQUIRE_OK(v);
    REQUIRE(v.value().has_value() && v.value()->type == MetadataValue::Type::Float);
    REQUIRE(std::fabs(v.value()->f - 0.95) < 2.3e-16);
}
TEST_HOST(get_metadata_field_not_present, "lib.rs:6840") {
    VectorEngine engine;
    REQUIRE_OK(engine.store_embedding("key", {1.0f}));
    auto v = engine.get_metadata_field("key", "nonexistent");
    REQUIRE_OK(v);
    REQUIRE(!v.value().has_value());
}
TEST_HOST(get_metadata_field_nonexistent_key, "lib.rs:6849") {
    VectorEngine engine;
    REQUIRE_ERR(engine.get_metadata_field("nonexistent", "field"), NotFound);
}
TEST_HOST(metadata_with_sparse_vector, "lib.rs:6856") {
    VectorEngine engine;
    Vec sparse(100, 0.0f);
    sparse[0] = 1.0f;
    sparse[50] = 2.0f;
    REQUIRE_OK(engine.store_embedding_with_metadata("sparse_key", sparse, meta({{"type", S("sparse")}})));
    auto got = engine.get_embedding("sparse_key");
    REQUIRE_OK(got);
    REQUIRE(got.value().size() == 100 && got.value()[0] == 1.0f && got.value()[50] == 2.0f);
    REQUIRE(engine.get_metadata("sparse_key").value().count("type") == 1);
}
TEST_HOST(metadata_multiple_types, "lib.rs:6891") {
    VectorEngine engine;
    REQUIRE_OK(engine.store_embedding_with_metadata(
        "multi_type", {1.0f, 2.0f},
        meta({{"int_field", I(42)}, {"float_field", F(3.14)}, {"string_field", S("hello")}, {"bool_field", B(true)}})));
    auto m = engine.get_metadata("multi_type");
    REQUIRE_OK(m);
    REQUIRE(m.value().size() == 4);
    REQUIRE(m.value().at("int_field").type == MetadataValue::Type::Int && m.value().at("int_field").i == 42);
    REQUIRE(m.value().at("float_field").type == MetadataValue::Type::Float &&
            std::fabs(m.value().at("float_field").f - 3.14) < 2.3e-16);
    REQUIRE(m.value().at("string_field").type == MetadataValue::Type::String && m.value().at("string_field").s == "hello");
    REQUIRE(m.value().at("bool_field").type == MetadataValue::Type::Bool && m.value().at("bool_field").b);
}
TEST_HOST(metadata_overwrites_on_store, "lib.rs:6940") {
    VectorEngine engine;
    REQUIRE_OK(engine.store_embedding_with_metadata("key", {1.0f}, meta({{"a", I(1)}, {"b", I(2)}})));
    REQUIRE_OK(engine.store_embedding_with_metadata("key", {2.0f}, meta({{"c", I(3)}})));
    auto m = engine.get_metadata("key");
    REQUIRE_OK(m);
    REQUIRE(m.value().size() == 1 && m.value().count("c") && !m.value().count("a") && !m.value().count("b"));
}

// ---------------------------------------------------------------------------------------------
// filtered search (lib.rs:6968-7720).  The filter is evaluated ON THE DEVICE over typed metadata
// columns (filter_kernels.cuh) for the pre-filter strategy and on the host for post-filtering.
// ---------------------------------------------------------------------------------------------
static void setup_filtered_search_engine(VectorEngine &engine) {  // lib.rs:6968-7001
    const char *categories[3] = {"electronics", "clothing", "food"};
    const int64_t prices[3] = {100, 50, 25};
    for (int i = 0; i < 3; ++i)
        engine.store_embedding_with_metadata(
            key_of("item", i), {(float)(i + 1), 1.0f, 1.0f},
            meta({{"category", S(categories[i])}, {"price", I(prices[i])}, {"active", B(i % 2 == 0)}}));
}
static FilterCondition Cmp(Op op, const char *field, MetadataValue v) { return FilterCondition::cmp(op, field, std::move(v)); }
#define FILTERED(engine, q, k, f) (engine).search_similar_filtered(q, k, f)
#define FILTER_COUNT_TEST(name, src, filter_expr, expected)              \
    TEST_GPU(name, src) {                                                \
        VectorEngine engine;                                             \
        setup_filtered_search_engine(engine);                            \
        auto results = FILTERED(engine, (Vec{1.0f, 1.0f, 1.0f}), 10, filter_expr); \
        REQUIRE_OK(results);                                             \
        REQUIRE(results.value().size() == (expected));                   \
    }
TEST_GPU(search_filtered_eq_string, "lib.rs:7004") {
    VectorEngine engine;
    setup_filtered_search_engine(engine);
    auto results = FILTERED(engine, (Vec{1.0f, 1.0f, 1.0f}), 10, Cmp(Op::Eq, "category", S("electronics")));
    REQUIRE_OK(results);
    REQUIRE(results.value().size() == 1);
    REQUIRE(results.value()[0].key == "item0");
}
TEST_GPU(search_filtered_eq_int, "lib.rs:7020") {
    VectorEngine engine;
    setup_filtered_search_engine(engine);
    auto results = FILTERED(engine, (Vec{1.0f, 0.0f, 0.0f}), 10, Cmp(Op::Eq, "price", I(50)));
    REQUIRE_OK(results);
    REQUIRE(results.value().size() == 1);
    REQUIRE(results.value()[0].key == "item1");
}
FILTER_COUNT_TEST(search_filtered_gt, "lib.rs:7033", Cmp(Op::Gt, "price", I(30)), 2)
FILTER_COUNT_TEST(search_filtered_lt, "lib.rs:7045", Cmp(Op::Lt, "price", I(60)), 2)
FILTER_COUNT_TEST(search_filtered_le, "lib.rs:7057", Cmp(Op::Le, "price", I(50)), 2)
FILTER_COUNT_TEST(search_filtered_ge, "lib.rs:7070", Cmp(Op::Ge, "price", I(50)), 2)
TEST_GPU(search_filtered_and, "lib.rs:7083") {
    VectorEngine engine;
    setup_filtered_search_engine(engine);
    auto results = FILTERED(engine, (Vec{1.0f, 1.0f, 1.0f}), 10,
                            Cmp(Op::Gt, "price", I(30)).and_(Cmp(Op::Lt, "price", I(80))));
    REQUIRE_OK(results);
    REQUIRE(results.value().size() == 1);
    REQUIRE(results.value()[0].key == "item1");
}
FILTER_COUNT_TEST(search_filtered_or, "lib.rs:7098",
                  Cmp(Op::Eq, "category", S("electronics")).or_(Cmp(Op::Eq, "category", S("food"))), 2)
FILTER_COUNT_TEST(search_filtered_true, "lib.rs:7117", FilterCondition::always(), 3)
TEST_GPU(search_filtered_exists, "lib.rs:7129") {
    VectorEngine engine;
    REQUIRE_OK(engine.store_embedding_with_metadata("with_tag", {1.0f, 0.0f}, meta({{"tag", S("a")}})));
    REQUIRE_OK(engine.store_embedding("without_tag", {0.0f, 1.0f}));
    auto results = FILTERED(engine, (Vec{1.0f, 0.0f}), 10, FilterCondition::exists("tag"));
    REQUIRE_OK(results);
    REQUIRE(results.value().size() == 1);
    REQUIRE(results.value()[0].key == "with_tag");
}
TEST_GPU(search_filtered_contains, "lib.rs:7155") {
    VectorEngine engine;
    REQUIRE_OK(engine.store_embedding_with_metadata("item1", {1.0f, 0.0f}, meta({{"description", S("blue shirt")}})));
    REQUIRE_OK(engine.store_embedding_with_metadata("item2", {0.0f, 1.0f}, meta({{"description", S("red pants")}})));
    auto results = FILTERED(engine, (Vec{1.0f, 0.0f}), 10, FilterCondition::contains("description", "shirt"));
    REQUIRE_OK(results);
    REQUIRE(results.value().size() == 1);
    REQUIRE(results.value()[0].key == "item1");
}
TEST_GPU(search_filtered_contains_on_non_string, "lib.rs:7186") {
    VectorEngine engine;
    REQUIRE_OK(engine.store_embedding_with_metadata("item", {1.0f, 0.0f}, meta({{"count", I(42)}})));
    auto results = FILTERED(engine, (Vec{1.0f, 0.0f}), 10, FilterCondition::contains("count", "4"));
    REQUIRE_OK(results);
    REQUIRE(results.value().empty());
}
TEST_GPU(search_filtered_starts_with, "lib.rs:7208") {
    VectorEngine engine;
    REQUIRE_OK(engine.store_embedding_with_metadata("item1", {1.0f, 0.0f}, meta({{"sku", S("ABC123")}})));
    REQUIRE_OK(engine.store_embedding_with_metadata("item2", {0.0f, 1.0f}, meta({{"sku", S("XYZ789")}})));
    auto results = FILTERED(engine, (Vec{1.0f, 0.0f}), 10, FilterCondition::starts_with("sku", "ABC"));
    REQUIRE_OK(results);
    REQUIRE(results.value().size() == 1);
    REQUIRE(results.value()[0].key == "item1");
}
TEST_GPU(search_filtered_starts_with_on_non_string, "lib.rs:7239") {
    VectorEngine engine;
    REQUIRE_OK(engine.store_embedding_with_metadata("item", {1.0f, 0.0f}, meta({{"count", I(123)}})));
    auto results = FILTERED(engine, (Vec{1.0f, 0.0f}), 10, FilterCondition::starts_with("count", "1"));
    REQUIRE_OK(results);
    REQUIRE(results.value().empty());
}
TEST_GPU(search_filtered_missing_field, "lib.rs:7261") {
    VectorEngine engine;
    REQUIRE_OK(engine.store_embedding("item", {1.0f, 0.0f}));
    auto results = FILTERED(engine, (Vec{1.0f, 0.0f}), 10, Cmp(Op::Eq, "missing", I(42)));
    REQUIRE_OK(results);
    REQUIRE(results.value().empty());
}
FILTER_COUNT_TEST(search_filtered_in, "lib.rs:7277", FilterCondition::in("category", {S("electronics"), S("food")}), 2)
FILTER_COUNT_TEST(search_filtered_ne, "lib.rs:7295", Cmp(Op::Ne, "category", S("electronics")), 2)
FILTER_COUNT_TEST(search_filtered_bool, "lib.rs:7310", Cmp(Op::Eq, "active", B(true)), 2)
FILTER_COUNT_TEST(search_filtered_empty_result, "lib.rs:7322", Cmp(Op::Eq, "category", S("nonexistent")), 0)
TEST_GPU(search_filtered_pre_filter_strategy, "lib.rs:7337") {
    VectorEngine engine;
    setup_filtered_search_engine(engine);
    auto results = engine.search_similar_filtered({1.0f, 1.0f, 1.0f}, 10, Cmp(Op::Eq, "category", S("electronics")),
                                                  FilteredSearchConfig::pre_filter());
    REQUIRE_OK(results);
    REQUIRE(results.value().size() == 1);
}
TEST_GPU(search_filtered_post_filter_strategy, "lib.rs:7353") {
    VectorEngine engine;
    setup_filtered_search_engine(engine);
    auto results = engine.search_similar_filtered({1.0f, 1.0f, 1.0f}, 10, Cmp(Op::Eq, "category", S("electronics")),
                                                  FilteredSearchConfig::post_filter());
    REQUIRE_OK(results);
    REQUIRE(results.value().size() == 1);
}
TEST_HOST(search_filtered_empty_vector_error, "lib.rs:7369") {
    VectorEngine engine;
    REQUIRE_ERR(engine.search_similar_filtered({}, 5, FilterCondition::always()), EmptyVector);
}
TEST_HOST(search_filtered_zero_top_k_error, "lib.rs:7377") {
    VectorEngine engine;
    REQUIRE_ERR(engine.search_similar_filtered({1.0f}, 0, FilterCondition::always()), InvalidTopK);
}
TEST_HOST(search_filtered_dimension_limit, "lib.rs:7385") {
    VectorEngineConfig config;
    config.max_dimension = 5;
    auto engine = VectorEngine::with_config(config);
    REQUIRE_OK(engine);
    auto r = engine.value()->search_similar_filtered(Vec(10, 0.0f), 5, FilterCondition::always());
    REQUIRE_ERR(r, DimensionMismatch);
    REQUIRE(r.error().expected == 5 && r.error().got == 10);
}
TEST_HOST(count_matching_basic, "lib.rs:7404") {
    VectorEngine engine;
    setup_filtered_search_engine(engine);
    REQUIRE(engine.count_matching(Cmp(Op::Gt, "price", I(30))) == 2);
}
TEST_HOST(list_keys_matching_basic, "lib.rs:7414") {
    VectorEngine engine;
    setup_filtered_search_engine(engine);
    auto keys = engine.list_keys_matching(Cmp(Op::Eq, "category", S("electronics")));
    REQUIRE(keys.size() == 1 && keys[0] == "item0");
}
TEST_HOST(estimate_filter_selectivity_basic, "lib.rs:7428") {
    VectorEngine engine;
    setup_filtered_search_engine(engine);
    REQUIRE(std::fabs(engine.estimate_filter_selectivity(FilterCondition::always()) - 1.0f) < 0.01f);
    const float s = engine.estimate_filter_selectivity(Cmp(Op::Eq, "category", S("electronics")));
    REQUIRE(s > 0.0f && s < 1.0f);
}
TEST_HOST(filter_condition_and_or_builders, "lib.rs:7444") {
    auto a = Cmp(Op::Eq, "x", I(1)), b = Cmp(Op::Eq, "y", I(2));
    REQUIRE(a.and_(b).op == Op::And);
    REQUIRE(a.or_(b).op == Op::Or);
}
TEST_HOST(filter_strategy_default, "lib.rs:7474") { REQUIRE(FilteredSearchConfig{}.strategy == FilterStrategy::Auto); }
TEST_HOST(filtered_search_config_builders, "lib.rs:7479") {
    REQUIRE(FilteredSearchConfig::pre_filter().strategy == FilterStrategy::PreFilter);
    REQUIRE(FilteredSearchConfig::post_filter().strategy == FilterStrategy::PostFilter);
    REQUIRE(FilteredSearchConfig{}.with_oversample(5).oversample_factor == 5);
}
TEST_GPU(search_filtered_float_comparison, "lib.rs:7491") {
    VectorEngine engine;
    REQUIRE_OK(engine.store_embedding_with_metadata("high", {1.0f, 0.0f}, meta({{"score", F(0.95)}})));
    REQUIRE_OK(engine.store_embedding_with_metadata("low", {0.0f, 1.0f}, meta({{"score", F(0.5)}})));
    auto results = FILTERED(engine, (Vec{1.0f, 0.0f}), 10, Cmp(Op::Gt, "score", F(0.8)));
    REQUIRE_OK(results);
    REQUIRE(results.value().size() == 1);
    REQUIRE(results.value()[0].key == "high");
}
TEST_GPU(search_filtered_mixed_int_float_comparison, "lib.rs:7522") {
    VectorEngine engine;
    REQUIRE_OK(engine.store_embedding_with_metadata("item", {1.0f, 0.0f}, meta({{"value", F(50.5)}})));
    auto results = FILTERED(engine, (Vec{1.0f, 0.0f}), 10, Cmp(Op::Gt, "value", I(50)));
    REQUIRE_OK(results);
    REQUIRE(results.value().size() == 1);
}
TEST_GPU(search_filtered_int_vs_float_filter, "lib.rs:7544") {
    VectorEngine engine;
    REQUIRE_OK(engine.store_embedding_with_metadata("item", {1.0f, 0.0f}, meta({{"count", I(100)}})));
    auto results = FILTERED(engine, (Vec{1.0f, 0.0f}), 10, Cmp(Op::Gt, "count", F(50.5)));
    REQUIRE_OK(results);
    REQUIRE(results.value().size() == 1);
    results = FILTERED(engine, (Vec{1.0f, 0.0f}), 10, Cmp(Op::Gt, "count", F(100.0)));
    REQUIRE_OK(results);
    REQUIRE(results.value().empty());
}
TEST_GPU(search_filtered_null_comparison, "lib.rs:7575") {
    VectorEngine engine;
    REQUIRE_OK(engine.store_embedding_with_metadata("with_null", {1.0f, 0.0f}, meta({{"optional", MetadataValue::null()}})));
    REQUIRE_OK(engine.store_embedding("without_field", {0.0f, 1.0f}));
    auto results = FILTERED(engine, (Vec{1.0f, 0.0f}), 10, Cmp(Op::Eq, "optional", MetadataValue::null()));
    REQUIRE_OK(results);
    REQUIRE(results.value().size() == 1);
    REQUIRE(results.value()[0].key == "with_null");
}
TEST_GPU(search_filtered_string_comparison, "lib.rs:7604") {
    VectorEngine engine;
    REQUIRE_OK(engine.store_embedding_with_metadata("item1", {1.0f, 0.0f}, meta({{"name", S("apple")}})));
    REQUIRE_OK(engine.store_embedding_with_metadata("item2", {0.0f, 1.0f}, meta({{"name", S("banana")}})));
    auto results = FILTERED(engine, (Vec{1.0f, 1.0f}), 10, Cmp(Op::Gt, "name", S("app")));
    REQUIRE_OK(results);
    REQUIRE(results.value().size() == 2);
    results = FILTERED(engine, (Vec{1.0f, 1.0f}), 10, Cmp(Op::Le, "name", S("apple")));
    REQUIRE_OK(results);
    REQUIRE(results.value().size() == 1);
    REQUIRE(results.value()[0].key == "item1");
}
TEST_GPU(search_filtered_bool_false, "lib.rs:7647") {
    VectorEngine engine;
    REQUIRE_OK(engine.store_embedding_with_metadata("active_item", {1.0f, 0.0f}, meta({{"active", B(true)}})));
    REQUIRE_OK(engine.store_embedding_with_metadata("inactive_item", {0.0f, 1.0f}, meta({{"active", B(false)}})));
    auto results = FILTERED(engine, (Vec{1.0f, 1.0f}), 10, Cmp(Op::Eq, "active", B(false)));
    REQUIRE_OK(results);
    REQUIRE(results.value().size() == 1);
    REQUIRE(results.value()[0].key == "inactive_item");
}
TEST_GPU(search_filtered_incompatible_types, "lib.rs:7679") {
    VectorEngine engine;
    REQUIRE_OK(engine.store_embedding_with_metadata("item", {1.0f, 0.0f}, meta({{"value", S("text")}})));
    auto results = FILTERED(engine, (Vec{1.0f, 0.0f}), 10, Cmp(Op::Eq, "value", I(42)));
    REQUIRE_OK(results);
    REQUIRE(results.value().empty());
}
TEST_GPU(search_filtered_respects_top_k, "lib.rs:7701") {
    VectorEngine engine;
    for (int i = 0; i < 10; ++i)
        REQUIRE_OK(engine.store_embedding_with_metadata(key_of("item", i), {(float)i, 0.0f}, meta({{"idx", I(i)}})));
    auto results = FILTERED(engine, (Vec{5.0f, 0.0f}), 3, FilterCondition::always());
    REQUIRE_OK(results);
    REQUIRE(results.value().size() == 3);
}

// ---------------------------------------------------------------------------------------------
// collections (lib.rs:7723-8055, 9336-9440)
// ---------------------------------------------------------------------------------------------
TEST_HOST(create_collection_basic, "lib.rs:7723") {
    VectorEngine engine;
    REQUIRE_OK(engine.create_collection("test", VectorCollectionConfig{}));
    REQUIRE(engine.collection_exists("test"));
    REQUIRE(engine.get_collection_config("test").has_value());
}
TEST_HOST(create_collection_already_exists, "lib.rs:7734") {
    VectorEngine engine;
    REQUIRE_OK(engine.create_collection("test", VectorCollectionConfig{}));
    REQUIRE_ERR(engine.create_collection("test", VectorCollectionConfig{}), CollectionExists);
}
TEST_HOST(delete_collection_basic, "lib.rs:7745") {
    VectorEngine engine;
    REQUIRE_OK(engine.create_collection("test", VectorCollectionConfig{}));
    REQUIRE_OK(engine.store_in_collection("test", "key1", {1.0f, 2.0f}));
    REQUIRE_OK(engine.store_in_collection("test", "key2", {3.0f, 4.0f}));
    REQUIRE(engine.collection_count("test") == 2);
    REQUIRE_OK(engine.delete_collection("test"));
    REQUIRE(!engine.collection_exists("test"));
    REQUIRE(engine.collection_count("test") == 0);
}
TEST_HOST(delete_collection_not_found, "lib.rs:7769") {
    VectorEngine engine;
    REQUIRE_ERR(engine.delete_collection("nonexistent"), CollectionNotFound);
}
TEST_HOST(list_collections_basic, "lib.rs:7777") {
    VectorEngine engine;
    REQUIRE_OK(engine.create_collection("alpha", VectorCollectionConfig{}));
    REQUIRE_OK(engine.create_collection("beta", VectorCollectionConfig{}));
    auto c = engine.list_collections();
    std::sort(c.begin(), c.end());
    REQUIRE(c == (std::vector<std::string>{"alpha", "beta"}));
}
TEST_HOST(store_in_collection_basic, "lib.rs:7793") {
    VectorEngine engine;
    REQUIRE_OK(engine.create_collection("products", VectorCollectionConfig{}));
    REQUIRE_OK(engine.store_in_collection("products", "item1", {1.0f, 2.0f, 3.0f}));
    REQUIRE(engine.get_from_collection("products", "item1").value() == (Vec{1.0f, 2.0f, 3.0f}));
}
TEST_HOST(store_in_collection_without_prior_create, "lib.rs:7808") {
    VectorEngine engine;
    REQUIRE_OK(engine.store_in_collection("auto_created", "key", {1.0f}));
    REQUIRE(engine.get_from_collection("auto_created", "key").value() == (Vec{1.0f}));
}
TEST_HOST(store_in_collection_dimension_constraint, "lib.rs:7821") {
    VectorEngine engine;
    REQUIRE_OK(engine.create_collection("fixed_dim", VectorCollectionConfig{}.with_dimension(3)));
    REQUIRE_OK(engine.store_in_collection("fixed_dim", "good", {1.0f, 2.0f, 3.0f}));
    auto r = engine.store_in_collection("fixed_dim", "bad", {1.0f, 2.0f});
    REQUIRE_ERR(r, DimensionMismatch);
    REQUIRE(r.error().expected == 3 && r.error().got == 2);
}
TEST_HOST(get_from_collection_not_found, "lib.rs:7843") {
    VectorEngine engine;
    REQUIRE_ERR(engine.get_from_collection("coll", "nonexistent"), NotFound);
}
TEST_HOST(delete_from_collection_basic, "lib.rs:7851") {
    VectorEngine engine;
    REQUIRE_OK(engine.store_in_collection("test", "key", {1.0f}));
    REQUIRE(engine.exists_in_collection("test", "key"));
    REQUIRE_OK(engine.delete_from_collection("test", "key"));
    REQUIRE(!engine.exists_in_collection("test", "key"));
}
TEST_HOST(delete_from_collection_not_found, "lib.rs:7865") {
    VectorEngine engine;
    REQUIRE_ERR(engine.delete_from_collection("coll", "nonexistent"), NotFound);
}
TEST_HOST(list_collection_keys_basic, "lib.rs:7873") {
    VectorEngine engine;
    REQUIRE_OK(engine.store_in_collection("test", "alpha", {1.0f}));
    REQUIRE_OK(engine.store_in_collection("test", "beta", {2.0f}));
    auto keys = engine.list_collection_keys("test");
    std::sort(keys.begin(), keys.end());
    REQUIRE(keys == (std::vector<std::string>{"alpha", "beta"}));
}
TEST_HOST(collection_count_basic, "lib.rs:7889") {
    VectorEngine engine;
    REQUIRE(engine.collection_count("empty") == 0);
    REQUIRE_OK(engine.store_in_collection("test", "a", {1.0f}));
    REQUIRE_OK(engine.store_in_collection("test", "b", {2.0f}));
    REQUIRE(engine.collection_count("test") == 2);
}
TEST_GPU(search_in_collection_basic, "lib.rs:7900") {
    VectorEngine engine;
    REQUIRE_OK(engine.store_in_collection("products", "p1", {1.0f, 0.0f, 0.0f}));
    REQUIRE_OK(engine.store_in_collection("products", "p2", {0.0f, 1.0f, 0.0f}));
    REQUIRE_OK(engine.store_in_collection("products", "p3", {0.0f, 0.0f, 1.0f}));
    auto results = engine.search_in_collection("products", {1.0f, 0.0f, 0.0f}, 2);
    REQUIRE_OK(results);
    REQUIRE(results.value().size() == 2);
    REQUIRE(results.value()[0].key == "p1");
}
TEST_HOST(search_in_collection_empty, "lib.rs:7922") {
    VectorEngine engine;
    auto results = engine.search_in_collection("empty", {1.0f, 2.0f}, 5);
    REQUIRE_OK(results);
    REQUIRE(results.value().empty());
}
TEST_HOST(search_in_collection_dimension_constraint, "lib.rs:7933") {
    VectorEngine engine;
    REQUIRE_OK(engine.create_collection("fixed", VectorCollectionConfig{}.with_dimension(3)));
    auto r = engine.search_in_collection("fixed", {1.0f, 2.0f}, 5);
    REQUIRE_ERR(r, DimensionMismatch);
    REQUIRE(r.error().expected == 3 && r.error().got == 2);
}
TEST_GPU(search_filtered_in_collection_basic, "lib.rs:7950") {
    VectorEngine engine;
    REQUIRE_OK(engine.store_in_collection_with_metadata("test", "item1", {1.0f, 0.0f}, meta({{"category", S("A")}})));
    REQUIRE_OK(engine.store_in_collection_with_metadata("test", "item2", {0.0f, 1.0f}, meta({{"category", S("B")}})));
    auto results = engine.search_filtered_in_collection("test", {1.0f, 0.0f}, 10, Cmp(Op::Eq, "category", S("A")));
    REQUIRE_OK(results);
    REQUIRE(results.value().size() == 1);
    REQUIRE(results.value()[0].key == "item1");
}
TEST_GPU(collection_isolation, "lib.rs:7982") {
    VectorEngine engine;
    REQUIRE_OK(engine.store_in_collection("coll_a", "key1", {1.0f}));
    REQUIRE_OK(engine.store_in_collection("coll_b", "key1", {2.0f}));
    REQUIRE(engine.get_from_collection("coll_a", "key1").value() == (Vec{1.0f}));
    REQUIRE(engine.get_from_collection("coll_b", "key1").value() == (Vec{2.0f}));
    auto a = engine.search_in_collection("coll_a", {1.0f}, 10), b = engine.search_in_collection("coll_b", {1.0f}, 10);
    REQUIRE_OK(a);
    REQUIRE_OK(b);
    REQUIRE(a.value().size() == 1 && b.value().size() == 1);
}
TEST_HOST(collection_and_default_isolation, "lib.rs:8008") {
    VectorEngine engine;
    REQUIRE_OK(engine.store_embedding("key1", {1.0f}));
    REQUIRE_OK(engine.store_in_collection("named", "key1", {2.0f}));
    REQUIRE(engine.get_embedding("key1").value() == (Vec{1.0f}));
    REQUIRE(engine.get_from_collection("named", "key1").value() == (Vec{2.0f}));
}
TEST_HOST(collection_config_with_dimension, "lib.rs:8028") {
    REQUIRE(VectorCollectionConfig{}.with_dimension(128).dimension == std::optional<size_t>(128));
}
TEST_HOST(collection_config_with_metric, "lib.rs:8034") {
    REQUIRE(VectorCollectionConfig{}.with_metric(DistanceMetric::Euclidean).distance_metric == DistanceMetric::Euclidean);
}
TEST_HOST(collection_config_with_auto_index, "lib.rs:8040") {
    auto c = VectorCollectionConfig{}.with_auto_index(500);
    REQUIRE(c.auto_index && c.auto_index_threshold == 500);
}
TEST_HOST(collection_config_default, "lib.rs:8047") {
    VectorCollectionConfig c;
    REQUIRE(!c.dimension.has_value() && c.distance_metric == DistanceMetric::Cosine && !c.auto_index &&
            c.auto_index_threshold == 1000);
}
TEST_HOST(collection_get_metadata_test, "lib.rs:9336") {
    VectorEngine engine;
    REQUIRE_OK(engine.create_collection("products", VectorCollectionConfig{}));
    REQUIRE_OK(engine.store_in_collection_with_metadata("products", "item1", {1.0f, 2.0f}, meta({{"price", I(100)}})));
    auto m = engine.get_collection_metadata("products", "item1");
    REQUIRE_OK(m);
    REQUIRE(m.value().count("price") == 1);
}
TEST_GPU(collection_search_filtered_test, "lib.rs:9356") {
    VectorEngine engine;
    REQUIRE_OK(engine.create_collection("items", VectorCollectionConfig{}));
    REQUIRE_OK(engine.store_in_collection_with_metadata("items", "item1", {1.0f, 0.0f}, meta({{"category", S("A")}})));
    REQUIRE_OK(engine.store_in_collection_with_metadata("items", "item2", {0.0f, 1.0f}, meta({{"category", S("B")}})));
    auto results = engine.search_filtered_in_collection("items", {1.0f, 0.0f}, 10, Cmp(Op::Eq, "category", S("A")));
    REQUIRE_OK(results);
    REQUIRE(results.value().size() == 1);
    REQUIRE(results.value()[0].key == "item1");
}
TEST_HOST(collection_exists_in_false, "lib.rs:9407") {
    VectorEngine engine;
    REQUIRE_OK(engine.create_collection("test", VectorCollectionConfig{}));
    REQUIRE(!engine.exists_in_collection("test", "nonexistent"));
}
TEST_HOST(collection_list_keys_empty, "lib.rs:9433") {
    VectorEngine engine;
    REQUIRE_OK(engine.create_collection("test", VectorCollectionConfig{}));
    REQUIRE(engine.list_collection_keys("test").empty());
}

// ---------------------------------------------------------------------------------------------
// search timeouts (lib.rs:8584-8690)
// ---------------------------------------------------------------------------------------------
TEST_HOST(config_with_search_timeout, "lib.rs:8584") {
    auto c = VectorEngineConfig{}.with_search_timeout(std::chrono::seconds(5));
    REQUIRE(c.search_timeout == std::optional<std::chrono::milliseconds>(std::chrono::milliseconds(5000)));
}
TEST_HOST(search_timeout_error_display, "lib.rs:8590") {
    VectorError e;
    e.kind = ErrorKind::SearchTimeout;
    e.operation = "search_similar";
    e.timeout_ms = 5000;
    REQUIRE(e.to_string().find("search_similar") != std::string::npos);
    REQUIRE(e.to_string().find("5000") != std::string::npos);
}
TEST_HOST(search_similar_respects_timeout, "lib.rs:8601") {
    auto engine = VectorEngine::with_config(VectorEngineConfig{}.with_search_timeout(std::chrono::nanoseconds(1)));
    REQUIRE_OK(engine);
    for (int i = 0; i < 1000; ++i) REQUIRE_OK(engine.value()->store_embedding(key_of("v", i), Vec(128, (float)i)));
    REQUIRE_ERR(engine.value()->search_similar(Vec(128, 0.5f), 10), SearchTimeout);
}
TEST_GPU(search_similar_no_timeout_when_none, "lib.rs:8616") {
    VectorEngine engine;
    for (int i = 0; i < 100; ++i) REQUIRE_OK(engine.store_embedding(key_of("v", i), {(float)i, 0.0f}));
    REQUIRE_OK(engine.search_similar({50.0f, 0.0f}, 10));
}
TEST_HOST(low_memory_config_has_timeout, "lib.rs:8648") {
    REQUIRE(VectorEngineConfig::low_memory().search_timeout ==
            std::optional<std::chrono::milliseconds>(std::chrono::milliseconds(30000)));
}
TEST_HOST(high_throughput_config_has_no_timeout, "lib.rs:8654") {
    REQUIRE(!VectorEngineConfig::high_throughput().search_timeout.has_value());
}
TEST_HOST(search_with_metric_respects_timeout, "lib.rs:8660") {
    auto engine = VectorEngine::with_config(VectorEngineConfig{}.with_search_timeout(std::chrono::nanoseconds(1)));
    REQUIRE_OK(engine);
    for (int i = 0; i < 1000; ++i) REQUIRE_OK(engine.value()->store_embedding(key_of("v", i), Vec(128, (float)i)));
    REQUIRE_ERR(engine.value()->search_similar_with_metric(Vec(128, 0.5f), 10, DistanceMetric::Cosine), SearchTimeout);
}
TEST_HOST(search_entities_respects_timeout, "lib.rs:8675") {
    auto engine = VectorEngine::with_config(VectorEngineConfig{}.with_search_timeout(std::chrono::nanoseconds(1)));
    REQUIRE_OK(engine);
    for (int i = 0; i < 1000; ++i) REQUIRE_OK(engine.value()->set_entity_embedding(key_of("entity:", i), Vec(128, (float)i)));
    REQUIRE_ERR(engine.value()->search_entities(Vec(128, 0.5f), 10), SearchTimeout);
}

// ---------------------------------------------------------------------------------------------
// later additions (lib.rs:9166-9570)
// ---------------------------------------------------------------------------------------------
TEST_HOST(batch_delete_multiple_keys, "lib.rs:9166") {
    VectorEngine engine;
    REQUIRE_OK(engine.store_embedding("a", {1.0f, 0.0f}));
    REQUIRE_OK(engine.store_embedding("b", {0.0f, 1.0f}));
    REQUIRE_OK(engine.store_embedding("c", {1.0f, 1.0f}));
    REQUIRE(engine.batch_delete_embeddings({"a", "b"}).value() == 2);
    REQUIRE(!engine.exists("a") && !engine.exists("b") && engine.exists("c"));
}
TEST_HOST(list_keys_paginated_no_limit_variant, "lib.rs:9216") {
    VectorEngine engine;
    for (int i = 0; i < 5; ++i) REQUIRE_OK(engine.store_embedding(key_of("key", i), {(float)i, 0.0f}));
    REQUIRE(engine.list_keys_paginated(Pg::skip_only(2)).items.size() == 3);
}
TEST_GPU(search_entities_paginated_no_count_variant, "lib.rs:9232") {
    VectorEngine engine;
    for (int i = 0; i < 5; ++i) REQUIRE_OK(engine.set_entity_embedding(key_of("entity:", i), {(float)i, 0.0f}));
    auto result = engine.search_entities_paginated({2.0f, 0.0f}, 5, Pg::with(0, 3));
    REQUIRE_OK(result);
    REQUIRE(result.value().items.size() == 3);
    REQUIRE(!result.value().total_count.has_value());
    REQUIRE(!result.value().has_more);
}
TEST_GPU(search_similar_with_metric_euclidean, "lib.rs:9482") {
    VectorEngine engine;
    REQUIRE_OK(engine.store_embedding("a", {0.0f, 0.0f}));
    REQUIRE_OK(engine.store_embedding("b", {1.0f, 0.0f}));
    REQUIRE_OK(engine.store_embedding("c", {3.0f, 4.0f}));
    auto results = engine.search_similar_with_metric({0.0f, 0.0f}, 3, DistanceMetric::Euclidean);
    REQUIRE_OK(results);
    REQUIRE(results.value().size() == 3);
    REQUIRE(results.value()[0].key == "a");
}
TEST_GPU(search_similar_with_metric_dot_product, "lib.rs:9498") {
    VectorEngine engine;
    REQUIRE_OK(engine.store_embedding("a", {1.0f, 0.0f}));
    REQUIRE_OK(engine.store_embedding("b", {0.0f, 1.0f}));
    auto results = engine.search_similar_with_metric({1.0f, 0.0f}, 2, DistanceMetric::DotProduct);
    REQUIRE_OK(results);
    REQUIRE(results.value().size() == 2);
}
TEST_HOST(list_keys_bounded_with_limit, "lib.rs:9528") {
    VectorEngineConfig config;
    config.max_keys_per_scan = 100;
    auto engine = VectorEngine::with_config(config);
    REQUIRE_OK(engine);
    for (int i = 0; i < 10; ++i) REQUIRE_OK(engine.value()->store_embedding(key_of("key", i), {(float)i}));
    REQUIRE(engine.value()->list_keys_bounded().size() == 10);
}
TEST_HOST(clear_all_embeddings, "lib.rs:9548") {
    VectorEngine engine;
    for (int i = 0; i < 5; ++i) REQUIRE_OK(engine.store_embedding(key_of("key", i), {(float)i}));
    REQUIRE(engine.count() == 5);
    REQUIRE(engine.clear().value() == 5);
    REQUIRE(engine.count() == 0);
}
TEST_HOST(clear_empty_engine, "lib.rs:9565") {
    VectorEngine engine;
    REQUIRE(engine.clear().value() == 0);
}

// ---------------------------------------------------------------------------------------------
// query_router: the SIMILAR operator and the EMBED command that feeds it
// (query_router/src/lib.rs tests; "QR" below).  Router = QueryRouter::new() + its own VectorEngine.
// ---------------------------------------------------------------------------------------------
struct Router {
    VectorEngine engine;
    QueryRouter router{engine};
    RouterOutcome execute(const std::string &c) { return router.execute(c); }
    RouterOutcome execute_parsed(const std::string &c) { return router.execute_parsed(c); }
};
#define REQUIRE_KIND(o, k) REQUIRE((o).ok && (o).result.kind == QueryResult::Kind::k)

TEST_HOST(routes_embed_to_vector, "QR:7656") {
    Router r;
    auto o = r.execute("EMBED doc1 [1.0, 0.0, 0.0]");
    REQUIRE_KIND(o, Empty);
    REQUIRE(r.engine.exists("doc1"));
}
TEST_GPU(routes_similar_to_vector, "QR:7669") {
    Router r;
    REQUIRE(r.execute("EMBED doc1 [1.0, 0.0, 0.0]").ok);
    REQUIRE(r.execute("EMBED doc2 [0.0, 1.0, 0.0]").ok);
    REQUIRE(r.execute("EMBED doc3 [0.9, 0.1, 0.0]").ok);
    auto o = r.execute("SIMILAR doc1 TOP 2");
    REQUIRE_KIND(o, Similar);
    REQUIRE(o.result.similar.size() == 2);
    REQUIRE(o.result.similar[0].key == "doc1");
}
TEST_HOST(handles_embedding_not_found, "QR:7815") {
    Router r;
    auto o = r.execute("SIMILAR nonexistent TOP 5");
    REQUIRE(!o.ok && o.error.kind == RouterError::Kind::VectorError);
}
TEST_GPU(embed_and_similar_inline, "QR:8051") {
    Router r;
    REQUIRE(r.execute("EMBED v1 [1.0, 0.0]").ok);
    REQUIRE(r.execute("EMBED v2 [0.0, 1.0]").ok);
    auto o = r.execute("SIMILAR [1.0, 0.0] TOP 1");
    REQUIRE_KIND(o, Similar);
    REQUIRE(o.result.similar.size() == 1 && o.result.similar[0].key == "v1");
}
TEST_GPU(similar_with_inline_vector, "QR:8467") {
    Router r;
    REQUIRE(r.execute("EMBED v1 [1.0, 0.0, 0.0]").ok);
    auto o = r.execute("SIMILAR [0.9, 0.1, 0.0] TOP 1");
    REQUIRE_KIND(o, Similar);
    REQUIRE(o.result.similar.size() == 1);
}
TEST_HOST(missing_embed_args, "QR:8537") {
    Router r;
    REQUIRE(!r.execute("EMBED").ok);
}
TEST_HOST(missing_similar_args, "QR:8544") {
    Router r;
    REQUIRE(!r.execute("SIMILAR").ok);
}
TEST_HOST(embed_with_empty_brackets, "QR:8858") {
    Router r;
    REQUIRE(!r.execute("EMBED emptykey []").ok);
}
TEST_HOST(similar_no_results, "QR:8948") {
    Router r;
    REQUIRE(!r.execute("SIMILAR nonexistent TOP 5").ok);
}
TEST_HOST(parsed_embed_store, "QR:9361") {
    Router r;
    REQUIRE_KIND(r.execute_parsed("EMBED STORE 'key1' [1.0, 2.0, 3.0]"), Empty);
}
TEST_HOST(parsed_embed_get, "QR:9370") {
    Router r;
    REQUIRE(r.execute("EMBED vec1 [1.0, 2.0, 3.0]").ok);
    auto o = r.execute_parsed("EMBED GET 'vec1'");
    REQUIRE_KIND(o, Value);
    REQUIRE(o.result.value == "[1.0, 2.0, 3.0]");  // format!("{vec:?}"); the reference asserts contains("1")
}
TEST_HOST(parsed_embed_delete, "QR:9382") {
    Router r;
    REQUIRE(r.execute("EMBED todelete [1.0, 2.0]").ok);
    auto o = r.execute_parsed("EMBED DELETE 'todelete'");
    REQUIRE_KIND(o, Count);
    REQUIRE(o.result.count == 1);
}
TEST_GPU(parsed_similar_by_key, "QR:9394") {
    Router r;
    REQUIRE(r.execute("EMBED item1 [1.0, 0.0, 0.0]").ok);
    REQUIRE(r.execute("EMBED item2 [0.9, 0.1, 0.0]").ok);
    auto o = r.execute_parsed("SIMILAR 'item1' LIMIT 5");
    REQUIRE_KIND(o, Similar);
    REQUIRE(!o.result.similar.empty());
}
TEST_GPU(parsed_similar_by_vector, "QR:9407") {
    Router r;
    REQUIRE(r.execute("EMBED vec1 [1.0, 0.0, 0.0]").ok);
    REQUIRE(r.execute("EMBED vec2 [0.0, 1.0, 0.0]").ok);
    auto o = r.execute_parsed("SIMILAR [1.0, 0.0, 0.0] LIMIT 5");
    REQUIRE_KIND(o, Similar);
    REQUIRE(!o.result.similar.empty());
}
TEST_GPU(parsed_similar_cosine_metric, "QR:9436") {
    Router r;
    REQUIRE(r.execute("EMBED cos_a [1.0, 0.0]").ok);
    REQUIRE(r.execute("EMBED cos_b [0.0, 1.0]").ok);
    REQUIRE(r.execute("EMBED cos_c [0.707, 0.707]").ok);
    // the metric comes AFTER the limit in the grammar: here `LIMIT 3` follows the statement and
    // is never read (parser::parse reads one statement) — the default limit of 10 applies
    auto o = r.execute_parsed("SIMILAR [1.0, 0.0] COSINE LIMIT 3");
    REQUIRE_KIND(o, Similar);
    REQUIRE(o.result.similar.size() == 3);
    REQUIRE(o.result.similar[0].key == "cos_a");
}
TEST_GPU(parsed_similar_euclidean_metric, "QR:9457") {
    Router r;
    REQUIRE(r.execute("EMBED euc_a [1.0, 0.0]").ok);
    REQUIRE(r.execute("EMBED euc_b [2.0, 0.0]").ok);
    REQUIRE(r.execute("EMBED euc_c [10.0, 0.0]").ok);
    auto o = r.execute_parsed("SIMILAR [1.0, 0.0] EUCLIDEAN LIMIT 3");
    REQUIRE_KIND(o, Similar);
    REQUIRE(o.result.similar.size() == 3);
    REQUIRE(o.result.similar[0].key == "euc_a");
}
TEST_GPU(parsed_similar_euclidean_zero_query, "QR:9478") {
    Router r;
    REQUIRE(r.execute("EMBED zero_origin [0.0, 0.0]").ok);
    REQUIRE(r.execute("EMBED zero_unit [1.0, 0.0]").ok);
    REQUIRE(r.execute("EMBED zero_far [10.0, 0.0]").ok);
    auto o = r.execute_parsed("SIMILAR [0.0, 0.0] EUCLIDEAN LIMIT 3");
    REQUIRE_KIND(o, Similar);
    REQUIRE(o.result.similar.size() == 3);
    REQUIRE(o.result.similar[0].key == "zero_origin");
    REQUIRE(std::fabs(o.result.similar[0].score - 1.0f) < 0.01f);
}
TEST_GPU(parsed_similar_dot_product_metric, "QR:9505") {
    Router r;
    REQUIRE(r.execute("EMBED dot_a [1.0, 0.0]").ok);
    REQUIRE(r.execute("EMBED dot_b [2.0, 0.0]").ok);
    REQUIRE(r.execute("EMBED dot_c [0.5, 0.0]").ok);
    auto o = r.execute_parsed("SIMILAR [1.0, 0.0] DOT_PRODUCT LIMIT 3");
    REQUIRE_KIND(o, Similar);
    REQUIRE(o.result.similar.size() == 3);
    REQUIRE(o.result.similar[0].key == "dot_b");
}
TEST_GPU(parsed_similar_with_limit_expr, "QR:10526") {
    Router r;
    REQUIRE(r.execute("EMBED v1 [1.0, 0.0]").ok);
    REQUIRE(r.execute("EMBED v2 [0.0, 1.0]").ok);
    REQUIRE_KIND(r.execute_parsed("SIMILAR 'v1' LIMIT 10"), Similar);
}
TEST_HOST(parsed_embed_store_with_list, "QR:10537") {
    Router r;
    REQUIRE(r.execute_parsed("EMBED STORE 'stored_vec' [1.0, 2.0, 3.0]").ok);
    REQUIRE_KIND(r.execute_parsed("EMBED GET 'stored_vec'"), Value);
}
TEST_HOST(parsed_embed_with_int_values, "QR:10635") {
    Router r;
    REQUIRE(r.execute_parsed("EMBED STORE 'intvec' [1, 2, 3]").ok);
    REQUIRE(r.engine.get_embedding("intvec").value() == (Vec{1.0f, 2.0f, 3.0f}));
}
TEST_HOST(parsed_similar_limit_not_integer, "QR:10709") {
    Router r;
    REQUIRE(r.execute("EMBED v [1.0, 2.0]").ok);
    REQUIRE(!r.execute_parsed("SIMILAR 'v' LIMIT 'ten'").ok);
}
TEST_HOST(parsed_embed_get_with_ident_key, "QR:10757") {
    Router r;
    REQUIRE(r.execute("EMBED mykey [1.0, 2.0]").ok);
    REQUIRE_KIND(r.execute_parsed("EMBED GET mykey"), Value);
}
TEST_GPU(parsed_similar_with_ident_key, "QR:10766") {
    Router r;
    REQUIRE(r.execute("EMBED vec1 [1.0, 0.0]").ok);
    REQUIRE_KIND(r.execute_parsed("SIMILAR vec1 LIMIT 5"), Similar);
}
TEST_HOST(parsed_embed_delete_nonexistent, "QR:10815") {
    Router r;
    REQUIRE(!r.execute_parsed("EMBED DELETE 'nonexistent'").ok);
}
TEST_HOST(parsed_embed_non_number_vector, "QR:10887") {
    Router r;
    REQUIRE(!r.execute_parsed("EMBED STORE 'k' ['a', 'b']").ok);
}
TEST_HOST(parsed_embed_batch_basic, "QR:12916") {
    Router r;
    auto o = r.execute_parsed("EMBED BATCH [('doc1', [1.0, 0.0]), ('doc2', [0.0, 1.0]), ('doc3', [0.5, 0.5])]");
    REQUIRE_KIND(o, Count);
    REQUIRE(o.result.count == 3);
    REQUIRE(r.execute_parsed("EMBED GET 'doc1'").ok);
}
TEST_HOST(parsed_embed_batch_empty, "QR:12935") {
    Router r;
    auto o = r.execute_parsed("EMBED BATCH []");
    REQUIRE_KIND(o, Count);
    REQUIRE(o.result.count == 0);
}
TEST_HOST(parsed_embed_store_into_collection, "QR:16287") {
    Router r;
    REQUIRE(r.execute_parsed("EMBED STORE 'doc1' [1.0, 2.0, 3.0] INTO my_collection").ok);
    REQUIRE_KIND(r.execute_parsed("EMBED GET 'doc1' INTO my_collection"), Value);
    REQUIRE(!r.engine.exists("doc1"));  // the default space is a different namespace
}
TEST_HOST(parsed_embed_delete_from_collection, "QR:16303") {
    Router r;
    REQUIRE(r.execute_parsed("EMBED STORE 'to_delete' [1.0, 2.0] INTO test_coll").ok);
    auto o = r.execute_parsed("EMBED DELETE 'to_delete' INTO test_coll");
    REQUIRE_KIND(o, Count);
    REQUIRE(o.result.count == 1);
    REQUIRE(!r.execute_parsed("EMBED GET 'to_delete' INTO test_coll").ok);
}
TEST_GPU(parsed_similar_into_collection, "QR:16323") {
    Router r;
    REQUIRE(r.execute_parsed("EMBED STORE 'vec_a' [1.0, 0.0, 0.0] INTO vectors").ok);
    REQUIRE(r.execute_parsed("EMBED STORE 'vec_b' [0.9, 0.1, 0.0] INTO vectors").ok);
    REQUIRE(r.execute_parsed("EMBED STORE 'vec_c' [0.0, 1.0, 0.0] INTO vectors").ok);
    auto o = r.execute_parsed("SIMILAR [1.0, 0.0, 0.0] LIMIT 3 INTO vectors");
    REQUIRE_KIND(o, Similar);
    REQUIRE(o.result.similar.size() == 3);
    REQUIRE(o.result.similar[0].key == "vec_a");
}
TEST_GPU(parsed_similar_with_where_clause, "QR:16378") {
    Router r;
    REQUIRE_OK(r.engine.store_embedding_with_metadata("item_a", {1.0f, 0.0f}, meta({{"category", S("science")}})));
    REQUIRE_OK(r.engine.store_embedding_with_metadata("item_b", {0.9f, 0.1f}, meta({{"category", S("tech")}})));
    REQUIRE_OK(r.engine.store_embedding_with_metadata("item_c", {0.8f, 0.2f}, meta({{"category", S("science")}})));
    auto o = r.execute_parsed("SIMILAR [1.0, 0.0] LIMIT 10 WHERE category = 'science'");
    REQUIRE_KIND(o, Similar);
    REQUIRE(o.result.similar.size() == 2);
    for (const auto &x : o.result.similar) REQUIRE(x.key == "item_a" || x.key == "item_c");
}
TEST_HOST(parsed_embed_batch_into_collection, "QR:16432") {
    Router r;
    auto o = r.execute_parsed("EMBED BATCH [('b1', [1.0, 2.0]), ('b2', [3.0, 4.0])] INTO batch_test");
    REQUIRE_KIND(o, Count);
    REQUIRE(o.result.count == 2);
    REQUIRE(r.execute_parsed("EMBED GET 'b1' INTO batch_test").ok);
    REQUIRE(r.execute_parsed("EMBED GET 'b2' INTO batch_test").ok);
}
TEST_GPU(test_similar_in_collection, "QR:19912") {
    // `COLLECTION 'grp'` and `TOP 2` are not part of either statement's grammar: the parser stops
    // after the vector, so c1 / c2 land in the default space and the search runs there
    Router r;
    REQUIRE(r.execute_parsed("EMBED STORE 'c1' [1.0, 0.0] COLLECTION 'grp'").ok);
    REQUIRE(r.execute_parsed("EMBED STORE 'c2' [0.9, 0.1] COLLECTION 'grp'").ok);
    auto o = r.execute_parsed("SIMILAR [1.0, 0.0] TOP 2 COLLECTION 'grp'");
    REQUIRE_KIND(o, Similar);
    REQUIRE(!o.result.similar.empty());
    REQUIRE(r.engine.exists("c1") && !r.engine.collection_exists("grp"));
}

// ---------------------------------------------------------------------------------------------
// gRPC PointsService::query (neumann_server/tests/grpc_vector_points.rs; "GP" below): the
// post-processing of points.rs:449-485 is VectorEngine::query_points here, the transport is not.
// ---------------------------------------------------------------------------------------------
static void setup_test_collection(VectorEngine &e, const std::string &name, size_t dimension) {  // GP:75-88
    e.create_collection(name, VectorCollectionConfig{}.with_dimension(dimension));
}
static void upsert_test_points(VectorEngine &e, const std::string &collection, size_t count, size_t dimension) {  // GP:91-112
    for (size_t i = 0; i < count; ++i) {
        Vec v(dimension);
        for (size_t j = 0; j < dimension; ++j) v[j] = (float)(i * 10 + j) / 10.0f;
        e.store_in_collection(collection, key_of("point_", i), v);
    }
}
TEST_GPU(test_points_query_basic, "GP:336") {
    VectorEngine engine;
    setup_test_collection(engine, "test_query", 3);
    upsert_test_points(engine, "test_query", 10, 3);
    auto r = engine.query_points("test_query", {0.5f, 1.0f, 1.5f}, 5, 0, std::nullopt, false);
    REQUIRE_OK(r);
    REQUIRE(r.value().size() <= 5 && !r.value().empty());
    for (size_t i = 1; i < r.value().size(); ++i) REQUIRE(r.value()[i - 1].score >= r.value()[i].score);
}
TEST_GPU(test_points_query_with_offset, "GP:371") {
    VectorEngine engine;
    setup_test_collection(engine, "test_query_offset", 3);
    upsert_test_points(engine, "test_query_offset", 10, 3);
    auto r = engine.query_points("test_query_offset", {0.0f, 0.0f, 0.0f}, 3, 2, std::nullopt, false);
    REQUIRE_OK(r);
    REQUIRE(r.value().size() <= 3);
    // with a query that has a direction: the page is ranks [2, 5) of the unpaged search
    auto full = engine.search_in_collection("test_query_offset", {0.5f, 1.0f, 1.5f}, 10);
    auto page = engine.query_points("test_query_offset", {0.5f, 1.0f, 1.5f}, 3, 2, std::nullopt, false);
    REQUIRE_OK(full);
    REQUIRE_OK(page);
    REQUIRE(page.value().size() == 3);
    for (size_t i = 0; i < 3; ++i) REQUIRE(page.value()[i].id == full.value()[i + 2].key);
}
TEST_GPU(test_points_query_with_score_threshold, "GP:402") {
    VectorEngine engine;
    setup_test_collection(engine, "test_query_threshold", 3);
    upsert_test_points(engine, "test_query_threshold", 10, 3);
    auto r = engine.query_points("test_query_threshold", {0.0f, 0.0f, 0.0f}, 10, 0, 0.8f, false);
    REQUIRE_OK(r);
    for (const auto &p : r.value()) REQUIRE(p.score >= 0.8f);
    // (a query with a direction: cos = 0.956, 0.951, 0.940, ... for point_0, point_1, point_2, ...)
    r = engine.query_points("test_query_threshold", {0.5f, 1.0f, 1.5f}, 10, 0, 0.945f, true);
    REQUIRE_OK(r);
    REQUIRE(r.value().size() == 2);
    for (const auto &p : r.value()) REQUIRE(p.score >= 0.945f && p.vector.size() == 3);
}
TEST_HOST(test_points_query_empty_collection, "GP:436") {
    VectorEngine engine;
    setup_test_collection(engine, "test_query_empty", 3);
    auto r = engine.query_points("test_query_empty", {1.0f, 2.0f, 3.0f}, 5, 0, std::nullopt, false);
    REQUIRE_OK(r);
    REQUIRE(r.value().empty());
}
TEST_HOST(test_points_query_missing_collection, "GP:604") {
    VectorEngine engine;
    auto r = engine.query_points("nonexistent_collection", {1.0f, 2.0f, 3.0f}, 5, 0, std::nullopt, false);
    REQUIRE_OK(r);  // "VectorEngine returns Ok with empty results for non-existent collections"
    REQUIRE(r.value().empty());
}
TEST_GPU(test_points_operation_after_collection_delete, "GP:629") {
    VectorEngine engine;
    setup_test_collection(engine, "test_deleted", 3);
    upsert_test_points(engine, "test_deleted", 5, 3);
    auto before = engine.query_points("test_deleted", {1.0f, 2.0f, 3.0f}, 5, 0, std::nullopt, false);
    REQUIRE_OK(before);
    REQUIRE(!before.value().empty());
    REQUIRE_OK(engine.delete_collection("test_deleted"));
    auto after = engine.query_points("test_deleted", {1.0f, 2.0f, 3.0f}, 5, 0, std::nullopt, false);
    REQUIRE_OK(after);
    REQUIRE(after.value().empty());
}
TEST_GPU(test_points_concurrent_query, "GP:901") {
    VectorEngine engine;
    setup_test_collection(engine, "test_concurrent_query", 3);
    upsert_test_points(engine, "test_concurrent_query", 20, 3);
    std::atomic<int> ok{0};
    std::vector<std::thread> ts;
    for (int i = 0; i < 5; ++i)
        ts.emplace_back([&, i] {
            auto r = engine.query_points("test_concurrent_query", {(float)i, (float)i + 1.0f, (float)i + 2.0f}, 5, 0,
                                         std::nullopt, false);
            if (r.is_ok() && r.value().size() == 5) ++ok;
        });
    for (auto &t : ts) t.join();
    REQUIRE(ok == 5);
}

// ---------------------------------------------------------------------------------------------
// tensor_store::hnsw::simd — the host restatement the zero-query short-circuit and
// compute_similarity use (tensor_store/src/hnsw.rs tests; "TS" below)
// ---------------------------------------------------------------------------------------------
TEST_HOST(test_simd_dot_product, "TS/hnsw.rs:2801") {
    const float a[4] = {1.0f, 2.0f, 3.0f, 4.0f}, b[4] = {1.0f, 1.0f, 1.0f, 1.0f};
    REQUIRE(std::fabs(simd::dot_product(a, b, 4) - 10.0f) < 1e-6f);
}
TEST_HOST(test_simd_magnitude, "TS/hnsw.rs:2809") {
    const float v[2] = {3.0f, 4.0f};
    REQUIRE(std::fabs(simd::magnitude(v, 2) - 5.0f) < 1e-6f);
}

}  // namespace

int main(int argc, char **argv) {
    const bool host_only = argc > 1 && std::strcmp(argv[1], "--host") == 0;
    const char *only = (argc > 2 && std::strcmp(argv[1], "--only") == 0) ? argv[2] : nullptr;
    int ran = 0, skipped = 0;
    for (const TestCase &t : registry()) {
        if (only && std::strcmp(only, t.name) != 0) continue;
        if (host_only && t.needs_device) {
            ++skipped;
            continue;
        }
        g_current = t.name;
        const int before = g_failures;
        t.fn();
        ++ran;
        if (g_failures != before) std::fprintf(stderr, "     ^ %s restates %s\n", t.name, t.src);
    }
    std::printf("reference_suite: %d tests run, %d skipped (need a device), %d failed\n", ran, skipped, g_failures);
    return g_failures ? 1 : 0;
}
