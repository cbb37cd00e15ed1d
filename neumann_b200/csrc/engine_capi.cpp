// engine_capi.cpp — extern "C" view of VectorEngine / QueryRouter (include/neumann_b200_engine.h).
#include <cctype>
#include <cstring>
#include <memory>
#include <optional>
#include <string>

#include "../../include/neumann_b200.h"
#include "../../include/neumann_b200_engine.h"
#include "similar_router.hpp"
#include "vector_engine.hpp"

using namespace neumann;

struct nm_engine {
    std::unique_ptr<VectorEngine> engine;
    std::unique_ptr<QueryRouter> router;
};
struct nm_results {
    std::vector<SearchResult> hits;
};

namespace {
thread_local std::string g_engine_error;
int fail(const VectorError &e) {
    g_engine_error = e.to_string();
    return e.status();
}
int fail(int code, const std::string &msg) {
    g_engine_error = msg;
    return code;
}
int give(Result<std::vector<SearchResult>> r, nm_results **out) {
    if (out) *out = nullptr;
    if (r.is_err()) return fail(r.error());
    if (out) {
        auto *res = new nm_results();
        res->hits = std::move(r.value());
        *out = res;
    }
    return NM_OK;
}
int give(const RouterOutcome &o, nm_results **out) {
    if (out) *out = nullptr;
    if (!o.ok) return fail(o.error.status, o.error.message);
    if (out && o.result.kind == QueryResult::Kind::Similar) {
        auto *res = new nm_results();
        for (auto &s : o.result.similar) res->hits.push_back(SearchResult{s.key, s.score});
        *out = res;
    }
    return NM_OK;
}
// No C++ exception may cross the C ABI (a host in another language cannot unwind it): whatever
// escapes the mirror — std::bad_alloc from its containers first of all — comes back as
// NM_ERR_STORAGE with the message in nm_engine_last_error().
template <class F>
int guarded(F &&body) noexcept {
    try {
        return body();
    } catch (const std::bad_alloc &) {
        try {
            g_engine_error = "out of host memory";
        } catch (...) {
        }
        return NM_ERR_STORAGE;
    } catch (const std::exception &e) {
        try {
            g_engine_error = std::string("unexpected exception: ") + e.what();
        } catch (...) {
        }
        return NM_ERR_STORAGE;
    } catch (...) {
        return NM_ERR_STORAGE;
    }
}
}  // namespace

extern "C" {

void nm_engine_config_default(nm_engine_config *cfg) {
    if (!cfg) return;
    std::memset(cfg, 0, sizeof(*cfg));
    cfg->sparse_threshold = 0.5f;
    cfg->parallel_threshold = 5000;
    cfg->default_metric = NM_COSINE;
    cfg->search_timeout_ms = -1;
}

int nm_engine_create(const nm_engine_config *cfg, nm_engine **out) {
    return guarded([&]() -> int {
        if (!out) return fail(NM_ERR_INVALID_ARGUMENT, "null out pointer");
        *out = nullptr;
        VectorEngineConfig c;
        if (cfg) {
            if (cfg->default_dimension) c.default_dimension = (size_t)cfg->default_dimension;
            c.sparse_threshold = cfg->sparse_threshold;
            c.parallel_threshold = (size_t)cfg->parallel_threshold;
            c.default_metric = (DistanceMetric)cfg->default_metric;
            if (cfg->max_dimension) c.max_dimension = (size_t)cfg->max_dimension;
            if (cfg->search_timeout_ms >= 0)
                c.search_timeout = std::chrono::milliseconds(cfg->search_timeout_ms);
            for (int i = 0; i < cfg->n_devices && i < 8; ++i) c.devices.push_back(cfg->devices[i]);
            c.device_prefilter = cfg->device_prefilter != 0;
            if (cfg->max_keys_per_scan) c.max_keys_per_scan = (size_t)cfg->max_keys_per_scan;
        }
        auto r = VectorEngine::with_config(std::move(c));
        if (r.is_err()) return fail(r.error());
        auto *e = new nm_engine();
        e->engine = std::move(r.value());
        e->router.reset(new QueryRouter(*e->engine));
        *out = e;
        return NM_OK;
    });
}

void nm_engine_destroy(nm_engine *e) { delete e; }
const char *nm_engine_last_error(void) { return g_engine_error.c_str(); }

int nm_engine_store_embedding(nm_engine *e, const char *key, const float *vec, size_t n) {
    return guarded([&]() -> int {
        if (!e || !key) return fail(NM_ERR_INVALID_ARGUMENT, "null argument");
        auto r = e->engine->store_embedding(key, std::vector<float>(vec, vec + n));
        return r.is_err() ? fail(r.error()) : NM_OK;
    });
}

int nm_engine_get_embedding(nm_engine *e, const char *key, float *out, size_t cap, size_t *len) {
    return guarded([&]() -> int {
        if (!e || !key) return fail(NM_ERR_INVALID_ARGUMENT, "null argument");
        auto r = e->engine->get_embedding(key);
        if (r.is_err()) return fail(r.error());
        if (len) *len = r.value().size();
        if (out && cap >= r.value().size())
            std::memcpy(out, r.value().data(), r.value().size() * sizeof(float));
        return NM_OK;
    });
}

int nm_engine_delete_embedding(nm_engine *e, const char *key) {
    return guarded([&]() -> int {
        if (!e || !key) return fail(NM_ERR_INVALID_ARGUMENT, "null argument");
        auto r = e->engine->delete_embedding(key);
        return r.is_err() ? fail(r.error()) : NM_OK;
    });
}

int nm_engine_exists(nm_engine *e, const char *key) { return e && key && e->engine->exists(key); }
uint64_t nm_engine_count(nm_engine *e) { return e ? e->engine->count() : 0; }

int nm_engine_search_similar(nm_engine *e, const float *query, size_t n, size_t top_k,
                             nm_results **out) {
    return guarded([&]() -> int {
        if (!e) return fail(NM_ERR_INVALID_ARGUMENT, "null engine");
        std::vector<float> q(query, query + (query ? n : 0));
        return give(e->engine->search_similar(q, top_k), out);
    });
}

int nm_engine_search_similar_with_metric(nm_engine *e, const float *query, size_t n, size_t top_k,
                                         int metric, nm_results **out) {
    return guarded([&]() -> int {
        if (!e) return fail(NM_ERR_INVALID_ARGUMENT, "null engine");
        if (metric < 0 || metric > 2) return fail(NM_ERR_INVALID_ARGUMENT, "unknown metric");
        std::vector<float> q(query, query + (query ? n : 0));
        return give(e->engine->search_similar_with_metric(q, top_k, (DistanceMetric)metric), out);
    });
}

int nm_engine_search_similar_batch(nm_engine *e, const float *queries, size_t nq, size_t n,
                                   size_t top_k, int metric, nm_results **out) {
    return guarded([&]() -> int {
        if (!e || !out) return fail(NM_ERR_INVALID_ARGUMENT, "null argument");
        if (metric < 0 || metric > 2) return fail(NM_ERR_INVALID_ARGUMENT, "unknown metric");
        for (size_t i = 0; i < nq; ++i) out[i] = nullptr;
        std::vector<std::vector<float>> qs(nq);
        for (size_t i = 0; i < nq; ++i)
            qs[i].assign(queries + i * n, queries + (queries ? (i + 1) * n : i * n));
        auto r = e->engine->search_similar_batch(qs, top_k, (DistanceMetric)metric);
        if (r.is_err()) return fail(r.error());
        for (size_t i = 0; i < nq; ++i) {
            auto *res = new nm_results();
            res->hits = std::move(r.value()[i]);
            out[i] = res;
        }
        return NM_OK;
    });
}

int nm_engine_compute_similarity(const float *a, size_t na, const float *b, size_t nb,
                                 float *out) {
    return guarded([&]() -> int {
        auto r = VectorEngine::compute_similarity(std::vector<float>(a, a + (a ? na : 0)),
                                                  std::vector<float>(b, b + (b ? nb : 0)));
        if (r.is_err()) return fail(r.error());
        if (out) *out = r.value();
        return NM_OK;
    });
}

int nm_engine_create_collection(nm_engine *e, const char *name, uint64_t dimension, int metric) {
    return guarded([&]() -> int {
        if (!e || !name) return fail(NM_ERR_INVALID_ARGUMENT, "null argument");
        VectorCollectionConfig c;
        if (dimension) c.dimension = (size_t)dimension;
        c.distance_metric = (DistanceMetric)metric;
        auto r = e->engine->create_collection(name, c);
        return r.is_err() ? fail(r.error()) : NM_OK;
    });
}

int nm_engine_delete_collection(nm_engine *e, const char *name) {
    return guarded([&]() -> int {
        if (!e || !name) return fail(NM_ERR_INVALID_ARGUMENT, "null argument");
        auto r = e->engine->delete_collection(name);
        return r.is_err() ? fail(r.error()) : NM_OK;
    });
}

int nm_engine_collection_exists(nm_engine *e, const char *name) {
    return guarded([&]() -> int {
        return e && name && e->engine->collection_exists(name);
    });
}

int nm_engine_store_in_collection(nm_engine *e, const char *collection, const char *key,
                                  const float *vec, size_t n) {
    return guarded([&]() -> int {
        if (!e || !collection || !key) return fail(NM_ERR_INVALID_ARGUMENT, "null argument");
        auto r = e->engine->store_in_collection(collection, key, std::vector<float>(vec, vec + n));
        return r.is_err() ? fail(r.error()) : NM_OK;
    });
}

int nm_engine_delete_from_collection(nm_engine *e, const char *collection, const char *key) {
    return guarded([&]() -> int {
        if (!e || !collection || !key) return fail(NM_ERR_INVALID_ARGUMENT, "null argument");
        auto r = e->engine->delete_from_collection(collection, key);
        return r.is_err() ? fail(r.error()) : NM_OK;
    });
}

uint64_t nm_engine_collection_count(nm_engine *e, const char *collection) {
    return (e && collection) ? e->engine->collection_count(collection) : 0;
}

int nm_engine_search_in_collection(nm_engine *e, const char *collection, const float *query,
                                   size_t n, size_t top_k, nm_results **out) {
    return guarded([&]() -> int {
        if (!e || !collection) return fail(NM_ERR_INVALID_ARGUMENT, "null argument");
        std::vector<float> q(query, query + (query ? n : 0));
        return give(e->engine->search_in_collection(collection, q, top_k), out);
    });
}

namespace {
bool parse_filter_arg(const char *where_expr, FilterCondition *f) {
    std::string text = where_expr ? where_expr : "";
    std::string up;
    for (char c : text)
        if (!isspace((unsigned char)c)) up += (char)toupper((unsigned char)c);
    if (up.empty() || up == "TRUE") {
        *f = FilterCondition::always();
        return true;
    }
    std::string why;
    if (!parse_where(text, f, &why)) {
        fail(NM_ERR_INVALID_ARGUMENT, "Parse error: " + why);
        return false;
    }
    return true;
}
FilteredSearchConfig filter_config(int strategy, size_t oversample) {
    FilteredSearchConfig c;
    c.strategy = (FilterStrategy)strategy;
    if (oversample) c.oversample_factor = oversample;
    return c;
}
}  // namespace

int nm_engine_store_embedding_with_metadata(nm_engine *e, const char *key, const float *vec,
                                            size_t n, const char *metadata) {
    return guarded([&]() -> int {
        if (!e || !key) return fail(NM_ERR_INVALID_ARGUMENT, "null argument");
        Metadata m;
        std::string why;
        if (!parse_metadata_wire(metadata ? metadata : "", &m, &why))
            return fail(NM_ERR_INVALID_ARGUMENT, why);
        auto r = e->engine->store_embedding_with_metadata(key, std::vector<float>(vec, vec + n),
                                                          std::move(m));
        return r.is_err() ? fail(r.error()) : NM_OK;
    });
}

int nm_engine_store_in_collection_with_metadata(nm_engine *e, const char *collection,
                                                const char *key, const float *vec, size_t n,
                                                const char *metadata) {
    return guarded([&]() -> int {
        if (!e || !collection || !key) return fail(NM_ERR_INVALID_ARGUMENT, "null argument");
        Metadata m;
        std::string why;
        if (!parse_metadata_wire(metadata ? metadata : "", &m, &why))
            return fail(NM_ERR_INVALID_ARGUMENT, why);
        auto r = e->engine->store_in_collection_with_metadata(collection, key,
                                                              std::vector<float>(vec, vec + n), std::move(m));
        return r.is_err() ? fail(r.error()) : NM_OK;
    });
}

int nm_engine_search_similar_filtered(nm_engine *e, const float *query, size_t n, size_t top_k,
                                      const char *where_expr, int strategy,
                                      size_t oversample_factor, nm_results **out) {
    return guarded([&]() -> int {
        if (!e) return fail(NM_ERR_INVALID_ARGUMENT, "null engine");
        if (out) *out = nullptr;
        FilterCondition f;
        if (!parse_filter_arg(where_expr, &f)) return NM_ERR_INVALID_ARGUMENT;
        std::vector<float> q(query, query + (query ? n : 0));
        return give(e->engine->search_similar_filtered(q, top_k, f,
                                                       filter_config(strategy, oversample_factor)),
                    out);
    });
}

int nm_engine_search_filtered_in_collection(nm_engine *e, const char *collection,
                                            const float *query, size_t n, size_t top_k,
                                            const char *where_expr, int strategy,
                                            size_t oversample_factor, nm_results **out) {
    return guarded([&]() -> int {
        if (!e || !collection) return fail(NM_ERR_INVALID_ARGUMENT, "null argument");
        if (out) *out = nullptr;
        FilterCondition f;
        if (!parse_filter_arg(where_expr, &f)) return NM_ERR_INVALID_ARGUMENT;
        std::vector<float> q(query, query + (query ? n : 0));
        return give(e->engine->search_filtered_in_collection(collection, q, top_k, f,
                                                             filter_config(strategy, oversample_factor)),
                    out);
    });
}

int nm_engine_count_matching(nm_engine *e, const char *where_expr, uint64_t *out) {
    return guarded([&]() -> int {
        if (!e || !out) return fail(NM_ERR_INVALID_ARGUMENT, "null argument");
        FilterCondition f;
        if (!parse_filter_arg(where_expr, &f)) return NM_ERR_INVALID_ARGUMENT;
        *out = e->engine->count_matching(f);
        return NM_OK;
    });
}

int nm_engine_update_metadata(nm_engine *e, const char *key, const char *metadata_wire) {
    return guarded([&]() -> int {
        if (!e || !key) return fail(NM_ERR_INVALID_ARGUMENT, "null argument");
        Metadata m;
        std::string perr;
        if (!parse_metadata_wire(metadata_wire ? metadata_wire : "", &m, &perr))
            return fail(NM_ERR_INVALID_ARGUMENT, perr);
        auto r = e->engine->update_metadata(key, m);
        return r.is_err() ? fail(r.error()) : NM_OK;
    });
}

int nm_engine_remove_metadata_field(nm_engine *e, const char *key, const char *field) {
    return guarded([&]() -> int {
        if (!e || !key || !field) return fail(NM_ERR_INVALID_ARGUMENT, "null argument");
        auto r = e->engine->remove_metadata_field(key, field);
        return r.is_err() ? fail(r.error()) : NM_OK;
    });
}

int nm_engine_has_metadata_field(nm_engine *e, const char *key, const char *field) {
    return guarded([&]() -> int {
        return e && key && field && e->engine->has_metadata_field(key, field);
    });
}

int nm_engine_clear(nm_engine *e, uint64_t *out) {
    return guarded([&]() -> int {
        if (!e) return fail(NM_ERR_INVALID_ARGUMENT, "null argument");
        auto r = e->engine->clear();
        if (r.is_err()) return fail(r.error());
        if (out) *out = r.value();
        return NM_OK;
    });
}

int nm_engine_batch_delete_embeddings(nm_engine *e, const char *keys_wire, uint64_t *out) {
    return guarded([&]() -> int {
        if (!e || !keys_wire) return fail(NM_ERR_INVALID_ARGUMENT, "null argument");
        std::vector<std::string> keys;
        std::string w(keys_wire);
        size_t pos = 0;
        while (pos <= w.size()) {
            size_t end = w.find('\x1f', pos);
            if (end == std::string::npos) end = w.size();
            if (end > pos) keys.push_back(w.substr(pos, end - pos));
            pos = end + 1;
        }
        auto r = e->engine->batch_delete_embeddings(keys);
        if (r.is_err()) return fail(r.error());
        if (out) *out = r.value();
        return NM_OK;
    });
}

int nm_engine_search_paginated(nm_engine *e, int entities, const float *query, size_t n, size_t top_k,
                               size_t skip, int64_t limit, int count_total, nm_results **out,
                               uint64_t *total_count, int *has_more) {
    return guarded([&]() -> int {
        if (!e) return fail(NM_ERR_INVALID_ARGUMENT, "null argument");
        if (out) *out = nullptr;
        std::vector<float> q(query, query + (query ? n : 0));
        VectorEngine::Pagination pg;
        pg.skip = skip;
        if (limit >= 0) pg.limit = (size_t)limit;
        pg.count_total = count_total != 0;
        auto r = entities ? e->engine->search_entities_paginated(q, top_k, pg)
                          : e->engine->search_similar_paginated(q, top_k, pg);
        if (r.is_err()) return fail(r.error());
        if (total_count) *total_count = r.value().total_count ? (uint64_t)*r.value().total_count : UINT64_MAX;
        if (has_more) *has_more = r.value().has_more ? 1 : 0;
        if (out) {
            auto *res = new nm_results();
            res->hits = std::move(r.value().items);
            *out = res;
        }
        return NM_OK;
    });
}

int nm_engine_debug_filter_program(nm_engine *e, uint32_t dim, const char *where_expr, char *out,
                                   size_t out_cap, size_t *out_len) {
    return guarded([&]() -> int {
        if (!e || !out_len) return fail(NM_ERR_INVALID_ARGUMENT, "null argument");
        FilterCondition f;
        if (!parse_filter_arg(where_expr, &f)) return NM_ERR_INVALID_ARGUMENT;
        const std::string js = e->engine->debug_filter_program(dim, f);
        *out_len = js.size();
        if (out && out_cap) {
            const size_t m = std::min(out_cap - 1, js.size());
            std::memcpy(out, js.data(), m);
            out[m] = 0;
        }
        return NM_OK;
    });
}

int nm_engine_query_points(nm_engine *e, const char *collection, const float *vector, size_t n,
                           size_t limit, size_t offset, int has_threshold, float score_threshold,
                           nm_results **out) {
    return guarded([&]() -> int {
        if (!e || !collection) return fail(NM_ERR_INVALID_ARGUMENT, "null argument");
        if (out) *out = nullptr;
        std::vector<float> q(vector, vector + (vector ? n : 0));
        std::optional<float> thr;
        if (has_threshold) thr = score_threshold;
        auto r = e->engine->query_points(collection, q, limit, offset, thr, false);
        if (r.is_err()) return fail(r.error());
        if (out) {
            auto *res = new nm_results();
            for (auto &p : r.value()) res->hits.push_back(SearchResult{p.id, p.score});
            *out = res;
        }
        return NM_OK;
    });
}

int nm_engine_set_entity_embedding(nm_engine *e, const char *entity_key, const float *vec, size_t n) {
    return guarded([&]() -> int {
        if (!e || !entity_key) return fail(NM_ERR_INVALID_ARGUMENT, "null argument");
        auto r = e->engine->set_entity_embedding(entity_key, std::vector<float>(vec, vec + n));
        return r.is_err() ? fail(r.error()) : NM_OK;
    });
}

int nm_engine_remove_entity_embedding(nm_engine *e, const char *entity_key) {
    return guarded([&]() -> int {
        if (!e || !entity_key) return fail(NM_ERR_INVALID_ARGUMENT, "null argument");
        auto r = e->engine->remove_entity_embedding(entity_key);
        return r.is_err() ? fail(r.error()) : NM_OK;
    });
}

int nm_engine_entity_has_embedding(nm_engine *e, const char *entity_key) {
    return guarded([&]() -> int {
        return e && entity_key && e->engine->entity_has_embedding(entity_key);
    });
}

int nm_engine_search_entities(nm_engine *e, const float *query, size_t n, size_t top_k,
                              nm_results **out) {
    return guarded([&]() -> int {
        if (!e) return fail(NM_ERR_INVALID_ARGUMENT, "null engine");
        std::vector<float> q(query, query + (query ? n : 0));
        return give(e->engine->search_entities(q, top_k), out);
    });
}

int nm_engine_execute(nm_engine *e, const char *command, nm_results **out) {
    return guarded([&]() -> int {
        if (!e || !command) return fail(NM_ERR_INVALID_ARGUMENT, "null argument");
        return give(e->router->execute(command), out);
    });
}

int nm_engine_execute_parsed(nm_engine *e, const char *command, nm_results **out) {
    return guarded([&]() -> int {
        if (!e || !command) return fail(NM_ERR_INVALID_ARGUMENT, "null argument");
        return give(e->router->execute_parsed(command), out);
    });
}

int nm_engine_mirror_rows(nm_engine *e, uint32_t dim, uint64_t *host_rows, uint64_t *device_rows) {
    return guarded([&]() -> int {
        if (!e) return fail(NM_ERR_INVALID_ARGUMENT, "null engine");
        for (auto &m : e->engine->mirror_info())
            if (m.dim == dim) {
                if (host_rows) *host_rows = m.host_rows;
                if (device_rows) *device_rows = m.device_rows;
                return NM_OK;
            }
        if (host_rows) *host_rows = 0;
        if (device_rows) *device_rows = 0;
        return NM_OK;
    });
}

size_t nm_results_len(const nm_results *r) { return r ? r->hits.size() : 0; }
const char *nm_results_key(const nm_results *r, size_t i) {
    return (r && i < r->hits.size()) ? r->hits[i].key.c_str() : "";
}
float nm_results_score(const nm_results *r, size_t i) {
    return (r && i < r->hits.size()) ? r->hits[i].score : 0.0f;
}
void nm_results_free(nm_results *r) { delete r; }

}  // extern "C"
