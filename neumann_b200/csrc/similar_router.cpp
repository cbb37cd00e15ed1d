// similar_router.cpp — see similar_router.hpp.
#include "similar_router.hpp"

#include <algorithm>
#include <cctype>
#include <cerrno>
#include <charconv>
#include <cmath>
#include <cstdlib>

namespace neumann {
namespace {

std::string trim(const std::string &s) {
    size_t a = 0, b = s.size();
    while (a < b && std::isspace((unsigned char)s[a])) ++a;
    while (b > a && std::isspace((unsigned char)s[b - 1])) --b;
    return s.substr(a, b - a);
}
std::string upper(std::string s) {
    for (char &c : s) c = (char)std::toupper((unsigned char)c);
    return s;
}
RouterOutcome fail(RouterError::Kind k, std::string msg, int status = 6) {
    RouterOutcome o;
    o.ok = false;
    o.error.kind = k;
    o.error.message = std::move(msg);
    o.error.status = status;
    return o;
}
RouterOutcome from_vector_error(const VectorError &e) {
    return fail(RouterError::Kind::VectorError, e.to_string(), e.status());
}
RouterOutcome ok_empty() {
    RouterOutcome o;
    o.ok = true;
    return o;
}
RouterOutcome ok_similar(const std::vector<SearchResult> &rs) {
    RouterOutcome o;
    o.ok = true;
    o.result.kind = QueryResult::Kind::Similar;
    for (auto &r : rs) o.result.similar.push_back(SimilarResult{r.key, r.score});
    return o;
}

RouterOutcome ok_value(std::string v) {
    RouterOutcome o;
    o.ok = true;
    o.result.kind = QueryResult::Kind::Value;
    o.result.value = std::move(v);
    return o;
}
RouterOutcome ok_count(size_t n) {
    RouterOutcome o;
    o.ok = true;
    o.result.kind = QueryResult::Kind::Count;
    o.result.count = n;
    return o;
}

// One f32 the way Rust's `{:?}` prints it: the shortest digits that round-trip, plain decimals
// with at least one fractional digit for 1e-4 <= |x| < 1e16, `1e30` / `1e-18` outside.
std::string rust_debug_f32(float x) {
    if (std::isnan(x)) return "NaN";
    if (std::isinf(x)) return x < 0 ? "-inf" : "inf";
    char buf[64];
    const float ax = std::fabs(x);
    if (x == 0.0f || (ax >= 1e-4f && ax < 1e16f)) {
        auto r = std::to_chars(buf, buf + sizeof buf, x, std::chars_format::fixed);
        std::string s(buf, r.ptr);
        if (s.find('.') == std::string::npos) s += ".0";
        return s;
    }
    auto r = std::to_chars(buf, buf + sizeof buf, x, std::chars_format::scientific);
    std::string s(buf, r.ptr);  // d.ddde+XX
    const size_t e = s.find('e');
    std::string mant = s.substr(0, e), ex = s.substr(e + 1);
    const bool neg = !ex.empty() && ex[0] == '-';
    if (!ex.empty() && (ex[0] == '+' || ex[0] == '-')) ex.erase(ex.begin());
    while (ex.size() > 1 && ex[0] == '0') ex.erase(ex.begin());
    return mant + "e" + (neg ? "-" : "") + ex;
}
std::string rust_debug_vec(const std::vector<float> &v) {
    std::string s = "[";
    for (size_t i = 0; i < v.size(); ++i) s += (i ? ", " : "") + rust_debug_f32(v[i]);
    return s + "]";
}

// QR:6884-6901 parse_vector: strip brackets, split on ',', Rust f32::from_str per item.
bool parse_f32(const std::string &tok, float *out) {
    if (tok.empty()) return false;
    errno = 0;
    char *end = nullptr;
    float v = std::strtof(tok.c_str(), &end);
    if (end == tok.c_str() || *end != '\0') return false;
    *out = v;
    return true;
}
bool parse_vector(const std::string &text, std::vector<float> *out, std::string *why) {
    std::string s = trim(text);
    size_t a = 0, b = s.size();
    while (a < b && s[a] == '[') ++a;
    while (b > a && s[b - 1] == ']') --b;
    s = s.substr(a, b - a);
    out->clear();
    size_t pos = 0;
    while (true) {
        size_t c = s.find(',', pos);
        std::string tok = trim(s.substr(pos, c == std::string::npos ? std::string::npos : c - pos));
        float v;
        if (!parse_f32(tok, &v)) {
            *why = "Invalid float: " + tok;
            return false;
        }
        out->push_back(v);
        if (c == std::string::npos) break;
        pos = c + 1;
    }
    if (out->empty()) {
        *why = "Empty vector";
        return false;
    }
    return true;
}
bool parse_usize(const std::string &tok, size_t *out) {
    if (tok.empty()) return false;
    for (char c : tok)
        if (!std::isdigit((unsigned char)c)) return false;
    errno = 0;
    unsigned long long v = std::strtoull(tok.c_str(), nullptr, 10);
    if (errno) return false;
    *out = (size_t)v;
    return true;
}
// split "<first-word> <rest>" on the first whitespace run
void split_first(const std::string &s, std::string *first, std::string *rest) {
    size_t i = 0;
    while (i < s.size() && !std::isspace((unsigned char)s[i])) ++i;
    *first = s.substr(0, i);
    *rest = i < s.size() ? trim(s.substr(i)) : std::string();
}

// A key expression of the AST grammar at the front of *s: 'string', "string" or identifier.
// Consumes it; anything else (a number, a bracket) is not a key (QR expr_to_string).
bool take_key_expr(std::string *s, std::string *key) {
    *s = trim(*s);
    if (s->empty()) return false;
    const char c = (*s)[0];
    if (c == '\'' || c == '"') {
        const size_t e = s->find(c, 1);
        if (e == std::string::npos) return false;
        *key = s->substr(1, e - 1);
        *s = trim(s->substr(e + 1));
        return true;
    }
    if (!(std::isalpha((unsigned char)c) || c == '_')) return false;
    size_t i = 0;
    while (i < s->size() && (std::isalnum((unsigned char)(*s)[i]) || (*s)[i] == '_' || (*s)[i] == ':' ||
                             (*s)[i] == '.' || (*s)[i] == '-'))
        ++i;
    *key = s->substr(0, i);
    *s = trim(s->substr(i));
    return true;
}
// Optional `INTO ident` at the front of *s.  false = malformed (INTO without a name).
bool take_into(std::string *s, std::string *collection, bool *has) {
    *has = false;
    std::string kw, rest;
    split_first(trim(*s), &kw, &rest);
    if (upper(kw) != "INTO") return true;
    std::string name;
    split_first(rest, &name, s);
    if (name.empty()) return false;
    *collection = name;
    *has = true;
    return true;
}

}  // namespace

// ---- legacy path --------------------------------------------------------------------------
RouterOutcome QueryRouter::execute(const std::string &command_in) {
    const std::string command = trim(command_in);
    std::string word, rest;
    split_first(command, &word, &rest);
    const std::string op = upper(word);

    if (op == "EMBED") {
        // QR:6617-6630: EMBED <key> [<val>, ...]
        std::string key, vec_text;
        split_first(rest, &key, &vec_text);
        if (key.empty() || vec_text.empty())
            return fail(RouterError::Kind::MissingArgument, "key and vector");
        std::vector<float> v;
        std::string why;
        if (!parse_vector(vec_text, &v, &why)) return fail(RouterError::Kind::InvalidArgument, why);
        auto r = vector_.store_embedding(key, std::move(v));
        if (r.is_err()) return from_vector_error(r.error());
        return ok_empty();
    }
    if (op == "SIMILAR") {
        // QR:6632-6665 + parse_similar_args QR:6903-6929
        if (rest.empty()) return fail(RouterError::Kind::MissingArgument, "key or vector");
        size_t top_k = 10;
        std::string query_part = rest;
        const std::string up = upper(rest);
        size_t top_pos = up.find(" TOP ");
        if (top_pos != std::string::npos) {
            if (!parse_usize(trim(rest.substr(top_pos + 5)), &top_k))
                return fail(RouterError::Kind::InvalidArgument, "Invalid TOP value");
            query_part = trim(rest.substr(0, top_pos));
        }
        std::vector<float> q;
        if (!query_part.empty() && query_part[0] == '[') {
            std::string why;
            if (!parse_vector(query_part, &q, &why))
                return fail(RouterError::Kind::InvalidArgument, why);
        } else {
            std::string key = trim(query_part);
            while (!key.empty() && key.front() == '"') key.erase(key.begin());
            while (!key.empty() && key.back() == '"') key.pop_back();
            auto g = vector_.get_embedding(key);
            if (g.is_err()) return from_vector_error(g.error());
            q = g.value();
        }
        auto r = vector_.search_similar(q, top_k);
        if (r.is_err()) return from_vector_error(r.error());
        return ok_similar(r.value());
    }
    return fail(RouterError::Kind::UnknownCommand,
                "Unknown command: " + word + " (only EMBED and SIMILAR are served by this build)");
}

// ---- AST path -----------------------------------------------------------------------------
RouterOutcome QueryRouter::execute_parsed(const std::string &command_in) {
    const std::string command = trim(command_in);
    std::string word, rest;
    split_first(command, &word, &rest);
    const std::string op = upper(word);

    if (op == "EMBED") {
        // parser.rs:1777-1851 + exec_embed QR:5228-5311; a bare `EMBED key [..]` is the legacy form
        std::string w2, r2;
        split_first(rest, &w2, &r2);
        const std::string sub = upper(w2);
        std::string collection;
        bool has_collection = false;
        if (sub == "GET" || sub == "DELETE") {
            std::string key;
            if (!take_key_expr(&r2, &key)) return fail(RouterError::Kind::ParseError, "expected a key after EMBED " + sub);
            if (!take_into(&r2, &collection, &has_collection))
                return fail(RouterError::Kind::ParseError, "expected collection name after INTO");
            if (sub == "GET") {
                auto g = has_collection ? vector_.get_from_collection(collection, key) : vector_.get_embedding(key);
                if (g.is_err()) return from_vector_error(g.error());
                return ok_value(rust_debug_vec(g.value()));
            }
            auto d = has_collection ? vector_.delete_from_collection(collection, key) : vector_.delete_embedding(key);
            if (d.is_err()) return from_vector_error(d.error());
            return ok_count(1);
        }
        if (sub == "BATCH") {
            // EMBED BATCH [('key', [v, ...]), ...] [INTO coll]: items are stored in order, the
            // first failure ends the statement (what was stored stays stored, as in the reference)
            std::string t = trim(r2);
            if (t.empty() || t[0] != '[') return fail(RouterError::Kind::ParseError, "expected [ after EMBED BATCH");
            t = trim(t.substr(1));
            std::vector<std::pair<std::string, std::vector<float>>> items;
            while (!t.empty() && t[0] != ']') {
                if (t[0] != '(') return fail(RouterError::Kind::ParseError, "expected ( in EMBED BATCH");
                t = t.substr(1);
                std::string key;
                if (!take_key_expr(&t, &key)) return fail(RouterError::Kind::ParseError, "expected a key in EMBED BATCH");
                if (t.empty() || t[0] != ',') return fail(RouterError::Kind::ParseError, "expected , after the key");
                t = trim(t.substr(1));
                const size_t rb = t.find(']');
                if (t.empty() || t[0] != '[' || rb == std::string::npos)
                    return fail(RouterError::Kind::ParseError, "expected [vector] in EMBED BATCH");
                std::vector<float> v;
                std::string why;
                if (trim(t.substr(1, rb - 1)).empty()) v.clear();
                else if (!parse_vector(t.substr(0, rb + 1), &v, &why))
                    return fail(RouterError::Kind::InvalidArgument, why);
                t = trim(t.substr(rb + 1));
                if (t.empty() || t[0] != ')') return fail(RouterError::Kind::ParseError, "expected ) in EMBED BATCH");
                t = trim(t.substr(1));
                items.emplace_back(std::move(key), std::move(v));
                if (!t.empty() && t[0] == ',') t = trim(t.substr(1));
                else break;
            }
            if (t.empty() || t[0] != ']') return fail(RouterError::Kind::ParseError, "expected ] after EMBED BATCH items");
            t = trim(t.substr(1));
            if (!take_into(&t, &collection, &has_collection))
                return fail(RouterError::Kind::ParseError, "expected collection name after INTO");
            size_t count = 0;
            for (auto &kv : items) {
                auto r = has_collection ? vector_.store_in_collection(collection, kv.first, std::move(kv.second))
                                        : vector_.store_embedding(kv.first, std::move(kv.second));
                if (r.is_err()) return from_vector_error(r.error());
                ++count;
            }
            return ok_count(count);
        }
        if (sub != "STORE") return execute(command);
        std::string key;
        if (!take_key_expr(&r2, &key)) return fail(RouterError::Kind::ParseError, "expected a key after EMBED STORE");
        const size_t rb = r2.find(']');
        if (r2.empty() || r2[0] != '[' || rb == std::string::npos)
            return fail(RouterError::Kind::ParseError, "expected [vector]");
        std::vector<float> v;
        std::string why;
        if (!parse_vector(r2.substr(0, rb + 1), &v, &why))
            return fail(RouterError::Kind::InvalidArgument, why);
        std::string tail = trim(r2.substr(rb + 1));
        if (!take_into(&tail, &collection, &has_collection))
            return fail(RouterError::Kind::ParseError, "expected collection name after INTO");
        Result<Unit> r = has_collection ? vector_.store_in_collection(collection, key, std::move(v))
                                        : vector_.store_embedding(key, std::move(v));
        if (r.is_err()) return from_vector_error(r.error());
        return ok_empty();
    }
    if (op != "SIMILAR")
        return fail(RouterError::Kind::UnknownCommand,
                    "Unknown command: " + word + " (only EMBED and SIMILAR are served by this build)");
    if (rest.empty()) return fail(RouterError::Kind::ParseError, "expected key or vector");

    // query: '[' exprs ']' | expr
    bool inline_vec = false;
    std::vector<float> q;
    std::string key, tail;
    if (rest[0] == '[') {
        size_t rb = rest.find(']');
        if (rb == std::string::npos) return fail(RouterError::Kind::ParseError, "expected ]");
        std::string why;
        std::string body = trim(rest.substr(1, rb - 1));
        if (!body.empty() && !parse_vector(rest.substr(0, rb + 1), &q, &why))
            return fail(RouterError::Kind::InvalidArgument, why);
        inline_vec = true;
        tail = trim(rest.substr(rb + 1));
    } else if (rest[0] == '\'' || rest[0] == '"') {
        size_t e = rest.find(rest[0], 1);
        if (e == std::string::npos) return fail(RouterError::Kind::ParseError, "unterminated string");
        key = rest.substr(1, e - 1);
        tail = trim(rest.substr(e + 1));
    } else {
        split_first(rest, &key, &tail);
    }

    // clauses, in grammar order: [CONNECTED TO e] [LIMIT e] [metric] [INTO ident] [WHERE e]
    size_t top_k = 10;  // QR:5317-5322
    DistanceMetric metric = DistanceMetric::Cosine;
    std::string collection;
    bool has_collection = false;
    FilterCondition filter;
    bool has_filter = false;
    int stage = 0;  // 0 limit, 1 metric, 2 into, 3 done
    while (!tail.empty()) {
        std::string kw, r2;
        split_first(tail, &kw, &r2);
        const std::string K = upper(kw);
        if (K == "CONNECTED")
            return fail(RouterError::Kind::ParseError,
                        "SIMILAR ... CONNECTED TO is a cross-engine query (out of scope here)");
        if (K == "WHERE") {
            // the rest of the statement is the filter expression (QR:5370-5375)
            std::string why;
            if (!parse_where(r2, &filter, &why)) return fail(RouterError::Kind::ParseError, why);
            has_filter = true;
            tail.clear();
            break;
        }
        if (K == "LIMIT" && stage <= 0) {
            std::string n;
            split_first(r2, &n, &tail);
            if (!parse_usize(n, &top_k))
                return fail(RouterError::Kind::ParseError, "expected integer after LIMIT");
            stage = 1;
        } else if ((K == "COSINE" || K == "EUCLIDEAN" || K == "DOT_PRODUCT" || K == "DOTPRODUCT") &&
                   stage <= 1) {
            metric = K == "COSINE" ? DistanceMetric::Cosine
                                   : K == "EUCLIDEAN" ? DistanceMetric::Euclidean
                                                      : DistanceMetric::DotProduct;
            tail = r2;
            stage = 2;
        } else if (K == "INTO" && stage <= 2) {
            split_first(r2, &collection, &tail);
            if (collection.empty())
                return fail(RouterError::Kind::ParseError, "expected collection name after INTO");
            has_collection = true;
            stage = 3;
        } else {
            break;  // not part of this statement: parser::parse never looks at what follows it
        }
    }

    if (!inline_vec) {
        auto g = has_collection ? vector_.get_from_collection(collection, key)
                                : vector_.get_embedding(key);
        if (g.is_err()) return from_vector_error(g.error());
        q = g.value();
    }
    // QR:5385-5447: (collection, filter) -> search_filtered_in_collection; (collection) ->
    // search_in_collection (its own metric); (filter) -> search_similar_filtered; otherwise
    // search_similar_with_metric.
    auto r = has_collection
                 ? (has_filter ? vector_.search_filtered_in_collection(collection, q, top_k, filter)
                               : vector_.search_in_collection(collection, q, top_k))
                 : (has_filter ? vector_.search_similar_filtered(q, top_k, filter)
                               : vector_.search_similar_with_metric(q, top_k, metric));
    if (r.is_err()) return from_vector_error(r.error());
    return ok_similar(r.value());
}

}  // namespace neumann
