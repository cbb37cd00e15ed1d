// filter.cpp — see filter.hpp.
#include "filter.hpp"

#include <cctype>
#include <cstdlib>

namespace neumann {

FilterCondition FilterCondition::cmp(Op op, std::string field, FilterValue v) {
    FilterCondition c;
    c.op = op;
    c.field = std::move(field);
    c.value = std::move(v);
    return c;
}
FilterCondition FilterCondition::exists(std::string field) {
    FilterCondition c;
    c.op = Op::Exists;
    c.field = std::move(field);
    return c;
}
FilterCondition FilterCondition::contains(std::string field, std::string substr) {
    FilterCondition c;
    c.op = Op::Contains;
    c.field = std::move(field);
    c.value = MetadataValue::string(std::move(substr));
    return c;
}
FilterCondition FilterCondition::starts_with(std::string field, std::string prefix) {
    FilterCondition c;
    c.op = Op::StartsWith;
    c.field = std::move(field);
    c.value = MetadataValue::string(std::move(prefix));
    return c;
}
FilterCondition FilterCondition::in(std::string field, std::vector<FilterValue> values) {
    FilterCondition c;
    c.op = Op::In;
    c.field = std::move(field);
    c.values = std::move(values);
    return c;
}
FilterCondition FilterCondition::and_(FilterCondition other) const {
    FilterCondition c;
    c.op = Op::And;
    c.lhs = std::make_shared<FilterCondition>(*this);
    c.rhs = std::make_shared<FilterCondition>(std::move(other));
    return c;
}
FilterCondition FilterCondition::or_(FilterCondition other) const {
    FilterCondition c;
    c.op = Op::Or;
    c.lhs = std::make_shared<FilterCondition>(*this);
    c.rhs = std::make_shared<FilterCondition>(std::move(other));
    return c;
}

namespace {
// compare_tensor_value_to_filter (lib.rs:3658-3684): -1/0/1, or 2 = incomparable
int compare(const MetadataValue &a, const FilterValue &b) {
    using T = MetadataValue::Type;
    auto ord = [](auto x, auto y) { return x < y ? -1 : (x > y ? 1 : 0); };
    if (a.type == T::Int && b.type == T::Int) return ord(a.i, b.i);
    if (a.type == T::Float && b.type == T::Float) {
        if (a.f != a.f || b.f != b.f) return 2;  // partial_cmp -> None on NaN
        return ord(a.f, b.f);
    }
    if (a.type == T::Float && b.type == T::Int) {
        if (a.f != a.f) return 2;
        return ord(a.f, (double)b.i);
    }
    if (a.type == T::Int && b.type == T::Float) {
        if (b.f != b.f) return 2;
        return ord((double)a.i, b.f);
    }
    if (a.type == T::String && b.type == T::String) return ord(a.s.compare(b.s), 0);
    if (a.type == T::Bool && b.type == T::Bool) return ord((int)a.b, (int)b.b);
    if (a.type == T::Null && b.type == T::Null) return 0;
    return 2;
}
bool compare_field(const Metadata &m, const std::string &field, const FilterValue &v,
                   bool (*pred)(int)) {
    auto it = m.find(field);
    if (it == m.end()) return false;
    int c = compare(it->second, v);
    return c != 2 && pred(c);
}
}  // namespace

bool evaluate_filter(const Metadata &m, const FilterCondition &f) {
    using Op = FilterCondition::Op;
    switch (f.op) {
    case Op::True: return true;
    case Op::And: return evaluate_filter(m, *f.lhs) && evaluate_filter(m, *f.rhs);
    case Op::Or: return evaluate_filter(m, *f.lhs) || evaluate_filter(m, *f.rhs);
    case Op::Exists: return m.count(f.field) != 0;
    case Op::Eq: return compare_field(m, f.field, f.value, [](int c) { return c == 0; });
    case Op::Ne: return compare_field(m, f.field, f.value, [](int c) { return c != 0; });
    case Op::Lt: return compare_field(m, f.field, f.value, [](int c) { return c < 0; });
    case Op::Le: return compare_field(m, f.field, f.value, [](int c) { return c <= 0; });
    case Op::Gt: return compare_field(m, f.field, f.value, [](int c) { return c > 0; });
    case Op::Ge: return compare_field(m, f.field, f.value, [](int c) { return c >= 0; });
    case Op::Contains: {
        auto it = m.find(f.field);
        return it != m.end() && it->second.type == MetadataValue::Type::String &&
               it->second.s.find(f.value.s) != std::string::npos;
    }
    case Op::StartsWith: {
        auto it = m.find(f.field);
        return it != m.end() && it->second.type == MetadataValue::Type::String &&
               it->second.s.compare(0, f.value.s.size(), f.value.s) == 0;
    }
    case Op::In:
        for (auto &v : f.values)
            if (compare_field(m, f.field, v, [](int c) { return c == 0; })) return true;
        return false;
    }
    return false;
}

// ---- WHERE parser ---------------------------------------------------------------------------
namespace {
struct Tok {
    enum K { End, Ident, Int, Float, Str, Op, LParen, RParen, Comma } k = End;
    std::string text;
};
struct Lexer {
    const std::string &s;
    size_t p = 0;
    std::string err;
    explicit Lexer(const std::string &str) : s(str) {}
    Tok next() {
        while (p < s.size() && std::isspace((unsigned char)s[p])) ++p;
        Tok t;
        if (p >= s.size()) return t;
        char c = s[p];
        if (c == '(') { ++p; t.k = Tok::LParen; return t; }
        if (c == ')') { ++p; t.k = Tok::RParen; return t; }
        if (c == ',') { ++p; t.k = Tok::Comma; return t; }
        if (c == '\'' || c == '"') {
            size_t e = s.find(c, p + 1);
            if (e == std::string::npos) { err = "unterminated string"; p = s.size(); return t; }
            t.k = Tok::Str;
            t.text = s.substr(p + 1, e - p - 1);
            p = e + 1;
            return t;
        }
        if (c == '=' || c == '!' || c == '<' || c == '>') {
            t.k = Tok::Op;
            t.text = std::string(1, c);
            ++p;
            if (p < s.size() && (s[p] == '=' || (c == '<' && s[p] == '>'))) t.text += s[p++];
            return t;
        }
        if (std::isdigit((unsigned char)c) || ((c == '-' || c == '+') && p + 1 < s.size() &&
                                               (std::isdigit((unsigned char)s[p + 1]) || s[p + 1] == '.'))) {
            size_t b = p++;
            bool is_float = false;
            while (p < s.size() && (std::isdigit((unsigned char)s[p]) || s[p] == '.' || s[p] == 'e' ||
                                    s[p] == 'E' || ((s[p] == '-' || s[p] == '+') &&
                                                    (s[p - 1] == 'e' || s[p - 1] == 'E')))) {
                if (s[p] == '.' || s[p] == 'e' || s[p] == 'E') is_float = true;
                ++p;
            }
            t.k = is_float ? Tok::Float : Tok::Int;
            t.text = s.substr(b, p - b);
            return t;
        }
        if (std::isalpha((unsigned char)c) || c == '_') {
            size_t b = p;
            while (p < s.size() && (std::isalnum((unsigned char)s[p]) || s[p] == '_' || s[p] == '.' ||
                                    s[p] == ':'))
                ++p;
            t.k = Tok::Ident;
            t.text = s.substr(b, p - b);
            return t;
        }
        err = std::string("unexpected character '") + c + "'";
        p = s.size();
        return t;
    }
};
std::string upper(std::string v) {
    for (char &c : v) c = (char)std::toupper((unsigned char)c);
    return v;
}
struct Parser {
    Lexer lx;
    Tok cur;
    std::string err;
    explicit Parser(const std::string &s) : lx(s) { advance(); }
    void advance() {
        cur = lx.next();
        if (!lx.err.empty() && err.empty()) err = lx.err;
    }
    bool is_kw(const char *kw) const { return cur.k == Tok::Ident && upper(cur.text) == kw; }
    bool value(FilterValue *out) {
        switch (cur.k) {
        case Tok::Int: *out = MetadataValue::integer(std::strtoll(cur.text.c_str(), nullptr, 10)); break;
        case Tok::Float: *out = MetadataValue::real(std::strtod(cur.text.c_str(), nullptr)); break;
        case Tok::Str: *out = MetadataValue::string(cur.text); break;
        case Tok::Ident: {
            std::string u = upper(cur.text);
            if (u == "TRUE") *out = MetadataValue::boolean(true);
            else if (u == "FALSE") *out = MetadataValue::boolean(false);
            else if (u == "NULL") *out = MetadataValue::string("null");  // QR:5894
            else *out = MetadataValue::string(cur.text);                 // QR:5900
            break;
        }
        default: err = "expected a literal"; return false;
        }
        advance();
        return true;
    }
    bool primary(FilterCondition *out) {
        if (cur.k == Tok::LParen) {
            advance();
            if (!expr(out)) return false;
            if (cur.k != Tok::RParen) { err = "expected )"; return false; }
            advance();
            return true;
        }
        if (cur.k != Tok::Ident) { err = "expected a field name"; return false; }
        std::string name = cur.text, U = upper(name);
        advance();
        if ((U == "EXISTS" || U == "CONTAINS" || U == "STARTS_WITH") && cur.k == Tok::LParen) {
            advance();
            if (cur.k != Tok::Ident && cur.k != Tok::Str) { err = "expected a field name"; return false; }
            std::string field = cur.text;
            advance();
            if (U == "EXISTS") {
                *out = FilterCondition::exists(field);
            } else {
                if (cur.k != Tok::Comma) { err = "expected ,"; return false; }
                advance();
                if (cur.k != Tok::Str) { err = "expected a string"; return false; }
                *out = U == "CONTAINS" ? FilterCondition::contains(field, cur.text)
                                       : FilterCondition::starts_with(field, cur.text);
                advance();
            }
            if (cur.k != Tok::RParen) { err = "expected )"; return false; }
            advance();
            return true;
        }
        if (is_kw("IN")) {
            advance();
            if (cur.k != Tok::LParen) { err = "expected ( after IN"; return false; }
            advance();
            std::vector<FilterValue> vals;
            while (cur.k != Tok::RParen) {
                FilterValue v;
                if (!value(&v)) return false;
                vals.push_back(std::move(v));
                if (cur.k == Tok::Comma) advance();
                else if (cur.k != Tok::RParen) { err = "expected , or )"; return false; }
            }
            advance();
            *out = FilterCondition::in(name, std::move(vals));
            return true;
        }
        if (cur.k != Tok::Op) { err = "expected a comparison operator after " + name; return false; }
        std::string op = cur.text;
        advance();
        FilterValue v;
        if (!value(&v)) return false;
        FilterCondition::Op o;
        if (op == "=" || op == "==") o = FilterCondition::Op::Eq;
        else if (op == "!=" || op == "<>") o = FilterCondition::Op::Ne;
        else if (op == "<") o = FilterCondition::Op::Lt;
        else if (op == "<=") o = FilterCondition::Op::Le;
        else if (op == ">") o = FilterCondition::Op::Gt;
        else if (op == ">=") o = FilterCondition::Op::Ge;
        else { err = "Unsupported operator in filter condition: " + op; return false; }
        *out = FilterCondition::cmp(o, name, std::move(v));
        return true;
    }
    bool conj(FilterCondition *out) {
        if (!primary(out)) return false;
        while (is_kw("AND")) {
            advance();
            FilterCondition r;
            if (!primary(&r)) return false;
            *out = out->and_(std::move(r));
        }
        return true;
    }
    bool expr(FilterCondition *out) {
        if (!conj(out)) return false;
        while (is_kw("OR")) {
            advance();
            FilterCondition r;
            if (!conj(&r)) return false;
            *out = out->or_(std::move(r));
        }
        return true;
    }
};
}  // namespace

bool parse_where(const std::string &text, FilterCondition *out, std::string *error) {
    Parser p(text);
    bool ok = p.expr(out) && p.err.empty();
    if (ok && p.cur.k != Tok::End) {
        ok = false;
        p.err = "unexpected trailing tokens in filter condition";
    }
    if (!ok && error) *error = p.err.empty() ? "Expected binary expression in filter condition" : p.err;
    return ok;
}

bool parse_metadata_wire(const std::string &wire, Metadata *out, std::string *error) {
    out->clear();
    size_t pos = 0;
    while (pos < wire.size()) {
        size_t end = wire.find('\x1f', pos);
        std::string rec = wire.substr(pos, end == std::string::npos ? std::string::npos : end - pos);
        pos = end == std::string::npos ? wire.size() : end + 1;
        if (rec.empty()) continue;
        size_t a = rec.find('\x1e'), b = a == std::string::npos ? a : rec.find('\x1e', a + 1);
        if (a == std::string::npos || b == std::string::npos || b != a + 2) {
            if (error) *error = "malformed metadata record";
            return false;
        }
        std::string name = rec.substr(0, a), val = rec.substr(b + 1);
        switch (rec[a + 1]) {
        case 'i': (*out)[name] = MetadataValue::integer(std::strtoll(val.c_str(), nullptr, 10)); break;
        case 'f': (*out)[name] = MetadataValue::real(std::strtod(val.c_str(), nullptr)); break;
        case 's': (*out)[name] = MetadataValue::string(val); break;
        case 'b': (*out)[name] = MetadataValue::boolean(val == "1"); break;
        case 'n': (*out)[name] = MetadataValue::null(); break;
        default:
            if (error) *error = "unknown metadata type tag";
            return false;
        }
    }
    return true;
}

}  // namespace neumann
