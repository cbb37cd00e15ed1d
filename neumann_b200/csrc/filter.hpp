// filter.hpp — metadata values and filter conditions of the reference's filtered search
// (vector_engine/src/lib.rs:291-445 types, :3580-3690 evaluation), plus the WHERE-expression
// parser the SIMILAR operator needs (query_router/src/lib.rs:5830-5906).
#pragma once
#include <cstdint>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

namespace neumann {

// ScalarValue as stored in an embedding's metadata (tensor_store ScalarValue subset).
struct MetadataValue {
    enum class Type { Null, Bool, Int, Float, String } type = Type::Null;
    bool b = false;
    int64_t i = 0;
    double f = 0.0;
    std::string s;
    static MetadataValue null() { return {}; }
    static MetadataValue boolean(bool v) { MetadataValue m; m.type = Type::Bool; m.b = v; return m; }
    static MetadataValue integer(int64_t v) { MetadataValue m; m.type = Type::Int; m.i = v; return m; }
    static MetadataValue real(double v) { MetadataValue m; m.type = Type::Float; m.f = v; return m; }
    static MetadataValue string(std::string v) { MetadataValue m; m.type = Type::String; m.s = std::move(v); return m; }
};
using Metadata = std::unordered_map<std::string, MetadataValue>;

// FilterValue (lib.rs:341-353)
using FilterValue = MetadataValue;

// FilterCondition (lib.rs:296-325)
struct FilterCondition {
    enum class Op { Eq, Ne, Lt, Le, Gt, Ge, And, Or, True, Exists, Contains, StartsWith, In } op = Op::True;
    std::string field;                 // all but And / Or / True
    FilterValue value;                 // comparisons; Contains/StartsWith use value.s
    std::vector<FilterValue> values;   // In
    std::shared_ptr<FilterCondition> lhs, rhs;  // And / Or

    static FilterCondition always() { return {}; }
    static FilterCondition cmp(Op op, std::string field, FilterValue v);
    static FilterCondition exists(std::string field);
    static FilterCondition contains(std::string field, std::string substr);
    static FilterCondition starts_with(std::string field, std::string prefix);
    static FilterCondition in(std::string field, std::vector<FilterValue> values);
    FilterCondition and_(FilterCondition other) const;
    FilterCondition or_(FilterCondition other) const;
};

// evaluate_filter (lib.rs:3590-3640): missing field or incompatible types -> false.
bool evaluate_filter(const Metadata &meta, const FilterCondition &f);

// FilterStrategy / FilteredSearchConfig (lib.rs:385-445)
enum class FilterStrategy : int { Auto = 0, PreFilter = 1, PostFilter = 2 };
struct FilteredSearchConfig {
    FilterStrategy strategy = FilterStrategy::Auto;
    float selectivity_threshold = 0.1f;
    size_t oversample_factor = 3;
    static FilteredSearchConfig pre_filter() {
        FilteredSearchConfig c;
        c.strategy = FilterStrategy::PreFilter;
        return c;
    }
    static FilteredSearchConfig post_filter() {
        FilteredSearchConfig c;
        c.strategy = FilterStrategy::PostFilter;
        return c;
    }
    FilteredSearchConfig with_oversample(size_t factor) const {
        FilteredSearchConfig c = *this;
        c.oversample_factor = factor;
        return c;
    }
};

// Parses `field op literal [AND|OR ...]` with parentheses (AND binds tighter than OR), the
// shape expr_to_filter_condition accepts (QR:5830-5881), plus EXISTS(f), CONTAINS(f,'s'),
// STARTS_WITH(f,'s') and `f IN (v, ...)` so the C API can express every FilterCondition.
// Literals: 123, 1.5, 'str' / "str", true/false, null (null -> String("null") as QR:5894),
// bare identifiers -> String (QR:5900).  Returns false and sets *error on a syntax error.
bool parse_where(const std::string &text, FilterCondition *out, std::string *error);

// Typed metadata wire format used by the C API: fields separated by 0x1f, each
// name 0x1e type 0x1e value, type one of i f s b n.
bool parse_metadata_wire(const std::string &wire, Metadata *out, std::string *error);

}  // namespace neumann
