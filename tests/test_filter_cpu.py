"""CPU tests of the filtered-search host logic: metadata wire format, filter evaluation rules
(vector_engine/src/lib.rs:3590-3690) and the WHERE parser (query_router/src/lib.rs:5830-5906)."""
import pytest

from neumann_b200 import engine as eng


def setup_filtered_search_engine():
    # vector_engine/src/lib.rs:6968-7000
    e = eng.VectorEngine()
    for i, (cat, price) in enumerate(zip(["electronics", "clothing", "food"], [100, 50, 25])):
        e.store_embedding_with_metadata(f"item{i}", [float(i + 1), 1.0, 1.0],
                                        {"category": cat, "price": price, "active": i % 2 == 0})
    return e


@pytest.mark.parametrize("where,count", [
    ("category = 'electronics'", 1),                    # search_filtered_eq_string :7004
    ("price = 50", 1),                                  # :7020
    ("price > 30", 2), ("price < 60", 2),               # :7033, :7045
    ("price <= 50", 2), ("price >= 50", 2),             # :7057, :7070
    ("price > 30 AND price < 80", 1),                   # :7083
    ("category = 'electronics' OR category = 'food'", 2),  # :7098
    ("TRUE", 3),                                        # :7117
    ("EXISTS(category)", 3), ("EXISTS(missing)", 0),    # :7129
    ("CONTAINS(category, 'cloth')", 1), ("CONTAINS(price, '5')", 0),   # :7155, :7186
    ("STARTS_WITH(category, 'foo')", 1), ("STARTS_WITH(active, 't')", 0),  # :7208, :7239
    ("missing = 1", 0),                                 # :7261
    ("price IN (25, 100, 7)", 2),                       # :7277
    ("category != 'food'", 2),                          # :7295
    ("active = true", 2), ("active = false", 1),        # :7310, :7647
    ("price = 50.0", 1), ("price < 50.5", 2),           # int field vs float filter :7544
    ("price = 'fifty'", 0), ("category > 5", 0),        # incompatible types :7679
    ("category < 'd'", 1), ("category >= 'electronics'", 2),  # string ordering :7604
    ("(category = 'food' OR category = 'clothing') AND active = false", 1),
    ("category = food", 1),                             # bare identifier -> String (QR:5900)
])
def test_filter_evaluation_rules(where, count):
    assert setup_filtered_search_engine().count_matching(where) == count


def test_float_and_null_metadata():
    e = eng.VectorEngine()
    e.store_embedding_with_metadata("a", [1.0, 0.0], {"score": 0.75, "nothing": None})
    e.store_embedding_with_metadata("b", [0.0, 1.0], {"score": 2})
    assert e.count_matching("score > 0.5") == 2       # float field / int field vs float filter
    assert e.count_matching("score < 1") == 1         # float field vs int filter (:7522)
    assert e.count_matching("EXISTS(nothing)") == 1
    assert e.count_matching("nothing = null") == 0    # NULL literal becomes String("null") (QR:5894)
    e.store_embedding("a", [1.0, 0.0])                # overwriting drops the old metadata
    assert e.count_matching("EXISTS(score)") == 1


@pytest.mark.parametrize("bad", ["price ~ 3", "price >", "AND price = 1", "price = 1 AND",
                                 "(price = 1", "price = 'x", "CONTAINS(category)", "price IN 1"])
def test_where_parse_errors(bad):
    with pytest.raises(eng.VectorError) as ei:
        setup_filtered_search_engine().count_matching(bad)
    assert ei.value.kind == "InvalidArgument"


def test_filtered_search_validation_without_device():
    e = setup_filtered_search_engine()
    with pytest.raises(eng.VectorError) as ei:
        e.search_similar_filtered([], 5, "TRUE")              # :7369
    assert ei.value.kind == "EmptyVector"
    with pytest.raises(eng.VectorError) as ei:
        e.search_similar_filtered([1.0, 1.0, 1.0], 0, "TRUE")  # :7377
    assert ei.value.kind == "InvalidTopK"
    e2 = eng.VectorEngine(max_dimension=2)
    with pytest.raises(eng.VectorError) as ei:
        e2.search_similar_filtered([1.0, 1.0, 1.0], 3, "TRUE")  # :7385
    assert ei.value.kind == "DimensionMismatch"
    # a zero query never reaches the device (the filter itself is evaluated ON the device now:
    # tests/test_gpu_filter.py covers the empty-subset case)
    assert e.search_similar_filtered([0.0, 0.0, 0.0], 5, "price > 0", eng.PRE_FILTER) == []
    with pytest.raises(eng.VectorError):
        e.execute_parsed("SIMILAR [1.0, 1.0, 1.0] LIMIT 2 WHERE price ~ 3")
