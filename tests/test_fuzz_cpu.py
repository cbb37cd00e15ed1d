"""Host-side robustness: token soups through the SIMILAR / EMBED router and the WHERE parser +
filter-program compiler.  Nothing may crash the process or leak a C++ exception across the C ABI —
every call either succeeds or comes back as a VectorError.  (No device involved: searches that get
as far as the scan fail with the documented storage error on a CPU-only host.)"""
import random

from neumann_b200 import engine as eng

ROUTER_TOKENS = ["SIMILAR", "EMBED", "STORE", "GET", "DELETE", "BATCH", "INTO", "WHERE", "LIMIT", "TOP", "COSINE",
                 "EUCLIDEAN", "DOT_PRODUCT", "CONNECTED", "TO", "[", "]", "(", ")", ",", "'a'", "'b", "\"c\"", "k1", "1",
                 "2.5", "-3", "1e30", "nan", "[1.0, 2.0]", "[1,2", "[]", "('a', [1,2])", "('a' [1])", "coll", "x = 1",
                 "AND", "OR", "EXISTS(x)", "CONTAINS(x,'y')", "IN (1,2)", "=", "<", "''", " ", "\t", "é"]
WHERE_TOKENS = ["x", "y", "price", "=", "!=", "<", "<=", ">", ">=", "AND", "OR", "(", ")", "1", "-2", "3.5", "'s'", "\"t\"",
                "true", "false", "null", "EXISTS(x)", "EXISTS(", "CONTAINS(x, 'a')", "CONTAINS(x,", "STARTS_WITH(y,'b')",
                "IN", "(1, 2, 'a')", ",", "TRUE", "NOT", "''", "1e999", "nan"]


def test_router_survives_token_soup():
    rng = random.Random(7)
    e = eng.VectorEngine()
    e.store_embedding("k1", [1.0, 2.0])
    outcomes = {"ok": 0, "err": 0}
    for _ in range(3000):
        cmd = " ".join(rng.choice(ROUTER_TOKENS) for _ in range(rng.randint(0, 9)))
        if rng.random() < 0.6:
            cmd = rng.choice(["SIMILAR ", "EMBED ", "EMBED STORE ", "EMBED BATCH ", "EMBED GET ", "similar "]) + cmd
        for fn in (e.execute, e.execute_parsed):
            try:
                fn(cmd)
                outcomes["ok"] += 1
            except eng.VectorError:
                outcomes["err"] += 1
    assert outcomes["ok"] > 0 and outcomes["err"] > 0
    e.store_embedding("after", [3.0])  # the engine is still usable
    assert e.exists("after")


def test_where_parser_and_filter_compiler_survive_token_soup():
    rng = random.Random(11)
    e = eng.VectorEngine()
    for i in range(50):
        e.store_embedding_with_metadata(f"k{i}", [1.0, float(i)], {"x": i, "y": f"b{i}", "price": i * 0.5})
    ok = err = 0
    for _ in range(3000):
        expr = " ".join(rng.choice(WHERE_TOKENS) for _ in range(rng.randint(0, 12)))
        try:
            n = e.count_matching(expr)
            assert 0 <= n <= 50
            ok += 1
        except eng.VectorError:
            err += 1
        try:
            e.debug_filter_program(2, expr)
        except eng.VectorError:
            pass
    assert ok > 0 and err > 0
    assert e.count_matching("x < 10") == 10
