"""Device-side filter evaluation (filter_kernels.cuh, nm_index_column_set / nm_search_filtered):
the FilterCondition tree is compiled to a postfix program and evaluated over typed metadata
columns ON THE DEVICE into the row bitmask of the scan.  Checked against an independent Python
restatement of the reference's rules (evaluate_filter / compare_tensor_value_to_filter,
vector_engine/src/lib.rs:3592-3684) and against the oracle on the eligible subset
(search_with_pre_filter, lib.rs:3514-3557)."""
import math
import struct

import numpy as np
import pytest

import oracle_ffi as o
from neumann_b200 import DeviceIndex, NmError, _ffi
from neumann_b200 import engine as eng
from neumann_b200._ffi import (NM_C_EQ, NM_C_GE, NM_C_GT, NM_C_LE, NM_C_LT, NM_C_NE, NM_F_AND,
                               NM_F_CMP, NM_F_EXISTS, NM_F_FALSE, NM_F_OR, NM_F_STR_TABLE, NM_F_TRUE,
                               NM_V_BOOL, NM_V_FLOAT, NM_V_INT, NM_V_MISSING, NM_V_NULL, NM_V_STRING,
                               NmFilterOp)

pytestmark = pytest.mark.gpu

MISSING = object()


def f64_bits(x: float) -> int:
    return struct.unpack("<Q", struct.pack("<d", x))[0]


def encode_column(values, dictionary):
    """python values (MISSING / None / bool / int / float / str) -> (tags u8, vals u64)."""
    tags = np.zeros(len(values), np.uint8)
    vals = np.zeros(len(values), np.uint64)
    for i, v in enumerate(values):
        if v is MISSING:
            tags[i] = NM_V_MISSING
        elif v is None:
            tags[i] = NM_V_NULL
        elif isinstance(v, bool):
            tags[i], vals[i] = NM_V_BOOL, int(v)
        elif isinstance(v, int):
            tags[i], vals[i] = NM_V_INT, np.uint64(v & 0xFFFFFFFFFFFFFFFF)
        elif isinstance(v, float):
            tags[i], vals[i] = NM_V_FLOAT, f64_bits(v)
        else:
            tags[i], vals[i] = NM_V_STRING, dictionary.setdefault(v, len(dictionary))
    return tags, vals


def ref_compare(a, b):
    """compare_tensor_value_to_filter (lib.rs:3658-3684): -1/0/1 or None (incomparable)."""
    def ordf(x, y):
        if math.isnan(x) or math.isnan(y):
            return None
        return (x > y) - (x < y)
    if isinstance(a, bool) or isinstance(b, bool):
        if isinstance(a, bool) and isinstance(b, bool):
            return (a > b) - (a < b)
        return None
    if a is None or b is None:
        return 0 if (a is None and b is None) else None
    if isinstance(a, int) and isinstance(b, int):
        return (a > b) - (a < b)
    if isinstance(a, (int, float)) and isinstance(b, (int, float)):
        return ordf(float(a), float(b))
    if isinstance(a, str) and isinstance(b, str):
        ab, bb = a.encode(), b.encode()
        return (ab > bb) - (ab < bb)
    return None


CMP = {NM_C_EQ: lambda c: c == 0, NM_C_NE: lambda c: c != 0, NM_C_LT: lambda c: c < 0,
       NM_C_LE: lambda c: c <= 0, NM_C_GT: lambda c: c > 0, NM_C_GE: lambda c: c >= 0}


def ref_cmp_field(v, lit, cmp):
    if v is MISSING:
        return False
    c = ref_compare(v, lit)
    return c is not None and CMP[cmp](c)


def cmp_op(column, cmp, lit):
    op = NmFilterOp(kind=NM_F_CMP, cmp=cmp, column=column)
    if lit is None:
        op.lit_tag = NM_V_NULL
    elif isinstance(lit, bool):
        op.lit_tag, op.lit = NM_V_BOOL, int(lit)
    elif isinstance(lit, int):
        op.lit_tag, op.lit = NM_V_INT, lit & 0xFFFFFFFFFFFFFFFF
    else:
        op.lit_tag, op.lit = NM_V_FLOAT, f64_bits(lit)
    return op


def table_op(column, dictionary, pred, tables):
    """NM_F_STR_TABLE leaf: the string predicate evaluated once per distinct string."""
    bits = np.zeros((len(dictionary) + 31) // 32, np.uint32)
    for s, code in dictionary.items():
        if pred(s):
            bits[code // 32] |= np.uint32(1 << (code % 32))
    op = NmFilterOp(kind=NM_F_STR_TABLE, column=column, table_off=len(tables), table_bits=len(dictionary))
    tables.extend(int(x) for x in bits)
    return op


def make_columns(n, seed=5):
    rng = np.random.default_rng(seed)
    num, txt, flag = [], [], []
    words = [f"w{j:03d}" for j in range(60)] + ["", "zebra", "Zebra", "éclair"]
    for i in range(n):
        r = rng.integers(0, 12)
        if r == 0:
            num.append(MISSING)
        elif r == 1:
            num.append(None)
        elif r == 2:
            num.append(float("nan"))
        elif r == 3:
            num.append(bool(i & 1))
        elif r < 8:
            num.append(int(rng.integers(-50, 50)))
        elif r == 8:
            num.append(int(rng.integers(-2**62, 2**62)))
        else:
            num.append(float(rng.integers(-100, 100)) / 2.0)
        txt.append(MISSING if rng.integers(0, 9) == 0 else
                   (int(rng.integers(0, 5)) if rng.integers(0, 15) == 0 else words[int(rng.integers(0, len(words)))]))
        flag.append([MISSING, True, False, None][int(rng.integers(0, 4))])
    return num, txt, flag


def test_filter_program_follows_the_reference_type_rules():
    n, d = 70_001, 16
    idx = DeviceIndex(d)
    idx.fill_synthetic(n, 3)
    num, txt, flag = make_columns(n)
    dic = {}
    idx.column_set(1, 0, *encode_column(num, dic))
    tdic = {}
    idx.column_set(2, 0, *encode_column(txt, tdic))
    idx.column_set(7, 0, *encode_column(flag, {}))
    cases = []
    for cmp in CMP:
        for lit in (0, 7, -3, 2**61, 0.5, -12.0, float("nan"), True, False, None):
            cases.append(([cmp_op(1, cmp, lit)], None,
                          [ref_cmp_field(v, lit, cmp) for v in num], f"num cmp{cmp} {lit!r}"))
        cases.append(([cmp_op(7, cmp, True)], None, [ref_cmp_field(v, True, cmp) for v in flag], f"flag cmp{cmp}"))
        cases.append(([cmp_op(7, cmp, None)], None, [ref_cmp_field(v, None, cmp) for v in flag], f"null cmp{cmp}"))
        tables = []
        op = table_op(2, tdic, lambda s, cmp=cmp: CMP[cmp](ref_compare(s, "w030")), tables)
        cases.append(([op], tables, [isinstance(v, str) and CMP[cmp](ref_compare(v, "w030")) for v in txt],
                      f"str cmp{cmp}"))
    cases.append(([NmFilterOp(kind=NM_F_EXISTS, column=1)], None, [v is not MISSING for v in num], "exists"))
    cases.append(([NmFilterOp(kind=NM_F_EXISTS, column=99)], None, [False] * n, "exists on an unknown column"))
    cases.append(([cmp_op(99, NM_C_NE, 1)], None, [False] * n, "!= on an unknown column is false"))
    cases.append(([NmFilterOp(kind=NM_F_TRUE)], None, [True] * n, "true"))
    cases.append(([NmFilterOp(kind=NM_F_FALSE)], None, [False] * n, "false"))
    tables = []
    contains = table_op(2, tdic, lambda s: "eb" in s, tables)
    starts = table_op(2, tdic, lambda s: s.startswith("w01"), tables)
    # (num > -5 AND num <= 20.5) OR (txt CONTAINS 'eb' AND NOT-missing flag) OR txt STARTS_WITH 'w01'
    prog = [cmp_op(1, NM_C_GT, -5), cmp_op(1, NM_C_LE, 20.5), NmFilterOp(kind=NM_F_AND),
            contains, NmFilterOp(kind=NM_F_EXISTS, column=7), NmFilterOp(kind=NM_F_AND),
            NmFilterOp(kind=NM_F_OR), starts, NmFilterOp(kind=NM_F_OR)]
    want = [(ref_cmp_field(a, -5, NM_C_GT) and ref_cmp_field(a, 20.5, NM_C_LE)) or
            (isinstance(t, str) and "eb" in t and f is not MISSING) or
            (isinstance(t, str) and t.startswith("w01")) for a, t, f in zip(num, txt, flag)]
    cases.append((prog, tables, want, "compound"))
    for ops, tabs, want, what in cases:
        got = idx.filter_mask(ops, tabs)
        assert np.array_equal(got, np.asarray(want, bool)), what
    # malformed programs are rejected before anything is launched
    for bad in ([NmFilterOp(kind=NM_F_AND)], [NmFilterOp(kind=NM_F_TRUE)] * 2, [NmFilterOp(kind=9)],
                [NmFilterOp(kind=NM_F_STR_TABLE, column=2, table_off=5, table_bits=64)],
                [NmFilterOp(kind=NM_F_CMP, column=1, lit_tag=NM_V_STRING)]):
        with pytest.raises(NmError):
            idx.filter_mask(bad, None)
    idx.close()


@pytest.mark.parametrize("n", [1, 31, 32, 255, 256, 257, 511, 512, 2048, 8193])
def test_filter_mask_at_the_edges_of_a_mask_block(n):
    """A warp evaluates one 256-row block of the mask (8 words); whole blocks take the unguarded
    loads, the last partial one the guarded ones; TRUE must not leak past the last row."""
    idx = DeviceIndex(8)
    idx.fill_synthetic(n, 11)
    rng = np.random.default_rng(n)
    a = rng.integers(-20, 20, n)
    tags = np.where(rng.integers(0, 7, n) == 0, NM_V_MISSING, NM_V_INT).astype(np.uint8)
    idx.column_set(3, 0, tags, a.astype(np.int64).view(np.uint64))
    have = tags == NM_V_INT
    cases = [([NmFilterOp(kind=NM_F_TRUE)], np.ones(n, bool)),
             ([cmp_op(3, NM_C_LT, 5)], have & (a < 5)),
             ([cmp_op(3, NM_C_NE, 0)], have & (a != 0)),
             ([cmp_op(3, NM_C_GE, -3), NmFilterOp(kind=NM_F_EXISTS, column=3), NmFilterOp(kind=NM_F_AND),
               NmFilterOp(kind=NM_F_FALSE), NmFilterOp(kind=NM_F_OR)], have & (a >= -3))]
    # a right-leaning chain 64 deep: every leaf is pushed before the first AND pops anything
    deep = [cmp_op(3, NM_C_GT, -20 + (j % 3)) for j in range(64)] + [NmFilterOp(kind=NM_F_AND)] * 63
    cases.append((deep, have & (a > -18)))
    for prog, want in cases:
        assert np.array_equal(idx.filter_mask(prog, None), want)
    idx.close()


def test_filter_masks_survive_growth_of_the_shard():
    """The mask buffers come from a per-shard pool sized for the shard's capacity: after appends
    that grow the mirror, many more distinct filters than the pool holds must still evaluate
    correctly (undersized buffers are dropped, not recycled)."""
    d = 8
    idx = DeviceIndex(d)
    rows = o.fill_synthetic(40_000, d, 21)
    idx.load(rows[:1000])
    a = (np.arange(40_000) * 7919) % 1000
    idx.column_set(1, 0, np.full(1000, NM_V_INT, np.uint8), a[:1000].astype(np.uint64))
    for lim in (10, 500, 900):
        assert np.array_equal(idx.filter_mask([cmp_op(1, NM_C_LT, lim)], None), a[:1000] < lim)
    idx.append(rows[1000:])                                   # 40x: the mirror and every column grow
    idx.column_set(1, 1000, np.full(39_000, NM_V_INT, np.uint8), a[1000:].astype(np.uint64))
    built0 = idx.stats().filter_masks_built
    for j in range(30):                                       # > 12 pool entries, > 8 cached masks
        lim = 17 + 31 * j
        assert np.array_equal(idx.filter_mask([cmp_op(1, NM_C_LT, lim)], None), a < lim), lim
    assert idx.stats().filter_masks_built - built0 == 30
    (g,) = idx.search_filtered(rows[5], 3, "cosine", [cmp_op(1, NM_C_EQ, int(a[5]))])
    assert g[0][0] == 5
    idx.close()


@pytest.mark.parametrize("metric", ["cosine", "euclidean", "dot"])
def test_search_filtered_equals_the_oracle_on_the_subset(metric):
    n, d, k = 90_000, 48, 12
    idx = DeviceIndex(d)
    idx.fill_synthetic(n, 0x5EED0001)
    rows = o.fill_synthetic(n, d, 0x5EED0001)
    bucket = [int(i * 2654435761 % 97) for i in range(n)]
    idx.column_set(1, 0, *encode_column(bucket, {}))
    qs = o.fill_synthetic(3, d, 0x5EED1001)
    for lo, hi in ((10, 12), (0, 96), (96, 96), (50, 40)):       # 2 %, all, 1 %, nothing
        prog = [cmp_op(1, NM_C_GE, lo), cmp_op(1, NM_C_LE, hi), NmFilterOp(kind=NM_F_AND)]
        sub = np.nonzero([(lo <= b <= hi) for b in bucket])[0]
        s0 = idx.stats()
        res = idx.search_filtered(qs, k, metric, prog)
        for i in range(3):
            if sub.size == 0:
                assert res[i][0].size == 0
                continue
            er, es = o.search(rows[sub], qs[i], k, metric, threads=8)
            assert np.array_equal(res[i][0], sub[er.astype(np.int64)].astype(np.uint64)), (metric, lo, hi, i)
            assert np.array_equal(res[i][1].view(np.uint32), es.view(np.uint32))
        # the mask was built once and is reused until the next mutation
        res2 = idx.search_filtered(qs[:1], k, metric, prog)
        s1 = idx.stats()
        assert s1.filter_masks_built - s0.filter_masks_built == 1
        assert s1.filter_mask_hits - s0.filter_mask_hits == 1
        assert np.array_equal(res2[0][0], res[0][0])
    # k > 1024 goes through chained passes with the mask
    prog = [cmp_op(1, NM_C_LT, 30)]
    sub = np.nonzero([b < 30 for b in bucket])[0]
    ((r, s),) = idx.search_filtered(qs[0], 1500, metric, prog)
    er, es = o.search(rows[sub], qs[0], 1500, metric, threads=8)
    assert np.array_equal(r, sub[er.astype(np.int64)].astype(np.uint64))
    assert np.array_equal(s.view(np.uint32), es.view(np.uint32))
    idx.close()


def test_columns_follow_mutations():
    d = 8
    rows = o.fill_synthetic(300, d, 9)
    idx = DeviceIndex(d)
    idx.load(rows[:200])
    vals = list(range(200))
    idx.column_set(4, 0, *encode_column(vals, {}))
    prog_ge = lambda x: [cmp_op(4, NM_C_GE, x)]                  # noqa: E731
    assert idx.filter_mask(prog_ge(150)).sum() == 50
    # appended rows read as missing until they are set
    idx.append(rows[200:260])
    m = idx.filter_mask([NmFilterOp(kind=NM_F_EXISTS, column=4)])
    assert m[:200].all() and not m[200:].any()
    idx.column_set(4, 200, *encode_column(list(range(1000, 1060)), {}))
    assert idx.filter_mask(prog_ge(150)).sum() == 110
    # swap_remove moves the last row's entry into the hole
    moved = idx.swap_remove(10)
    assert moved == 259
    m = idx.filter_mask(prog_ge(1059))
    assert m.sum() == 1 and m[10] and idx.rows == 259
    # ... and a search through the filter ranks exactly that row
    ((r, s),) = idx.search_filtered(rows[0], 5, "dot", prog_ge(1059))
    assert list(r) == [10]
    # removing the last row itself; growth afterwards must not resurrect its entry
    idx.swap_remove(258)
    idx.append(rows[260:262])
    m = idx.filter_mask([NmFilterOp(kind=NM_F_EXISTS, column=4)])
    assert m[:258].all() and not m[258:].any()
    # load replaces the rows: columns are gone
    idx.load(rows[:50])
    assert not idx.filter_mask([NmFilterOp(kind=NM_F_EXISTS, column=4)]).any()
    with pytest.raises(NmError):
        idx.column_set(4, 40, *encode_column(list(range(20)), {}))   # rows out of range
    idx.close()


WHERE_CASES = [
    "bucket = 7", "bucket != 7", "bucket < 3 OR bucket >= 47", "price > 10.5 AND price <= 30",
    "price = 20", "bucket >= 12.5", "name >= 'n0100' AND name < 'n0200'", "name = 'n0042'",
    "STARTS_WITH(name, 'n01')", "CONTAINS(name, '99')", "EXISTS(opt)", "opt = true", "opt != true",
    "bucket IN (1, 2, 3, 'x')", "name IN ('n0001', 'n0002', 7)", "nothing = 1", "nothing != 1",
    "(bucket = 1 OR bucket = 2) AND (price < 5 OR CONTAINS(name, '7')) AND EXISTS(name)",
    "mixed = 3", "mixed = 'three'", "mixed > 2.5", "TRUE",
]


def _meta(i):
    m = {"bucket": i % 50, "price": (i % 97) * 0.5, "name": f"n{i:04d}"}
    if i % 3 == 0:
        m["opt"] = (i % 2 == 0)
    if i % 5 == 0:
        m["opt"] = None
    m["mixed"] = [3, "three", 3.5, True][i % 4]
    return m


def _py_eval(meta, where):
    """Independent evaluation of the WHERE cases above (reference rules)."""
    g = lambda f: meta.get(f, MISSING)                            # noqa: E731
    c = ref_cmp_field
    w = where
    if w == "bucket = 7": return c(g("bucket"), 7, NM_C_EQ)
    if w == "bucket != 7": return c(g("bucket"), 7, NM_C_NE)
    if w == "bucket < 3 OR bucket >= 47": return c(g("bucket"), 3, NM_C_LT) or c(g("bucket"), 47, NM_C_GE)
    if w == "price > 10.5 AND price <= 30": return c(g("price"), 10.5, NM_C_GT) and c(g("price"), 30, NM_C_LE)
    if w == "price = 20": return c(g("price"), 20, NM_C_EQ)
    if w == "bucket >= 12.5": return c(g("bucket"), 12.5, NM_C_GE)
    if w == "name >= 'n0100' AND name < 'n0200'": return c(g("name"), "n0100", NM_C_GE) and c(g("name"), "n0200", NM_C_LT)
    if w == "name = 'n0042'": return c(g("name"), "n0042", NM_C_EQ)
    if w == "STARTS_WITH(name, 'n01')": return isinstance(g("name"), str) and g("name").startswith("n01")
    if w == "CONTAINS(name, '99')": return isinstance(g("name"), str) and "99" in g("name")
    if w == "EXISTS(opt)": return g("opt") is not MISSING
    if w == "opt = true": return c(g("opt"), True, NM_C_EQ)
    if w == "opt != true": return c(g("opt"), True, NM_C_NE)
    if w == "bucket IN (1, 2, 3, 'x')": return any(c(g("bucket"), v, NM_C_EQ) for v in (1, 2, 3, "x"))
    if w == "name IN ('n0001', 'n0002', 7)": return any(c(g("name"), v, NM_C_EQ) for v in ("n0001", "n0002", 7))
    if w in ("nothing = 1", "nothing != 1"): return False
    if w.startswith("(bucket = 1"):
        return ((c(g("bucket"), 1, NM_C_EQ) or c(g("bucket"), 2, NM_C_EQ)) and
                (c(g("price"), 5, NM_C_LT) or (isinstance(g("name"), str) and "7" in g("name"))) and
                g("name") is not MISSING)
    if w == "mixed = 3": return c(g("mixed"), 3, NM_C_EQ)
    if w == "mixed = 'three'": return c(g("mixed"), "three", NM_C_EQ)
    if w == "mixed > 2.5": return c(g("mixed"), 2.5, NM_C_GT)
    if w == "TRUE": return True
    raise AssertionError(w)


def test_engine_pre_filter_runs_on_the_device_and_matches_the_reference_rules():
    n, d, k = 3000, 24, 15
    e = eng.VectorEngine()
    rows = o.fill_synthetic(n + 10, d, 41)
    metas = [_meta(i) for i in range(n)]
    for i in range(n):
        e.store_embedding_with_metadata(f"k{i:04d}", rows[i], metas[i])
    q = o.fill_synthetic(1, d, 42)[0]

    def check(keys, vecs, ms):
        for w in WHERE_CASES:
            sub = [j for j in range(len(keys)) if _py_eval(ms[j], w)]
            got = e.search_similar_filtered(q, k, w, eng.PRE_FILTER)
            if not sub:
                assert got == [], w
                continue
            er, es = o.search(np.stack([vecs[j] for j in sub]), q, k, "cosine", threads=4)
            assert [x.key for x in got] == [keys[sub[int(j)]] for j in er], w
            assert [np.float32(x.score).view(np.uint32) for x in got] == list(es.view(np.uint32)), w

    keys = [f"k{i:04d}" for i in range(n)]
    vecs = [rows[i] for i in range(n)]
    check(keys, vecs, metas)
    # metadata replaced, rows deleted (swap-remove) and new rows stored: the columns follow
    metas[7] = {"bucket": 7, "price": 20.0, "name": "n9999"}
    e.store_embedding_with_metadata("k0007", rows[7], metas[7])
    metas[8] = {}
    e.store_embedding("k0008", rows[8])                       # plain store drops the metadata
    for victim in (5, 1234, -1):
        victim = victim % len(keys)                             # -1: the last row itself
        e.delete_embedding(keys[victim])
        last = len(keys) - 1
        keys[victim], vecs[victim], metas[victim] = keys[last], vecs[last], metas[last]
        keys.pop(); vecs.pop(); metas.pop()
    for j in range(4):
        keys.append(f"new{j}")
        vecs.append(rows[n + j])
        metas.append({"bucket": 7, "name": f"n{j:04d}", "fresh": 1.5})
        e.store_embedding_with_metadata(keys[-1], vecs[-1], metas[-1])
    check(keys, vecs, metas)
    got = e.search_similar_filtered(q, 3, "fresh > 1", eng.PRE_FILTER)
    assert len(got) == 3 and all(x.key.startswith("new") for x in got)
    e.close()
