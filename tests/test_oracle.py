"""CPU tests of the oracle (oracle/nm_oracle.c): pinned against every KAT the reference's own
tests hold for the path, against an independent numpy restatement, and against committed
golden vectors."""
import json
from pathlib import Path

import numpy as np
import pytest

import np_ref
import oracle_ffi as o

GOLD = Path(__file__).parent / "golden"
KATS = json.loads((GOLD / "reference_kats.json").read_text())


@pytest.mark.parametrize("kat", KATS["scalar"], ids=lambda k: k["src"])
def test_reference_scalar_kats(kat):
    fn = kat["fn"]
    if fn == "dot_product":
        got = o.dot_product(kat["a"], kat["b"])
    elif fn == "magnitude":
        got = o.magnitude(kat["a"])
    elif fn == "compute_similarity":
        got = o.compute_similarity(kat["a"], kat["b"])
    else:
        got = o.euclidean_distance(kat["a"], kat["b"])
    assert not np.isnan(got)
    assert abs(float(got) - kat["expect"]) <= kat["tol"]


def _run_search_kat(kat, search_fn):
    """search_fn(rows[n,d], query, k, metric) -> (rows, scores); applies the host-wrapper rules
    (dimension bucketing, zero-query short-circuit) the way VectorEngine does."""
    q = np.asarray(kat["query"], np.float32)
    metric = kat.get("metric", "cosine")
    keys = [k for k, v in kat["store"].items() if len(v) == q.size]
    if (o.magnitude(q) == 0 and metric != "euclidean") or not keys:
        return []
    rows = np.asarray([kat["store"][k] for k in keys], np.float32)
    r, s = search_fn(rows, q, kat["k"], metric)
    return [(keys[int(i)], float(x)) for i, x in zip(r, s)]


def check_search_kat(kat, hits):
    assert len(hits) == kat["len"]
    if "first" in kat:
        assert hits[0][0] in kat["first"]
    if "last" in kat:
        assert hits[-1][0] == kat["last"]
    if "order" in kat:
        assert [h[0] for h in hits] == kat["order"]
    if "first_score" in kat:
        assert abs(hits[0][1] - kat["first_score"]) < kat["tol"]
    for key, want in kat.get("scores", {}).items():
        got = dict(hits)[key]
        assert abs(got - want) < kat["tol"], (key, got, want)


@pytest.mark.parametrize("kat", KATS["search"], ids=lambda k: k["src"])
def test_reference_search_kats(kat):
    check_search_kat(kat, _run_search_kat(kat, o.search))
    check_search_kat(kat, _run_search_kat(kat, lambda *a: o.search(*a, threads=3)))


def test_reference_merge_top_k_kat():
    kat = KATS["merge_top_k"][0]
    keys, rows, scores = [], [], []
    for shard in kat["shards"]:
        rows.append([len(keys) + i for i in range(len(shard))])
        scores.append([s for _, s in shard])
        keys += [k for k, _ in shard]
    r, s = o.merge_top_k(rows, scores, kat["k"])
    assert [keys[int(i)] for i in r] == kat["expect_keys"]


def create_test_vector(dim, seed):
    """vector_engine/src/lib.rs:4029-4038 (f32 sin is libm-dependent: ranks, not bits)."""
    i = np.arange(dim)
    x = (seed * 31 + i * 17).astype(np.float32)
    return (np.sin((x * np.float32(0.0001)).astype(np.float32)).astype(np.float32)
            * ((seed + i).astype(np.float32) * np.float32(0.001)).astype(np.float32)).astype(np.float32)


def test_store_10000_vectors_search_kat():
    # vector_engine/src/lib.rs:4256-4276 (BASELINE config 1): 10k x 128, cosine top-5
    rows = np.stack([create_test_vector(128, i) for i in range(10000)])
    q = create_test_vector(128, 5000)
    r, s = o.search(rows, q, 5, "cosine")
    assert len(r) == 5 and r[0] == 5000 and abs(s[0] - 1.0) < 1e-5
    r2, s2 = o.search(rows, q, 5, "cosine", threads=4)
    assert np.array_equal(r, r2) and np.array_equal(s.view(np.uint32), s2.view(np.uint32))


@pytest.mark.parametrize("dim,probe", [(768, 50), (1536, 75)])
def test_high_dimensional_kats(dim, probe):
    # vector_engine/src/lib.rs:4279-4312
    rows = np.stack([create_test_vector(dim, i) for i in range(100)])
    r, _ = o.search(rows, create_test_vector(dim, probe), 5, "cosine")
    assert r[0] == probe


def test_high_dimension_4096_kat():
    # vector_engine/src/lib.rs:6024-6037
    i = np.arange(4096, dtype=np.float32)
    v1 = np.sin(i * np.float32(0.001)).astype(np.float32)
    v2 = np.sin(i * np.float32(0.002)).astype(np.float32)
    r, _ = o.search(np.stack([v1, v2]), v1, 2, "cosine")
    assert list(r) == [0, 1]


def test_example_vector_search():
    # examples/vector_search.rs:26-131 (8 docs x 8 dims, three TOP-3 queries)
    ex = json.loads((GOLD / "example_vector_search.json").read_text())
    keys = list(ex["docs"])
    rows = np.asarray([ex["docs"][k] for k in keys], np.float32)
    for qname, spec in ex["queries"].items():
        r, s = o.search(rows, np.asarray(spec["vector"], np.float32), 3, "cosine")
        got = [keys[int(i)] for i in r]
        assert len(got) == 3
        assert set(got[:len(spec["top_set"])]) == set(spec["top_set"]), (qname, got)


# ---- independent restatement ---------------------------------------------------------------
@pytest.mark.parametrize("dim", [1, 3, 7, 8, 9, 15, 16, 31, 33, 100, 128, 257])
def test_oracle_matches_numpy_restatement_bits(dim):
    rng = np.random.default_rng(dim)
    for scale in (1.0, 1e-20, 1e18):
        a = (rng.standard_normal(dim) * scale).astype(np.float32)
        b = (rng.standard_normal(dim) * scale).astype(np.float32)
        assert o.dot_product(a, b).view(np.uint32) == np_ref.dot_product(a, b).view(np.uint32)
        assert o.magnitude(a).view(np.uint32) == np_ref.magnitude(a).view(np.uint32)
        assert o.euclidean_distance(a, b).view(np.uint32) == np_ref.euclidean_distance(a, b).view(np.uint32)
        for m in ("cosine", "dot", "euclidean"):
            assert o.compute_score(a, b, m).view(np.uint32) == np_ref.score(a, b, m).view(np.uint32)


@pytest.mark.parametrize("metric", ["cosine", "dot", "euclidean"])
@pytest.mark.parametrize("n,dim", [(500, 128), (300, 77), (64, 768)])
def test_oracle_rows_and_topk_match_numpy(metric, n, dim):
    rows = o.fill_synthetic(n, dim, 0x5EED0001)
    q = o.fill_synthetic(1, dim, 0x5EED1001)[0]
    sc = o.score_rows(rows, q, metric)
    ref = np_ref.score_rows_vectorised(rows, q, metric)
    assert np.array_equal(sc.view(np.uint32), ref.view(np.uint32))
    r, s = o.search(rows, q, 10, metric)
    er, es = np_ref.topk(ref, 10)
    assert np.array_equal(r, er) and np.array_equal(s.view(np.uint32), es.view(np.uint32))


def test_lane_order_is_observable():
    """The 8-lane tree differs from a plain left fold on ordinary data, so the bit-parity
    tests really pin the summation order."""
    rows = o.fill_synthetic(200, 768, 7)
    q = o.fill_synthetic(1, 768, 8)[0]
    seq = np.zeros(200, np.float32)
    for i in range(768):
        seq = (seq + (rows[:, i] * q[i]).astype(np.float32)).astype(np.float32)
    tree = o.score_rows(rows, q, "dot")
    assert np.count_nonzero(seq.view(np.uint32) != tree.view(np.uint32)) > 50


# ---- ordering rules ------------------------------------------------------------------------
def test_ties_negzero_nan_order():
    # rows 0,2 tie at +0.0/-0.0 (dot with +-0), row 1 NaN, row 3 negative, row 4 positive
    rows = np.array([[0.0, 0.0], [np.nan, 0.0], [-0.0, 0.0], [-1.0, 0.0], [1.0, 0.0]], np.float32)
    q = np.array([1.0, 0.0], np.float32)
    r, s = o.search(rows, q, 5, "dot")
    assert list(r) == [4, 0, 2, 3, 1]
    assert np.isnan(s[4])
    r2, _ = o.search(rows, q, 5, "dot", threads=2)
    assert list(r2) == list(r)


def test_duplicate_rows_tie_by_row():
    base = o.fill_synthetic(50, 16, 3)
    rows = np.concatenate([base, base, base])
    q = base[7]
    r, s = o.search(rows, q, 6, "cosine")
    assert list(r[:3]) == [7, 57, 107]
    assert s[0] == s[1] == s[2]


def test_k_larger_than_n_and_empty():
    rows = o.fill_synthetic(5, 8, 1)
    r, s = o.search(rows, rows[0], 100, "euclidean")
    assert len(r) == 5 and r[0] == 0 and s[0] == 1.0


def test_merge_equals_global_search():
    rows = o.fill_synthetic(3000, 64, 11)
    rows[100] = rows[2900]  # a cross-shard exact tie
    q = rows[2900]
    bounds = [0, 700, 1500, 3000]
    sr, ss = [], []
    for a, b in zip(bounds[:-1], bounds[1:]):
        r, s = o.search(rows[a:b], q, 10, "cosine")
        sr.append(r + a)
        ss.append(s)
    mr, ms = o.merge_top_k(sr, ss, 10)
    gr, gs = o.search(rows, q, 10, "cosine")
    assert np.array_equal(mr, gr) and np.array_equal(ms.view(np.uint32), gs.view(np.uint32))
    assert list(gr[:2]) == [100, 2900]


# ---- synthetic generator + golden drift check ----------------------------------------------
def test_synthetic_generator_properties():
    a = o.fill_synthetic(1000, 96, 0x5EED0001)
    assert a.min() >= -1.0 and a.max() < 1.0 and abs(float(a.mean())) < 0.01
    b = o.fill_synthetic(10, 96, 0x5EED0001, row_offset=990)
    assert np.array_equal(a[990:], b)
    c = o.fill_synthetic(1000, 96, 0x5EED1001)
    assert not np.array_equal(a, c)
    # exactly representable: value * 2^23 is an integer
    assert np.all(np.modf(a.astype(np.float64) * 2 ** 23)[0] == 0)


def test_oracle_golden_vectors():
    """Committed outputs of the oracle on seeded inputs (tests/golden/make_oracle_vectors.py):
    any drift in the oracle's arithmetic or ordering shows up as a diff here."""
    g = np.load(GOLD / "oracle_vectors.npz")
    for name in ("cosine", "euclidean", "dot"):
        n, dim, k = (int(x) for x in g[f"{name}_shape"])
        rows = o.fill_synthetic(n, dim, int(g["seed_rows"]))
        q = o.fill_synthetic(1, dim, int(g["seed_query"]))[0]
        r, s = o.search(rows, q, k, name)
        assert np.array_equal(r, g[f"{name}_rows"])
        assert np.array_equal(s.view(np.uint32), g[f"{name}_score_bits"])


def test_numpy_generator_twin_matches_oracle_generator():
    from neumann_b200.synth import synth_rows
    a = synth_rows(64, 77, 0x5EED1001, row_offset=5)
    b = o.fill_synthetic(64, 77, 0x5EED1001, row_offset=5)
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
