"""the int8 pre-filter must return exactly what the f32 scan returns — same
row ids in the same order, bit-identical scores — on friendly and on adversarial data, and
must fall back to the f32 scan when its candidate list overflows or the query is not finite."""
import numpy as np
import pytest

import oracle_ffi as o
from neumann_b200 import DeviceIndex

pytestmark = pytest.mark.gpu


def assert_same(got, exp, ctx=""):
    assert np.array_equal(got[0], exp[0]), f"{ctx}: rows {got[0][:8]} != {exp[0][:8]}"
    nan = np.isnan(exp[1])
    assert np.array_equal(np.isnan(got[1]), nan), ctx
    assert np.array_equal(got[1].view(np.uint32)[~nan], exp[1].view(np.uint32)[~nan]), ctx


def check(idx, rows, queries, k, metric, ctx=""):
    s0 = idx.stats()
    for i, q in enumerate(queries):
        (got,) = idx.search(q, k, metric)
        assert_same(got, o.search(rows, q, k, metric, threads=8), f"{ctx} {metric} q{i} k={k}")
    s1 = idx.stats()
    return s1.prefilter_queries - s0.prefilter_queries, s1.prefilter_fallbacks - s0.prefilter_fallbacks, \
        s1.prefilter_kept - s0.prefilter_kept


@pytest.mark.parametrize("metric", ["cosine", "dot", "euclidean"])
@pytest.mark.parametrize("n,dim,k", [(50_000, 128, 10), (20_000, 768, 10), (30_000, 100, 100),
                                     (9_000, 1536, 1), (70_000, 64, 1000), (3_000, 13, 5),
                                     (400, 8, 7), (5_000, 4096, 3)])
def test_prefilter_equals_exact_scan_synthetic(metric, n, dim, k):
    rows = o.fill_synthetic(n, dim, 0x5EED0001)
    idx = DeviceIndex(dim)
    idx.load(rows)
    idx.set_prefilter(1)
    qs = o.fill_synthetic(4, dim, 0x5EED1001)
    qs[1] = rows[n // 3]                       # query equal to a stored row
    used, fell, kept = check(idx, rows, qs, k, metric, f"n={n} d={dim}")
    assert used == 4
    idx.close()


@pytest.mark.parametrize("metric", ["cosine", "dot", "euclidean"])
def test_prefilter_adversarial_data(metric):
    """Near-duplicates inside the quantisation error, exact duplicates, outlier elements (coarse
    int8 scale), tiny and huge magnitudes, zero rows, negative copies."""
    n, d, k = 40_000, 96, 20
    rng = np.random.default_rng(5)
    base = o.fill_synthetic(n, d, 21)
    rows = base.copy()
    q = o.fill_synthetic(1, d, 22)[0]
    rows[1000:1400] = q + rng.normal(0, 1e-4, (400, d)).astype(np.float32)   # indistinguishable in int8
    rows[2000:2100] = rows[1000]                                              # exact ties
    rows[3000:3200, 0] = 1000.0                                               # outlier -> coarse scale
    rows[4000:4100] *= np.float32(1e-20)
    rows[4100:4200] *= np.float32(1e15)
    rows[5000:5050] = 0.0
    rows[6000:6100] = -rows[1000:1100]
    rows[7000:7050] *= np.float32(1e-38)          # denormal elements: scale below FLT_MIN
    rows[7050:7060] = np.float32(1e-45)
    idx = DeviceIndex(d)
    idx.load(rows)
    idx.set_prefilter(1)
    for kk in (1, k, 500):
        check(idx, rows, [q, rows[3100], rows[4050], rows[4150], np.abs(q), rows[7010],
                          q * np.float32(1e30)], kk, metric, "adversarial")
    idx.close()


def test_prefilter_falls_back_on_overflow_and_nonfinite():
    n, d = 1_200_000, 8
    rows = np.ones((n, d), np.float32)           # every row ties: all 1.2M are candidates
    idx = DeviceIndex(d)
    idx.load(rows)
    idx.set_prefilter(1)
    q = np.ones(d, np.float32)
    used, fell, _ = check(idx, rows, [q], 10, "cosine", "overflow")
    assert used == 1 and fell == 1
    qn = q.copy(); qn[3] = np.inf
    (got,) = idx.search(qn, 5, "dot")            # non-finite query -> exact scan decides
    assert_same(got, o.search(rows, qn, 5, "dot", threads=8), "inf query")
    assert idx.stats().prefilter_fallbacks >= 2
    idx.close()


def test_prefilter_rows_with_nonfinite_values_stay_exact():
    n, d = 20_000, 32
    rows = o.fill_synthetic(n, d, 9)
    rows[7, 3] = np.nan
    rows[8, 0] = np.inf
    rows[9] = 3e38                                # overflows in the reference arithmetic
    idx = DeviceIndex(d)
    idx.load(rows)
    idx.set_prefilter(1)
    for metric in ("cosine", "dot", "euclidean"):
        check(idx, rows, [o.fill_synthetic(1, d, 10)[0]], 1000, metric, "nonfinite rows")
    idx.close()


def test_prefilter_follows_mirror_mutations():
    d = 48
    host = o.fill_synthetic(6000, d, 31)
    idx = DeviceIndex(d)
    idx.set_prefilter(1)                          # enabled on an empty index
    idx.load(host[:4000])
    idx.append(host[4000:])
    q = o.fill_synthetic(1, d, 32)[0]
    check(idx, host, [q], 10, "cosine", "after append")
    host = host.copy()
    host[123] = q
    idx.update(123, q)
    (got,) = idx.search(q, 3, "cosine")
    assert got[0][0] == 123
    check(idx, host, [q], 10, "dot", "after update")
    moved = idx.swap_remove(123)
    host[123] = host[moved]
    host = host[:-1]
    check(idx, host, [q], 10, "cosine", "after swap_remove")
    idx.set_prefilter(0)
    s = idx.stats().prefilter_queries
    check(idx, host, [q], 10, "cosine", "prefilter off")
    assert idx.stats().prefilter_queries == s
    idx.close()


def test_prefilter_full_size_10m_x_768_matches_exact_scan():
    n, d, k = 10_000_000, 768, 10
    idx = DeviceIndex(d)
    idx.fill_synthetic(n, 0x5EED0001)
    qs = o.fill_synthetic(6, d, 0x5EED1001)
    exact = [idx.search(q, k, "cosine")[0] for q in qs]
    idx.set_prefilter(1)
    for i, q in enumerate(qs):
        (got,) = idx.search(q, k, "cosine")
        assert_same(got, exact[i], f"10M q{i}")
    st = idx.stats()
    assert st.prefilter_queries == 6 and st.prefilter_fallbacks == 0
    print("kept per query:", st.prefilter_kept / 6)
    idx.close()
