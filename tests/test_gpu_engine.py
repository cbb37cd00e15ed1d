"""The reference's own behavioural tests for the path, replayed through the C++ host mirror
(VectorEngine / QueryRouter) with the scan on the GPU."""
import json
import threading
from pathlib import Path

import numpy as np
import pytest

import oracle_ffi as o
from neumann_b200 import engine as eng
from test_oracle import check_search_kat, create_test_vector

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).parent / "golden"
KATS = json.loads((GOLD / "reference_kats.json").read_text())
METRIC_ID = {"cosine": eng.COSINE, "euclidean": eng.EUCLIDEAN, "dot": eng.DOT_PRODUCT}


@pytest.mark.parametrize("kat", KATS["search"], ids=lambda k: k["src"])
def test_reference_search_kats_on_gpu(kat):
    e = eng.VectorEngine()
    for k, v in kat["store"].items():
        e.store_embedding(k, v)
    if kat["api"] == "search_similar":
        res = e.search_similar(kat["query"], kat["k"])
    else:
        res = e.search_similar_with_metric(kat["query"], kat["k"], METRIC_ID[kat["metric"]])
    check_search_kat(kat, [(r.key, r.score) for r in res])


def test_store_10000_vectors_search():
    # vector_engine/src/lib.rs:4256-4276
    e = eng.VectorEngine()
    for i in range(10000):
        e.store_embedding(f"v{i}", create_test_vector(128, i))
    assert e.count() == 10000
    res = e.search_similar(create_test_vector(128, 5000), 5)
    assert len(res) == 5 and res[0].key == "v5000" and abs(res[0].score - 1.0) < 1e-5
    assert e.mirror_rows(128) == (10000, 10000)


@pytest.mark.parametrize("dim,probe", [(768, 50), (1536, 75)])
def test_high_dimensional(dim, probe):
    e = eng.VectorEngine()
    for i in range(100):
        e.store_embedding(f"v{i}", create_test_vector(dim, i))
    assert e.search_similar(create_test_vector(dim, probe), 5)[0].key == f"v{probe}"


def test_example_vector_search():
    ex = json.loads((GOLD / "example_vector_search.json").read_text())
    e = eng.VectorEngine()
    for k, v in ex["docs"].items():
        e.store_embedding(k, v)
    keys = list(ex["docs"])
    rows = np.asarray([ex["docs"][k] for k in keys], np.float32)
    for spec in ex["queries"].values():
        res = e.search_similar(spec["vector"], 3)
        assert {r.key for r in res[:len(spec["top_set"])]} == set(spec["top_set"])
        er, es = o.search(rows, np.asarray(spec["vector"], np.float32), 3, "cosine")
        assert [r.key for r in res] == [keys[int(i)] for i in er]
        assert [np.float32(r.score).view(np.uint32) for r in res] == list(es.view(np.uint32))


def test_search_with_metric_parallel_path():
    # vector_engine/src/lib.rs:6583-6603
    e = eng.VectorEngine(parallel_threshold=5)
    for i in range(10):
        e.store_embedding(f"vec_{i}", [float(i), 0.0, 0.0])
    res = e.search_similar_with_metric([5.0, 0.0, 0.0], 3, eng.EUCLIDEAN)
    assert len(res) == 3 and res[0].key == "vec_5" and res[0].score == 1.0
    assert {res[1].key, res[2].key} == {"vec_4", "vec_6"} and res[1].score == 0.5


def test_mirror_is_maintained_incrementally():
    e = eng.VectorEngine()
    rows = o.fill_synthetic(500, 32, 3)
    for i in range(300):
        e.store_embedding(f"k{i}", rows[i])
    q = rows[10]
    assert e.search_similar(q, 1)[0].key == "k10"
    assert e.mirror_rows(32) == (300, 300)
    for i in range(300, 500):
        e.store_embedding(f"k{i}", rows[i])          # appended lazily at the next search
    assert e.mirror_rows(32) == (500, 300)
    assert e.search_similar(rows[400], 1)[0].key == "k400"
    assert e.mirror_rows(32) == (500, 500)
    e.store_embedding("k7", rows[400])               # overwrite in place -> exact tie with k400
    res = e.search_similar(rows[400], 2)
    assert {res[0].key, res[1].key} == {"k7", "k400"} and res[0].score == res[1].score
    e.delete_embedding("k400")
    res = e.search_similar(rows[400], 2)
    assert res[0].key == "k7" and res[1].key != "k400"
    assert e.mirror_rows(32) == (499, 499)
    e.store_embedding("k8", [1.0, 2.0])              # key moves to another dimension bucket
    assert e.mirror_rows(32)[0] == 498
    assert e.search_similar([1.0, 2.0], 5)[0].key == "k8"


def test_engine_results_match_oracle_order_and_bits():
    e = eng.VectorEngine()
    rows = o.fill_synthetic(3000, 48, 12)
    for i in range(3000):
        e.store_embedding(f"k{i}", rows[i])
    q = o.fill_synthetic(1, 48, 13)[0]
    for name, mid in METRIC_ID.items():
        res = e.search_similar_with_metric(q, 25, mid)
        er, es = o.search(rows, q, 25, name)
        assert [r.key for r in res] == [f"k{int(i)}" for i in er]
        assert [np.float32(r.score).view(np.uint32) for r in res] == list(es.view(np.uint32))


def test_search_similar_batch_matches_single_calls_and_oracle():
    """search_similar_batch[i] == search_similar_with_metric(queries[i]): by default batches go
    through the tensor-core pre-filter (auto mode) and single calls through the f32 scan; with
    device_prefilter=True single calls use the dp4a pre-filter too."""
    n, d, k = 70_000, 40, 9
    rows = o.fill_synthetic(n, d, 12)
    qs = o.fill_synthetic(7, d, 13)
    qs[3] = 0.0                                   # zero query: empty for cosine / dot
    for pre in (False, True):
        e = eng.VectorEngine(device_prefilter=pre)
        for i in range(n):
            e.store_embedding(f"k{i}", rows[i])
        for name, mid in METRIC_ID.items():
            res = e.search_similar_batch(qs, k, mid)
            assert len(res) == 7
            for i in range(7):
                single = e.search_similar_with_metric(qs[i], k, mid)
                assert [(r.key, np.float32(r.score).view(np.uint32)) for r in res[i]] == \
                       [(r.key, np.float32(r.score).view(np.uint32)) for r in single], (pre, name, i)
                if i == 3 and name != "euclidean":
                    assert res[i] == []
                    continue
                er, es = o.search(rows, qs[i], k, name, threads=8)
                assert [r.key for r in res[i]] == [f"k{int(j)}" for j in er], (pre, name, i)
                assert [np.float32(r.score).view(np.uint32) for r in res[i]] == list(es.view(np.uint32))
        e.close()


def test_collections_on_gpu():
    e = eng.VectorEngine()
    e.create_collection("euc", dimension=2, metric=eng.EUCLIDEAN)
    for k, v in {"origin": [0.0, 0.0], "unit": [1.0, 0.0], "far": [10.0, 0.0]}.items():
        e.store_in_collection("euc", k, v)
    e.store_embedding("outside", [0.0, 0.0])
    res = e.search_in_collection("euc", [0.0, 0.0], 3)   # zero query is fine for Euclidean
    assert [r.key for r in res] == ["origin", "unit", "far"]
    assert res[0].score == 1.0 and res[1].score == 0.5
    e.store_in_collection("cos", "a", [1.0, 0.0])
    assert e.search_in_collection("cos", [0.0, 0.0], 3) == []  # cosine zero query -> empty
    assert e.search_in_collection("cos", [2.0, 0.0], 3)[0].key == "a"
    e.delete_from_collection("euc", "origin")
    assert e.search_in_collection("euc", [0.0, 0.0], 3)[0].key == "unit"


def test_router_similar_operator():
    # query_router/src/lib.rs:7669-7685, 8051-8064, 8467-8474
    e = eng.VectorEngine()
    e.execute("EMBED doc1 [1.0, 0.0, 0.0]")
    e.execute("EMBED doc2 [0.0, 1.0, 0.0]")
    e.execute("EMBED doc3 [0.9, 0.1, 0.0]")
    res = e.execute("SIMILAR doc1 TOP 2")
    assert len(res) == 2 and res[0].key == "doc1" and res[1].key == "doc3"
    assert len(e.execute("SIMILAR doc1")) == 3           # default k = 10
    res = e.execute("SIMILAR [0.0, 1.0, 0.0] TOP 1")
    assert len(res) == 1 and res[0].key == "doc2"
    assert len(e.execute('SIMILAR "doc2" TOP 1')) == 1
    # integration_tests/tests/distance_metrics.rs:39-140 (AST path, LIMIT + metric keyword)
    e.execute("EMBED vec:1 1.0, 0.0, 0.0, 0.0")
    e.execute("EMBED vec:2 0.9, 0.1, 0.0, 0.0")
    e.execute("EMBED vec:3 0.0, 1.0, 0.0, 0.0")
    res = e.execute_parsed("SIMILAR 'vec:1' LIMIT 3")
    assert [r.key for r in res] == ["vec:1", "vec:2", "vec:3"]
    res = e.execute_parsed("SIMILAR 'vec:1' LIMIT 3 EUCLIDEAN")
    assert res[0].key == "vec:1" and res[0].score == 1.0
    res = e.execute_parsed("SIMILAR [1.0, 0.0, 0.0, 0.0] LIMIT 2 DOT_PRODUCT")
    assert [r.key for r in res] == ["vec:1", "vec:2"] and res[0].score == 1.0
    e.execute_parsed("EMBED STORE 'c1' [0.0, 3.0] INTO things")
    e.execute_parsed("EMBED STORE 'c2' [4.0, 0.0] INTO things")
    res = e.execute_parsed("SIMILAR [1.0, 0.0] LIMIT 5 INTO things")
    assert [r.key for r in res] == ["c2", "c1"]


def test_search_timeout_is_reported():
    e = eng.VectorEngine(search_timeout_ms=0)
    e.store_embedding("a", [1.0, 0.0])
    with pytest.raises(eng.VectorError) as ei:
        e.search_similar([1.0, 0.0], 1)
    assert ei.value.kind == "SearchTimeout" and "search_similar" in str(ei.value)


def test_concurrent_search_and_store():
    # vector_engine/src/lib.rs:5615-5711: searches from many threads while stores happen
    e = eng.VectorEngine()
    rows = o.fill_synthetic(4000, 64, 31)
    for i in range(2000):
        e.store_embedding(f"k{i}", rows[i])
    errors = []

    def searcher(tid):
        try:
            for j in range(40):
                probe = (tid * 97 + j * 13) % 2000
                res = e.search_similar(rows[probe], 3)
                if res[0].key != f"k{probe}":
                    errors.append((tid, j, res[0].key))
        except Exception as ex:  # noqa: BLE001
            errors.append(repr(ex))

    def writer():
        try:
            for i in range(2000, 4000):
                e.store_embedding(f"k{i}", rows[i])
        except Exception as ex:  # noqa: BLE001
            errors.append(repr(ex))

    ts = [threading.Thread(target=searcher, args=(t,)) for t in range(6)] + \
         [threading.Thread(target=writer)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errors, errors[:3]
    assert e.count() == 4000
    assert e.search_similar(rows[3999], 1)[0].key == "k3999"


# ---- filtered search (vector_engine/src/lib.rs:6968-7720) and PointsService::query ----------
from test_filter_cpu import setup_filtered_search_engine  # noqa: E402


@pytest.mark.parametrize("strategy", [eng.AUTO, eng.PRE_FILTER, eng.POST_FILTER])
def test_search_filtered_reference_kats(strategy):
    e = setup_filtered_search_engine()
    q = [1.0, 1.0, 1.0]
    r = e.search_similar_filtered(q, 10, "category = 'electronics'", strategy)
    assert [x.key for x in r] == ["item0"]                       # :7004-7017
    r = e.search_similar_filtered([1.0, 0.0, 0.0], 10, "price = 50", strategy)
    assert [x.key for x in r] == ["item1"]                       # :7020-7030
    assert len(e.search_similar_filtered(q, 10, "price > 30", strategy)) == 2
    r = e.search_similar_filtered(q, 10, "price > 30 AND price < 80", strategy)
    assert [x.key for x in r] == ["item1"]                       # :7083-7095
    assert len(e.search_similar_filtered(q, 10, "TRUE", strategy)) == 3   # :7117
    assert e.search_similar_filtered(q, 10, "price > 1000", strategy) == []  # :7322
    assert len(e.search_similar_filtered(q, 1, "price > 30", strategy)) == 1  # respects top_k :7701


def test_pre_filter_matches_oracle_on_subset_and_post_filter_oversamples():
    e = eng.VectorEngine()
    rows = o.fill_synthetic(5000, 32, 77)
    for i in range(5000):
        e.store_embedding_with_metadata(f"k{i}", rows[i], {"bucket": i % 50, "name": f"n{i:05d}"})
    q = o.fill_synthetic(1, 32, 78)[0]
    sub = np.nonzero(np.arange(5000) % 50 == 7)[0]
    er, es = o.search(rows[sub], q, 10, "cosine")
    r = e.search_similar_filtered(q, 10, "bucket = 7", eng.PRE_FILTER)
    assert [x.key for x in r] == [f"k{int(sub[int(i)])}" for i in er]
    assert [np.float32(x.score).view(np.uint32) for x in r] == list(es.view(np.uint32))
    assert [x.key for x in e.search_similar_filtered(q, 10, "bucket = 7")] == [x.key for x in r]  # auto -> pre (2 % selective)
    # post-filter: filter(top (k*oversample)) — approximate by design (lib.rs:3560-3578)
    gr, _ = o.search(rows, q, 30, "cosine")
    want = [f"k{int(i)}" for i in gr if int(i) % 2 == 0][:10]
    r = e.search_similar_filtered(q, 10, "bucket IN (0,2,4,6,8,10,12,14,16,18,20,22,24,26,28,30,32,34,36,38,40,42,44,46,48)",
                                  eng.POST_FILTER, oversample_factor=3)
    assert [x.key for x in r] == want
    r = e.execute_parsed("SIMILAR 'k7' LIMIT 3 WHERE bucket = 7 AND name >= 'n00007'")
    assert r[0].key == "k7" and all(int(x.key[1:]) % 50 == 7 for x in r)


def test_search_filtered_in_collection_and_query_points():
    e = eng.VectorEngine()
    e.create_collection("docs", dimension=3, metric=eng.COSINE)
    for i in range(20):
        e.store_in_collection_with_metadata("docs", f"d{i}", [1.0, float(i), 0.5], {"even": i % 2 == 0})
    r = e.search_filtered_in_collection("docs", [1.0, 0.0, 0.5], 3, "even = true", eng.PRE_FILTER)
    assert [x.key for x in r] == ["d0", "d2", "d4"]
    r = e.execute_parsed("SIMILAR [1.0, 0.0, 0.5] LIMIT 2 INTO docs WHERE even = false")
    assert [x.key for x in r] == ["d1", "d3"]
    # points.rs:449-485: search(limit+offset) -> skip(offset) -> take(limit) -> threshold
    full = e.search_in_collection("docs", [1.0, 0.0, 0.5], 20)
    page = e.query_points("docs", [1.0, 0.0, 0.5], limit=5, offset=3)
    assert [x.key for x in page] == [x.key for x in full[3:8]]
    thr = full[5].score
    page = e.query_points("docs", [1.0, 0.0, 0.5], limit=10, offset=0, score_threshold=thr)
    assert [x.key for x in page] == [x.key for x in full[:10] if x.score >= thr]


def test_search_entities_on_gpu():
    # vector_engine/src/lib.rs:3155-3219 (the call tensor_unified makes, :913-914)
    e = eng.VectorEngine()
    rows = o.fill_synthetic(800, 24, 5)
    for i in range(800):
        e.set_entity_embedding(f"user:{i}", rows[i])
    e.store_embedding("not_an_entity", rows[3])          # different namespace: never returned
    q = rows[77]
    res = e.search_entities(q, 5)
    er, es = o.search(rows, q, 5, "cosine")
    assert [r.key for r in res] == [f"user:{int(i)}" for i in er]
    assert [np.float32(r.score).view(np.uint32) for r in res] == list(es.view(np.uint32))
    e.remove_entity_embedding("user:77")
    assert e.search_entities(q, 1)[0].key != "user:77"


def test_delete_collection_is_safe_against_concurrent_searches_and_stores():
    """Advisor finding (round 1): delete_collection used to free a Space (rows + device mirror)
    that concurrent search_in_collection / store_in_collection calls were still using.  Spaces are
    shared_ptr-held now: hammer one collection name with searches, stores and delete/recreate
    cycles from several threads; nothing may crash and every result must be well formed."""
    import threading
    e = eng.VectorEngine()
    d = 32
    rows = o.fill_synthetic(600, d, 91)
    stop = threading.Event()
    errors = []

    def fill():
        for i in range(300):
            e.store_in_collection("hot", f"k{i}", rows[i])

    e.create_collection("hot", dimension=d)
    fill()

    def searcher():
        try:
            while not stop.is_set():
                res = e.search_in_collection("hot", rows[7], 5)
                assert len(res) <= 5 and all(r.key.startswith("k") for r in res)
                scores = [r.score for r in res]
                assert scores == sorted(scores, reverse=True)
        except Exception as ex:  # noqa: BLE001
            errors.append(repr(ex))

    def storer():
        try:
            i = 300
            while not stop.is_set():
                e.store_in_collection("hot", f"k{i % 600}", rows[i % 600])
                i += 1
        except Exception as ex:  # noqa: BLE001
            errors.append(repr(ex))

    ts = [threading.Thread(target=searcher) for _ in range(3)] + [threading.Thread(target=storer)]
    for t in ts:
        t.start()
    for _ in range(25):
        try:
            e.delete_collection("hot")
        except eng.VectorError:
            pass                      # the storer may have re-created the rows without a config
        try:
            e.create_collection("hot", dimension=d)
        except eng.VectorError:
            pass
        fill()
    stop.set()
    for t in ts:
        t.join()
    assert not errors, errors[:2]
    res = e.search_in_collection("hot", rows[7], 1)
    assert res and res[0].key == "k7"
    e.close()


def test_paginated_searches_and_metadata_updates_on_gpu():
    """search_similar_paginated / search_entities_paginated (vector_engine/src/lib.rs:2988-3058,
    KATs :5367-5397, :9232-9250) slice the ranked hits of the device scan; update_metadata /
    remove_metadata_field (:3346-3384) reach the device-side metadata columns before the next
    filtered search; clear / batch_delete keep the mirror in step."""
    e = eng.VectorEngine()
    for i in range(10):
        e.store_embedding(f"v{i}", [float(i), 0.5])
    full = e.search_similar([5.0, 0.5], 10)
    items, total, more = e.search_paginated([5.0, 0.5], 10, skip=0, limit=3, count_total=True)
    assert [x.key for x in items] == [x.key for x in full[:3]] and total == 3 and not more
    # (the reference fetches min(skip + limit, top_k) hits and counts THOSE: total == 3, :2994-3003)
    items, total, more = e.search_paginated([5.0, 0.5], 10, skip=2, limit=3, count_total=True)
    assert [x.key for x in items] == [x.key for x in full[2:5]] and total == 5 and not more
    items, total, more = e.search_paginated([5.0, 0.5], 10, skip=4)          # skip only: top_k hits
    assert [x.key for x in items] == [x.key for x in full[4:]] and total is None and not more
    items, total, more = e.search_paginated([5.0, 0.5], 4, skip=1, limit=10, count_total=True)
    assert [x.key for x in items] == [x.key for x in full[1:4]] and total == 4 and not more
    with pytest.raises(eng.VectorError):
        e.search_paginated([], 10, limit=3)
    for i in range(5):
        e.set_entity_embedding(f"entity:{i}", [float(i), 0.0])
    items, total, more = e.search_paginated([2.0, 0.0], 5, limit=3, entities=True)   # :9232-9250
    assert len(items) == 3 and total is None and not more
    # metadata updates reach the device columns
    rows = o.fill_synthetic(400, 16, 5)
    for i in range(400):
        e.store_embedding_with_metadata(f"m{i}", rows[i], {"grp": i % 4, "tag": "old"})
    q = rows[7]
    assert [x.key for x in e.search_similar_filtered(q, 3, "grp = 3", eng.PRE_FILTER)][0] == "m7"
    e.update_metadata("m7", {"grp": 9, "tag": "new"})
    hits = e.search_similar_filtered(q, 3, "grp = 3", eng.PRE_FILTER)
    assert "m7" not in [x.key for x in hits] and all(int(x.key[1:]) % 4 == 3 for x in hits)
    assert [x.key for x in e.search_similar_filtered(q, 3, "grp = 9 AND tag = 'new'", eng.PRE_FILTER)] == ["m7"]
    e.remove_metadata_field("m7", "grp")
    assert e.search_similar_filtered(q, 3, "grp = 9", eng.PRE_FILTER) == []
    assert [x.key for x in e.search_similar_filtered(q, 3, "tag = 'new'", eng.PRE_FILTER)] == ["m7"]
    # batch delete + clear keep host and device in step
    assert e.batch_delete_embeddings([f"m{i}" for i in range(0, 400, 2)] + ["nope"]) == 200
    hits = e.search_similar(rows[8], 400)
    assert len(hits) == 200 and all(int(x.key[1:]) % 2 == 1 for x in hits)
    er, es = o.search(rows[1::2], rows[8], 5, "cosine")
    assert [x.key for x in hits[:5]] == [f"m{2 * int(i) + 1}" for i in er]
    assert e.clear() == 210 and e.count() == 0
    assert e.search_similar(rows[8], 5) == []
    e.store_embedding("again", rows[3])
    assert [x.key for x in e.search_similar(rows[3], 5)] == ["again"]
    e.close()
