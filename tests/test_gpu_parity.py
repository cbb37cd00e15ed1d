"""GPU parity tests proper: the CUDA path, called through the C ABI, against the CPU oracle on
the same seeded inputs.  Bar: identical row ids in identical rank order and bit-identical
scores (stronger than the 1e-5 relative the north star asks for)."""
import numpy as np
import pytest

import oracle_ffi as o
from neumann_b200 import DeviceIndex, NmError, _ffi

pytestmark = pytest.mark.gpu

METRICS = ["cosine", "euclidean", "dot"]


def assert_same(got, exp, ctx=""):
    gr, gs = got
    er, es = exp
    assert np.array_equal(gr, er), f"{ctx}: rows {gr[:8]} != {er[:8]}"
    # NaN payloads are canonicalised on the device; compare NaN-ness there, bits elsewhere
    nan = np.isnan(es)
    assert np.array_equal(np.isnan(gs), nan), ctx
    assert np.array_equal(gs.view(np.uint32)[~nan], es.view(np.uint32)[~nan]), f"{ctx}: score bits"


def synth_index(n, dim, seed=0x5EED0001):
    idx = DeviceIndex(dim)
    idx.fill_synthetic(n, seed)
    return idx, o.fill_synthetic(n, dim, seed)


@pytest.mark.parametrize("metric", METRICS)
@pytest.mark.parametrize("n,dim,k", [
    (1, 1, 1), (3, 3, 3), (255, 8, 5), (256, 32, 5), (257, 33, 10), (300, 100, 7), (1000, 7, 4),
    (1000, 128, 5), (5000, 768, 10), (4096, 1536, 100), (70000, 64, 100), (50000, 96, 1000),
    (40000, 31, 1024), (2000, 4096, 10), (600, 5000, 3),
])
def test_parity_synthetic(metric, n, dim, k):
    idx, rows = synth_index(n, dim)
    for qi in range(2):
        q = o.fill_synthetic(1, dim, 0x5EED1001 + qi)[0]
        (got,) = idx.search(q, k, metric)
        assert_same(got, o.search(rows, q, k, metric, threads=8), f"{metric} n={n} d={dim} k={k}")
    idx.close()


def test_parity_committed_golden_vectors():
    """CUDA path vs tests/golden/oracle_vectors.npz (frozen oracle outputs)."""
    from pathlib import Path
    g = np.load(Path(__file__).parent / "golden" / "oracle_vectors.npz")
    for name in METRICS:
        n, dim, k = (int(x) for x in g[f"{name}_shape"])
        idx = DeviceIndex(dim)
        idx.fill_synthetic(n, int(g["seed_rows"]))
        q = o.fill_synthetic(1, dim, int(g["seed_query"]))[0]
        ((r, s),) = idx.search(q, k, name)
        assert np.array_equal(r, g[f"{name}_rows"])
        assert np.array_equal(s.view(np.uint32), g[f"{name}_score_bits"])
        idx.close()


def create_test_vector(dim, seed):
    i = np.arange(dim)
    x = (seed * 31 + i * 17).astype(np.float32)
    return (np.sin((x * np.float32(0.0001)).astype(np.float32)).astype(np.float32)
            * ((seed + i).astype(np.float32) * np.float32(0.001)).astype(np.float32)).astype(np.float32)


def test_config1_store_10000_vectors_search():
    # BASELINE config 1 / vector_engine/src/lib.rs:4256-4276: 10k x 128 cosine TOP 5
    rows = np.stack([create_test_vector(128, i) for i in range(10000)])
    idx = DeviceIndex(128)
    idx.load(rows)
    q = create_test_vector(128, 5000)
    (got,) = idx.search(q, 5, "cosine")
    assert got[0][0] == 5000 and abs(got[1][0] - 1.0) < 1e-5
    assert_same(got, o.search(rows, q, 5, "cosine"))
    idx.close()


def test_reference_integer_pattern_corpus():
    # vector_engine/src/lib.rs:6534-6545 (large_scale_million_vectors), scaled to 200k rows
    n, d = 200_000, 128
    flat = (np.arange(n * d, dtype=np.int64) % 1000).astype(np.float32) / np.float32(1000.0)
    rows = flat.reshape(n, d)
    q = (np.arange(d, dtype=np.float32) / np.float32(d)).astype(np.float32)
    idx = DeviceIndex(d)
    idx.load(rows)
    for m in METRICS:
        (got,) = idx.search(q, 10, m)
        assert_same(got, o.search(rows, q, 10, m, threads=8), m)  # many exact ties (period 1000)
    idx.close()


@pytest.mark.parametrize("metric", METRICS)
def test_exact_ties_break_by_row(metric):
    base = o.fill_synthetic(3000, 64, 5)
    rows = np.concatenate([base, base[:1500], base])  # every row duplicated 2-3 times
    idx = DeviceIndex(64)
    idx.load(rows)
    q = base[17]
    (got,) = idx.search(q, 50, metric)
    assert_same(got, o.search(rows, q, 50, metric), metric)
    assert list(got[0][:3]) == [17, 3017, 4517]
    idx.close()


def test_special_values_rank_like_the_oracle():
    d = 16
    rows = o.fill_synthetic(600, d, 9)
    rows[5] = 0.0                      # zero-norm row: cosine 0.0 exactly
    rows[6] = -0.0
    rows[7, 3] = np.nan                # NaN score ranks last
    rows[8, 0] = np.inf
    rows[9] = 1e30                     # overflow to inf inside sum of squares
    rows[10] = 1e-30                   # underflow to denormal/zero
    rows[11] = np.float32(1.1754944e-38) / 2  # denormals are kept (no FTZ)
    rows[12] = -rows[13]
    idx = DeviceIndex(d)
    idx.load(rows)
    for m in METRICS:
        for q in (o.fill_synthetic(1, d, 77)[0], np.full(d, 1e-30, np.float32), rows[11].copy()):
            (got,) = idx.search(q, 600, m)
            assert_same(got, o.search(rows, q, 600, m), m)
    # zero query: cosine scores every row 0.0 (guard in cosine_similarity), order = row order
    (got,) = idx.search(np.zeros(d, np.float32), 5, "cosine")
    assert list(got[0]) == [0, 1, 2, 3, 4] and not got[1].any()
    idx.close()


def test_negative_zero_score_is_preserved():
    rows = np.array([[1e-30, 0.0], [-1e-30, 0.0], [0.0, 1.0]], np.float32)
    q = np.array([1e-30, 0.0], np.float32)   # products underflow to +0.0 / -0.0
    idx = DeviceIndex(2)
    idx.load(rows)
    (got,) = idx.search(q, 3, "dot")
    exp = o.search(rows, q, 3, "dot")
    assert_same(got, exp)
    assert np.array_equal(got[1].view(np.uint32), exp[1].view(np.uint32))
    idx.close()


@pytest.mark.parametrize("metric", METRICS)
def test_adversarial_ascending_scores_exercise_pruning(metric):
    """Scores increase with the row index, so every row beats the running threshold and the
    candidate buffer is pruned over and over."""
    n, d = 60000, 8
    t = np.linspace(0.01, 1.5, n, dtype=np.float32)
    rows = np.zeros((n, d), np.float32)
    if metric == "euclidean":
        rows[:, 0] = t[::-1]          # distance to q=0 shrinks with row
        q = np.zeros(d, np.float32)
    elif metric == "dot":
        rows[:, 0] = t
        q = np.eye(1, d, 0, dtype=np.float32)[0]
    else:
        rows[:, 0] = np.cos(t[::-1]); rows[:, 1] = np.sin(t[::-1])
        q = np.eye(1, d, 0, dtype=np.float32)[0]
    idx = DeviceIndex(d)
    idx.load(rows)
    for k in (1, 10, 1000):
        (got,) = idx.search(q, k, metric)
        assert_same(got, o.search(rows, q, k, metric, threads=8), f"{metric} k={k}")
    idx.close()


def test_very_long_rows_stream_the_query():
    """dim = 40000: the query no longer fits next to the stage ring, so even a single query is
    served by the batched kernels (query streamed chunk by chunk)."""
    n, d = 300, 40000
    idx, rows = synth_index(n, d)
    q = o.fill_synthetic(1, d, 3)[0]
    for m in METRICS:
        (got,) = idx.search(q, 5, m)
        assert_same(got, o.search(rows, q, 5, m), m)
    idx.close()


def test_k_edge_cases_and_errors():
    idx, rows = synth_index(100, 24)
    q = o.fill_synthetic(1, 24, 1)[0]
    (got,) = idx.search(q, 1024, "cosine")          # k > n returns all rows, sorted
    assert len(got[0]) == 100
    assert_same(got, o.search(rows, q, 1024, "cosine"))
    with pytest.raises(NmError) as ei:
        idx.search(q, 0, "cosine")
    assert ei.value.code == _ffi.NM_ERR_INVALID_TOP_K
    with pytest.raises(NmError) as ei:
        idx.search(np.zeros(25, np.float32), 3, "cosine")
    assert ei.value.code == _ffi.NM_ERR_DIMENSION_MISMATCH
    idx.clear()
    assert idx.rows == 0
    (got,) = idx.search(q, 5, "cosine")
    assert len(got[0]) == 0
    idx.close()


@pytest.mark.parametrize("metric", METRICS)
def test_large_k_chained_passes(metric):
    """k > 1024 is served by chained passes under a key ceiling; result must equal the
    oracle's single sort, including across pass boundaries and exact ties."""
    base = o.fill_synthetic(6000, 40, 61)
    rows = np.concatenate([base, base[:3000]])      # ties straddle the pass boundaries
    idx = DeviceIndex(40)
    idx.load(rows)
    q = o.fill_synthetic(1, 40, 62)[0]
    for k in (1025, 2048, 3000, 9000, 20000):
        (got,) = idx.search(q, k, metric)
        assert len(got[0]) == min(k, 9000)
        assert_same(got, o.search(rows, q, k, metric, threads=4), f"{metric} k={k}")
    idx.close()


def test_multi_query_batch_equals_single_queries():
    idx, rows = synth_index(20000, 96)
    qs = o.fill_synthetic(7, 96, 0xABC)
    for m in METRICS:
        res = idx.search(qs, 10, m)
        assert len(res) == 7
        for i in range(7):
            assert_same(res[i], o.search(rows, qs[i], 10, m, threads=4), f"{m} q{i}")
    idx.close()


@pytest.mark.parametrize("metric", METRICS)
@pytest.mark.parametrize("n,dim,nq,k", [
    (3000, 64, 2, 5), (6000, 64, 3, 5), (3000, 64, 8, 5), (20000, 96, 17, 10), (5000, 768, 70, 10), (1000, 40, 256, 100),
    (70000, 32, 33, 1000), (300, 24, 9, 1024), (9000, 100, 20, 7), (4000, 1536, 64, 100),
    (2500, 13, 12, 4),
])
def test_batched_queries_share_one_pass(metric, n, dim, nq, k):
    """nq >= 2 goes through score_batch_kernel + select_batch_kernel (BASELINE config 4 shape);
    every (query, row) score and every per-query ranking must equal nq independent searches.
    dim % 8 != 0 with dot/cosine and dim % 32 != 0 with Euclidean cover the routing rules."""
    base = o.fill_synthetic(n, dim, 0x5EED0001)
    rows = base.copy()
    rows[n // 2:n // 2 + n // 10] = base[:n // 10]      # exact ties
    idx = DeviceIndex(dim)
    idx.set_prefilter(0)    # this test is about the exact batched kernels (70000 rows would
    idx.load(rows)          # otherwise take the tensor-core pre-filter, tests/test_gpu_tc.py)
    qs = o.fill_synthetic(nq, dim, 0xBEEF)
    qs[1] = rows[7]                                      # a query equal to a stored row
    if metric != "euclidean" and nq > 2:
        qs[2] = 0.0                                      # zero query: cosine scores all 0.0
    res = idx.search(qs, k, metric)
    launches_batched = idx.stats().scan_launches
    assert len(res) == nq
    for i in range(nq):
        assert_same(res[i], o.search(rows, qs[i], k, metric, threads=8), f"{metric} q{i}")
    idx.set_batching(False)
    res2 = idx.search(qs, k, metric)
    for i in range(nq):
        assert np.array_equal(res[i][0], res2[i][0])
        assert np.array_equal(res[i][1].view(np.uint32), res2[i][1].view(np.uint32))
    if metric == "euclidean" or dim % 8 == 0:        # batched: prepare + score + select per pass
        per_pass = 64 if (metric == "euclidean" and nq > 16) else 16
        assert launches_batched == 3 * -(-nq // per_pass)
    else:                                             # dot/cosine tail rule: one scan per query
        assert launches_batched == nq
    idx.close()


@pytest.mark.parametrize("metric", METRICS)
def test_masked_search_equals_oracle_on_the_subset(metric):
    """nm_search_masked == the oracle run on only the eligible rows (search_with_pre_filter,
    vector_engine/src/lib.rs:3514-3557), for dense, sparse, clustered, single-row and empty
    masks (clustered masks exercise the 'skip whole row blocks' path)."""
    n, d, k = 40000, 48, 20
    idx, rows = synth_index(n, d)
    q = o.fill_synthetic(1, d, 0x5EED1001)[0]
    rng = np.random.default_rng(3)
    masks = {
        "half": rng.random(n) < 0.5,
        "one_percent": rng.random(n) < 0.01,
        "clustered": (np.arange(n) // 1000) % 7 == 3,
        "single": np.arange(n) == 31337,
        "fewer_than_k": np.isin(np.arange(n), [5, 300, 39999]),
        "all": np.ones(n, bool),
    }
    for name, m in masks.items():
        sub = np.nonzero(m)[0]
        er, es = o.search(rows[sub], q, k, metric, threads=4)
        ((gr, gs),) = idx.search_masked(q, k, metric, m)
        assert np.array_equal(gr, sub[er.astype(np.int64)].astype(np.uint64)), (name, gr[:5])
        assert np.array_equal(gs.view(np.uint32), es.view(np.uint32)), name
    ((gr, gs),) = idx.search_masked(q, k, metric, np.zeros(n, bool))
    assert len(gr) == 0
    idx.close()


def test_mirror_mutations_load_append_update_swap_remove():
    d = 40
    host = o.fill_synthetic(5000, d, 21)
    idx = DeviceIndex(d)
    idx.load(host[:3000])
    idx.append(host[3000:4000])
    idx.append(host[4000:])
    assert idx.rows == 5000
    q = o.fill_synthetic(1, d, 22)[0]
    (got,) = idx.search(q, 20, "cosine")
    assert_same(got, o.search(host, q, 20, "cosine"))
    # overwrite one row with the query itself -> becomes the top hit
    host = host.copy()
    host[1234] = q
    idx.update(1234, q)
    (got,) = idx.search(q, 20, "cosine")
    assert got[0][0] == 1234
    assert_same(got, o.search(host, q, 20, "cosine"))
    assert np.array_equal(idx.get_row(1234), q)
    # delete = move the last row into the hole; the caller mirrors the swap
    moved = idx.swap_remove(1234)
    assert moved == 4999
    host[1234] = host[4999]
    host = host[:4999]
    assert idx.rows == 4999
    for m in METRICS:
        (got,) = idx.search(q, 20, m)
        assert_same(got, o.search(host, q, 20, m), m)
    assert idx.swap_remove(4998) == 4998  # removing the last row itself
    (got,) = idx.search(q, 5, "dot")
    assert_same(got, o.search(host[:4998], q, 5, "dot"))
    idx.close()


def test_load_from_pageable_memory_with_ragged_dim():
    # dim % 4 != 0 -> padded device pitch; dim % 8 != 0 -> scalar tail in the lane tree
    for d in (5, 13, 97, 770):
        host = o.fill_synthetic(3000, d, d)
        idx = DeviceIndex(d)
        idx.load(host)
        assert np.array_equal(idx.get_row(2999), host[2999])
        q = o.fill_synthetic(1, d, 1000 + d)[0]
        for m in METRICS:
            (got,) = idx.search(q, 9, m)
            assert_same(got, o.search(host, q, 9, m), f"d={d} {m}")
        idx.close()


def test_search_device_resident_buffers():
    import torch
    idx, rows = synth_index(30000, 128)
    q = o.fill_synthetic(3, 128, 5)
    dq = torch.from_numpy(q).cuda()
    k = 10
    d_rows = torch.zeros((3, k), dtype=torch.int64, device="cuda")
    d_scores = torch.zeros((3, k), dtype=torch.float32, device="cuda")
    d_counts = torch.zeros(3, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    idx.search_device(dq.data_ptr(), 3, k, "cosine", d_rows.data_ptr(), d_scores.data_ptr(),
                      d_counts.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    for i in range(3):
        er, es = o.search(rows, q[i], k, "cosine", threads=4)
        assert int(d_counts[i]) == k
        assert np.array_equal(d_rows[i].cpu().numpy().astype(np.uint64), er)
        assert np.array_equal(d_scores[i].cpu().numpy().view(np.uint32), es.view(np.uint32))
    idx.close()


@pytest.mark.parametrize("n,dim,k", [(300_000, 128, 10), (20_000, 96, 100), (700, 64, 5), (2_000_000, 64, 1)])
def test_pipelined_device_searches_overlap_and_match(n, dim, k):
    """nm_index_set_pipelining(1): consecutive asynchronous single-query calls on one caller stream
    overlap (programmatic dependent launch, alternating scratch sets).  Every result must equal
    the strictly serial call; outputs written into ONE shared buffer must end up holding the last
    query's result; a mutation issued right behind the searches must wait for all of them."""
    import torch
    idx, rows = synth_index(n, dim)
    nq = 40
    q = o.fill_synthetic(nq, dim, 0x77)
    dq = torch.from_numpy(q).cuda()
    stream = torch.cuda.Stream()
    d_rows = torch.zeros((nq, k), dtype=torch.int64, device="cuda")
    d_scores = torch.zeros((nq, k), dtype=torch.float32, device="cuda")
    d_counts = torch.zeros(nq, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()

    def run(metric, into_one=False):
        for i in range(nq):
            j = 0 if into_one else i
            idx.search_device(dq[i].data_ptr(), 1, k, metric, d_rows[j].data_ptr(),
                              d_scores[j].data_ptr(), d_counts[j].data_ptr(), stream.cuda_stream)

    for metric in METRICS:
        idx.set_pipelining(False)
        run(metric)
        torch.cuda.synchronize()
        ref = (d_rows.cpu().numpy().copy(), d_scores.cpu().numpy().copy(), d_counts.cpu().numpy().copy())
        d_rows.zero_(); d_scores.zero_(); d_counts.zero_()
        torch.cuda.synchronize()
        idx.set_pipelining(True)
        s0 = idx.stats().scan_launches
        run(metric)
        run(metric, into_one=True)
        assert idx.stats().scan_launches - s0 == 2 * nq
        torch.cuda.synchronize()
        got = (d_rows.cpu().numpy(), d_scores.cpu().numpy(), d_counts.cpu().numpy())
        assert np.array_equal(got[2][1:], ref[2][1:])
        assert np.array_equal(got[0][1:], ref[0][1:]), metric
        assert np.array_equal(got[1][1:].view(np.uint32), ref[1][1:].view(np.uint32)), metric
        # slot 0 was overwritten by every query of the second run, in call order: last one wins
        assert np.array_equal(got[0][0], ref[0][nq - 1]) and got[2][0] == ref[2][nq - 1]
        for i in (1, nq - 1):
            er, es = o.search(rows, q[i], k, metric, threads=8)
            m = int(ref[2][i])
            assert np.array_equal(ref[0][i][:m].astype(np.uint64), er)
            assert np.array_equal(ref[1][i][:m].view(np.uint32), es.view(np.uint32))
    # a mutation right behind a burst of pipelined searches waits for them (no torn scan): the
    # query itself is planted as the last row while 40 searches are in flight; all 40 must still
    # report the corpus as it was
    idx.set_pipelining(False)
    run("cosine")
    torch.cuda.synchronize()
    before = d_rows.cpu().numpy().copy()
    d_rows.zero_()
    torch.cuda.synchronize()
    idx.set_pipelining(True)
    run("cosine")
    idx.update(n - 1, q[3])
    torch.cuda.synchronize()
    assert np.array_equal(d_rows.cpu().numpy(), before)
    (after,) = idx.search(q[3], k, "cosine")
    assert o.compute_score(q[3], q[3], "cosine").view(np.uint32) == after[1][0].view(np.uint32)
    assert n - 1 in after[0]
    idx.release_stream(stream.cuda_stream)
    idx.close()


def test_concurrent_single_queries_are_coalesced():
    """16 host threads issue single-query nm_search calls at once (the reference's serving
    pattern): calls queue behind the running scan and share corpus passes; every result must
    equal the isolated search."""
    import threading
    n, d, k = 400_000, 128, 10
    idx, rows = synth_index(n, d)
    qs = o.fill_synthetic(16 * 12, d, 0xC0A1)
    expected = [o.search(rows, q, k, "cosine", threads=16) for q in qs[:48]]
    results = [None] * len(qs)
    errors = []

    def worker(tid):
        try:
            for j in range(12):
                i = tid * 12 + j
                (results[i],) = idx.search(qs[i], k, "cosine")
        except Exception as e:  # noqa: BLE001
            errors.append(repr(e))

    ts = [threading.Thread(target=worker, args=(t,)) for t in range(16)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errors, errors[:2]
    st = idx.stats()
    assert st.coalesced_queries == len(qs)
    assert st.coalesced_batches < st.coalesced_queries      # some calls really shared a pass
    for i in range(48):
        assert_same(results[i], expected[i], f"coalesced q{i}")
    idx.set_coalescing(1)
    (alone,) = idx.search(qs[5], k, "cosine")
    assert_same(alone, expected[5])
    assert idx.stats().coalesced_queries == len(qs)          # coalescer bypassed
    idx.close()


def test_stats_counters():
    idx, _ = synth_index(1000, 64)
    q = o.fill_synthetic(2, 64, 5)
    idx.set_batching(False)              # two queries would otherwise share one batched pass
    idx.search(q, 3, "dot")
    s = idx.stats()
    assert s.searches == 2 and s.scan_launches == 2
    assert s.rows_scanned == 2000 and s.bytes_streamed == 2 * 1000 * 64 * 4
    assert s.last_scan_ms == 0           # timing events are recorded only while profiling is on
    idx.set_profiling(True)
    idx.search(q, 3, "dot")
    assert idx.stats().last_scan_ms > 0
    idx.set_profiling(False)
    idx.search(q, 3, "dot")
    assert idx.stats().last_scan_ms == 0
    idx.close()


@pytest.mark.parametrize("mode", [2, 0])
@pytest.mark.parametrize("metric", METRICS)
def test_parity_gate_1000_queries_with_tie_stress(metric, mode):
    """SURVEY 8d parity gate: >= 1000 queries per metric on a corpus with 1 % duplicated rows
    (exact ties); ids position by position, scores bit for bit.  The 1000 queries go through the
    default batch path (mode 2: tensor-core pre-filter, copy built by the first batch) and through
    the exact batched kernels (mode 0); the first 40 are repeated through single-query scans."""
    n, d, k = 120_000, 128, 10
    rows = o.fill_synthetic(n, d, 0x5EED0001)
    rng = np.random.default_rng(11)
    dup = rng.choice(n, n // 100, replace=False)
    rows[dup] = rows[(dup * 7 + 13) % n]
    idx = DeviceIndex(d)
    idx.set_prefilter(mode)
    idx.load(rows)
    qs = o.fill_synthetic(1000, d, 0x5EED1001)
    qs[::50] = rows[dup[:20]]            # 20 queries that hit a tie group exactly
    res = idx.search(qs, k, metric)
    assert (idx.stats().tc_queries == 1000) == (mode == 2)
    for i in range(1000):
        assert_same(res[i], o.search(rows, qs[i], k, metric, threads=16), f"{metric} q{i}")
    idx.set_batching(False)
    res1 = idx.search(qs[:40], k, metric)
    for i in range(40):
        assert np.array_equal(res1[i][0], res[i][0])
        assert np.array_equal(res1[i][1].view(np.uint32), res[i][1].view(np.uint32))
    idx.close()


# ---- BASELINE configs at (or near) full size -------------------------------------------------
def test_config2_1m_x_768_cosine_top10_vs_oracle():
    n, d = 1_000_000, 768
    idx, rows = synth_index(n, d)
    q = o.fill_synthetic(1, d, 0x5EED1001)[0]
    (got,) = idx.search(q, 10, "cosine")
    assert_same(got, o.search(rows, q, 10, "cosine", threads=16))
    idx.close()


def _full_size_properties(n, d, k, metric):
    """Size-independent checks where the oracle cannot score the whole corpus in seconds:
    (1) every returned score is bit-equal to the oracle's score of that very row (row fetched
    back from the device), (2) results are sorted by the total order, (3) no row of a large
    random sample beats the k-th hit, (4) a planted copy of the query becomes the top hit with
    the oracle's self-score, (5) the search is idempotent."""
    idx = DeviceIndex(d)
    idx.fill_synthetic(n, 0x5EED0001)
    q = o.fill_synthetic(1, d, 0x5EED1001)[0]
    ((r, s),) = idx.search(q, k, metric)
    assert len(r) == k
    for i in range(0, k, max(1, k // 16)):
        row = idx.get_row(int(r[i]))
        assert np.array_equal(row, o.fill_synthetic(1, d, 0x5EED0001, row_offset=int(r[i]))[0])
        assert o.compute_score(q, row, metric).view(np.uint32) == s[i].view(np.uint32)
    import np_ref
    ordk = np_ref.orderable(s).astype(np.int64)
    assert all((ordk[i] > ordk[i + 1]) or (ordk[i] == ordk[i + 1] and r[i] < r[i + 1])
               for i in range(k - 1))
    rng = np.random.default_rng(0)
    for start in rng.integers(0, n - 50_000, 4):
        block = o.fill_synthetic(50_000, d, 0x5EED0001, row_offset=int(start))
        sc = o.score_rows(block, q, metric)
        better = np_ref.orderable(sc).astype(np.int64) > ordk[-1]
        rows_better = set((np.nonzero(better)[0] + int(start)).tolist())
        assert rows_better <= set(int(x) for x in r), "a sampled row beats the k-th hit"
    plant = n - 12345
    idx.update(plant, q)
    ((r2, s2),) = idx.search(q, k, metric)
    assert r2[0] == plant
    assert s2[0].view(np.uint32) == o.compute_score(q, q, metric).view(np.uint32)
    ((r3, s3),) = idx.search(q, k, metric)
    assert np.array_equal(r2, r3) and np.array_equal(s2.view(np.uint32), s3.view(np.uint32))
    idx.close()


def test_config3_10m_x_768_cosine_top10_properties():
    _full_size_properties(10_000_000, 768, 10, "cosine")


def test_config4_shape_10m_x_1536_l2_top100_properties():
    _full_size_properties(10_000_000, 1536, 100, "euclidean")


def test_config4_batch_256_queries_full_size():
    """BASELINE config 4 as stated: 10M x 1536, Euclidean, TOP 100, ONE batch of 256 queries
    (batched kernels).  Checked through size-independent properties on a spread of queries:
    returned scores re-derived by the oracle from the rows themselves, total-order sortedness,
    no row of a sampled block beats the 100th hit, and equality with the single-query scan."""
    import np_ref
    n, d, k, nq = 10_000_000, 1536, 100, 256
    idx = DeviceIndex(d)
    idx.set_prefilter(0)                                # the exact batched kernels
    idx.fill_synthetic(n, 0x5EED0001)
    qs = o.fill_synthetic(nq, d, 0x5EED1001)
    res = idx.search(qs, k, "euclidean")
    assert idx.stats().scan_launches == 3 * 4           # 4 passes of 64 queries
    block_start = 6_543_210
    block = o.fill_synthetic(40_000, d, 0x5EED0001, row_offset=block_start)
    for qi in (0, 63, 64, 200, 255):
        r, s = res[qi]
        assert len(r) == k
        ordk = np_ref.orderable(s).astype(np.int64)
        assert all((ordk[i] > ordk[i + 1]) or (ordk[i] == ordk[i + 1] and r[i] < r[i + 1])
                   for i in range(k - 1))
        for i in (0, 1, 50, 99):
            row = o.fill_synthetic(1, d, 0x5EED0001, row_offset=int(r[i]))[0]
            assert o.compute_score(qs[qi], row, "euclidean").view(np.uint32) == s[i].view(np.uint32)
        sc = o.score_rows(block, qs[qi], "euclidean")
        better = np.nonzero(np_ref.orderable(sc).astype(np.int64) > ordk[-1])[0] + block_start
        assert set(better.tolist()) <= set(int(x) for x in r)
    idx.set_batching(False)
    (single,) = idx.search(qs[200], k, "euclidean")
    assert np.array_equal(single[0], res[200][0])
    assert np.array_equal(single[1].view(np.uint32), res[200][1].view(np.uint32))
    idx.close()


# ---- full-corpus oracle at the BASELINE sizes (chunked: 1M-row chunks -> per-chunk top-k ->
#      nmo_merge_top_k, the reference's merge_top_k, distributed.rs:413-433) ------------------
def _chunked_oracle(n, d, queries, k, metric, seed=0x5EED0001, chunk=1_000_000, threads=16,
                    index=None):
    """Per query the exact top-k of the WHOLE synthetic corpus, every row scored by the oracle.
    Chunks come from the host generator; when `index` is given each chunk is also compared bit for
    bit with the rows read back from the device mirror (nm_index_get_rows)."""
    lists = [([], []) for _ in queries]
    for c0 in range(0, n, chunk):
        m = min(chunk, n - c0)
        rows = o.fill_synthetic(m, d, seed, row_offset=c0, threads=threads)
        if index is not None:
            assert np.array_equal(index.get_rows(c0, m).view(np.uint32), rows.view(np.uint32))
        for qi, q in enumerate(queries):
            r, s = o.search(rows, q, k, metric, threads=threads)
            lists[qi][0].append(r + np.uint64(c0))
            lists[qi][1].append(s)
    return [o.merge_top_k(lr, ls, k) for lr, ls in lists]


def test_get_rows_matches_generator_and_get_row():
    idx, rows = synth_index(70_001, 131)
    got = idx.get_rows(0, 70_001)
    assert np.array_equal(got.view(np.uint32), rows.view(np.uint32))
    part = idx.get_rows(65_000, 1_234)
    assert np.array_equal(part, rows[65_000:66_234])
    assert np.array_equal(idx.get_row(66_000), part[1_000])
    with pytest.raises(NmError):
        idx.get_rows(70_000, 2)
    idx.close()


def test_config3_full_oracle():
    """BASELINE config 3 (headline): 10M x 768 cosine TOP 10, 4 queries, EVERY row scored by the
    oracle — identical ids in identical order and bit-identical scores (VE:2013-2034)."""
    n, d, k = 10_000_000, 768, 10
    idx = DeviceIndex(d)
    idx.fill_synthetic(n, 0x5EED0001)
    qs = o.fill_synthetic(4, d, 0x5EED1001)
    exp = _chunked_oracle(n, d, qs, k, "cosine", index=idx)
    idx.set_batching(False)
    for qi in range(4):
        (got,) = idx.search(qs[qi], k, "cosine")
        assert_same(got, exp[qi], f"config 3 query {qi}")
    idx.set_batching(True)
    for qi, got in enumerate(idx.search(qs, k, "cosine")):      # the same 4 as one batch (default:
        assert_same(got, exp[qi], f"config 3 batched query {qi}")  # tensor-core pre-filter)
    assert idx.stats().tc_queries == 4
    idx.set_prefilter(0)
    for qi, got in enumerate(idx.search(qs, k, "cosine")):      # and through the exact batched kernels
        assert_same(got, exp[qi], f"config 3 exact batched query {qi}")
    idx.close()


def test_config4_full_oracle():
    """BASELINE config 4: 10M x 1536 Euclidean TOP 100, a batch of 256 queries; queries 0, 100 and
    255 of the batch are checked against the oracle over EVERY row, through the default batch path
    and through the exact batched kernels."""
    n, d, k, nq = 10_000_000, 1536, 100, 256
    idx = DeviceIndex(d)
    idx.fill_synthetic(n, 0x5EED0001)
    qs = o.fill_synthetic(nq, d, 0x5EED2001)
    pick = [0, 100, 255]
    exp = _chunked_oracle(n, d, qs[pick], k, "euclidean")
    res = idx.search(qs, k, "euclidean")                       # default path: tensor-core pre-filter
    assert idx.stats().tc_queries == nq
    for j, qi in enumerate(pick):
        assert_same(res[qi], exp[j], f"config 4 default path, query {qi}")
    idx.set_tensor_core(False)                                  # exact batched kernels
    res64 = idx.search(qs[64:128], k, "euclidean")
    assert_same(res64[100 - 64], exp[1], "config 4 exact batched kernels, query 100")
    for qi in range(64, 128):                                   # and the two paths agree on all 64
        assert_same(res64[qi - 64], res[qi], f"default vs exact batched, query {qi}")
    idx.close()
