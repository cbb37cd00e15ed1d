"""Tensor-core batch pre-filter (tc_prefilter_kernels.cuh; BASELINE config 4 "batch 256 queries").

A batch is nq independent search_similar_with_metric calls (vector_engine/src/lib.rs:2049-2101),
so every query of a batch must return exactly what the exact f32 kernels return: same row ids
in the same order, bit-identical scores.  Checked against the CPU oracle at sizes it finishes in
seconds and against the exact batched kernels (nm_index_set_tensor_core(0)) at larger sizes;
the tcgen05 integer GEMM itself is checked against numpy integer dot products."""
import numpy as np
import pytest

import oracle_ffi as o
from neumann_b200 import DeviceIndex
from neumann_b200.synth import synth_rows

pytestmark = pytest.mark.gpu

METRICS = ["cosine", "euclidean", "dot"]


def assert_same(got, exp, ctx=""):
    assert np.array_equal(got[0], exp[0]), f"{ctx}: rows {got[0][:8]} != {exp[0][:8]}"
    nan = np.isnan(exp[1])
    assert np.array_equal(np.isnan(got[1]), nan), ctx
    assert np.array_equal(got[1].view(np.uint32)[~nan], exp[1].view(np.uint32)[~nan]), ctx


def tc_search(idx, qs, k, metric):
    """-> results, tc queries served, fallbacks, survivors"""
    s0 = idx.stats()
    res = idx.search(qs, k, metric)
    s1 = idx.stats()
    return (res, s1.tc_queries - s0.tc_queries, s1.tc_fallbacks - s0.tc_fallbacks,
            s1.tc_survivors - s0.tc_survivors)


def quantise(q):
    """The library's int8 quantisation of a query (tc_prepare_queries_kernel)."""
    q = q.astype(np.float32)
    s = np.float32(np.abs(q).max()) / np.float32(127.0)
    return np.clip(np.rint(q / s), -127, 127).astype(np.int32)


@pytest.mark.parametrize("n,dim,nq", [(70_000, 128, 16), (66_000, 200, 21), (70_000, 768, 256),
                                      (65_600, 100, 3)])
def test_tensor_core_dot_products_are_exact(n, dim, nq):
    """tcgen05.mma kind::i8 over the TMA-staged SWIZZLE_128B tiles == integer dot products."""
    idx = DeviceIndex(dim)
    idx.fill_synthetic(n, 0x5EED0001)
    idx.set_prefilter(1)
    qs = synth_rows(nq, dim, 0x5EED1001)
    dots = idx.debug_tc_dots(qs)
    q8 = np.stack([quantise(q) for q in qs])
    rng = np.random.default_rng(1)
    rows = sorted(set([0, 1, 31, 32, 127, 128, 129, 2047, 2048, n - 1, n - 2, n // 2] +
                      [int(r) for r in rng.integers(0, n, 150)]))
    for r in rows:
        x8, _ = idx.debug_q8_row(r)
        assert np.array_equal(q8 @ x8.astype(np.int32), dots[:, r]), f"row {r}"
    idx.close()


@pytest.mark.parametrize("metric", METRICS)
@pytest.mark.parametrize("n,dim,nq,k", [(70_000, 128, 16, 10), (100_000, 96, 37, 100),
                                        (66_000, 131, 5, 7), (80_000, 768, 8, 1),
                                        (120_000, 64, 2, 1000), (70_000, 1536, 4, 20),
                                        (70_000, 8, 3, 5), (66_000, 24, 9, 3), (66_000, 2052, 3, 4),
                                        (66_000, 1, 2, 3)])
def test_tc_batch_equals_oracle(metric, n, dim, nq, k):
    rows = o.fill_synthetic(n, dim, 0x5EED0001)
    idx = DeviceIndex(dim)
    idx.load(rows)
    idx.set_prefilter(1)
    qs = o.fill_synthetic(nq, dim, 0x5EED1001)
    qs[1] = rows[n // 3]                               # a query equal to a stored row
    res, used, fell, _ = tc_search(idx, qs, k, metric)
    assert used == nq
    assert fell == 0 or dim == 1      # dim 1: every cosine is +-1, the tie flood falls back
    for i in range(nq):
        assert_same(res[i], o.search(rows, qs[i], k, metric, threads=8), f"{metric} n={n} d={dim} q{i}")
    idx.close()


@pytest.mark.parametrize("metric", METRICS)
def test_tc_batch_equals_exact_kernels_large(metric):
    """300 queries (two passes of the 256-query GEMM) over 1M x 256 against the exact batched
    kernels; the screen must also be a no-op for the result (NM_TC_SCREEN is exercised by
    test_tc_adversarial_data through rows that bypass it)."""
    n, dim, nq, k = 1_000_000, 256, 300, 50
    idx = DeviceIndex(dim)
    idx.fill_synthetic(n, 0x5EED0001)
    qs = synth_rows(nq, dim, 0x5EED1001)
    exact = idx.search(qs, k, metric)
    idx.set_prefilter(1)
    res, used, fell, surv = tc_search(idx, qs, k, metric)
    assert used == nq and fell == 0
    assert surv < nq * 40 * k, "the filter should keep a small multiple of k rows per query"
    for i in range(nq):
        assert_same(res[i], exact[i], f"{metric} q{i}")
    idx.set_tensor_core(False)
    res2, used2, _, _ = tc_search(idx, qs[:8], k, metric)
    assert used2 == 0
    for i in range(8):
        assert_same(res2[i], exact[i], f"{metric} tensor core off q{i}")
    idx.close()


@pytest.mark.parametrize("metric", METRICS)
def test_tc_adversarial_data(metric):
    """Near-duplicates inside the quantisation error, exact ties, outlier elements (coarse int8
    scale), tiny / huge / denormal magnitudes, zero rows, NaN and Inf rows, and queries that are
    zero, huge, tiny or not finite (those are redone by the exact path)."""
    n, d, k = 70_000, 96, 20
    rng = np.random.default_rng(5)
    rows = o.fill_synthetic(n, d, 21)
    q = o.fill_synthetic(1, d, 22)[0]
    rows[1000:1400] = q + rng.normal(0, 1e-4, (400, d)).astype(np.float32)
    rows[2000:2100] = rows[1000]
    rows[3000:3200, 0] = 1000.0
    rows[4000:4100] *= np.float32(1e-20)
    rows[4100:4200] *= np.float32(1e15)
    rows[5000:5050] = 0.0
    rows[6000:6100] = -rows[1000:1100]
    rows[7000:7050] *= np.float32(1e-38)
    rows[7050:7060] = np.float32(1e-45)
    rows[8000, 3] = np.nan
    rows[8001, 5] = np.inf
    rows[8002, 7] = -np.inf
    rows[8003] *= np.float32(3e38)                      # squares overflow in the reference too
    idx = DeviceIndex(d)
    idx.load(rows)
    idx.set_prefilter(1)
    qs = np.stack([q, rows[3100], rows[4050], rows[4150], np.abs(q), rows[7010],
                   q * np.float32(1e30), np.zeros(d, np.float32), q * np.float32(1e-30),
                   rows[5000] + np.float32(1.0)])
    for kk in (1, k, 500):
        res, used, fell, _ = tc_search(idx, qs, kk, metric)
        assert used == len(qs)
        for i in range(len(qs)):
            assert_same(res[i], o.search(rows, qs[i], kk, metric, threads=8),
                        f"adversarial {metric} k={kk} q{i}")
    bad = qs.copy()
    bad[2, 4] = np.nan
    bad[5, 1] = np.inf
    res, used, fell, _ = tc_search(idx, bad, k, metric)
    assert fell >= 2
    for i in range(len(bad)):
        assert_same(res[i], o.search(rows, bad[i], k, metric, threads=8), f"non-finite {metric} q{i}")
    idx.close()


@pytest.mark.parametrize("metric", ["cosine", "euclidean"])
def test_tc_sorted_corpus_falls_back_or_survives(metric):
    """Rows ordered so that every new row beats everything before it: the worst case for the
    running threshold (each phase keeps nearly all of its rows).  Whatever mix of tensor-core
    results and exact fallbacks the library chooses, the answer must not change."""
    n, d, k, nq = 200_000, 32, 10, 6
    rng = np.random.default_rng(9)
    qs = rng.normal(0, 1, (nq, d)).astype(np.float32)
    t = np.linspace(0.0, 1.0, n, dtype=np.float32)[:, None]
    noise = rng.normal(0, 1, (n, d)).astype(np.float32)
    rows = (t * qs[0][None, :] + (1 - t) * noise).astype(np.float32)   # drifts towards query 0
    idx = DeviceIndex(d)
    idx.load(rows)
    idx.set_prefilter(1)
    res, used, fell, _ = tc_search(idx, qs, k, metric)
    assert used == nq
    for i in range(nq):
        assert_same(res[i], o.search(rows, qs[i], k, metric, threads=8), f"sorted {metric} q{i}")
    idx.close()


def test_tc_many_ties_overflow_falls_back():
    """Every row identical: every (row, query) interval reaches the threshold, the kept lists
    overflow and the exact path must take over."""
    n, d = 300_000, 16
    rows = np.ones((n, d), np.float32)
    idx = DeviceIndex(d)
    idx.load(rows)
    idx.set_prefilter(1)
    qs = np.ones((3, d), np.float32)
    qs[1] *= 2
    res, used, fell, _ = tc_search(idx, qs, 10, "cosine")
    assert used == 3 and fell == 3
    for i in range(3):
        assert np.array_equal(res[i][0], np.arange(10, dtype=np.uint64))
    idx.close()


def test_tc_follows_mutations_and_routing():
    """The int8 copy follows append / update / swap-remove; small indexes, single queries and
    masked searches stay on their own paths."""
    n, d, k = 70_000, 64, 5
    rows = o.fill_synthetic(n, d, 3)
    idx = DeviceIndex(d)
    idx.load(rows[:60_000])
    idx.set_prefilter(1)
    qs = o.fill_synthetic(4, d, 4)
    _, used, _, _ = tc_search(idx, qs, k, "euclidean")
    assert used == 0, "below the row threshold the exact kernels serve the batch"
    idx.append(rows[60_000:])
    res, used, fell, _ = tc_search(idx, qs, k, "euclidean")
    assert used == 4 and fell == 0
    for i in range(4):
        assert_same(res[i], o.search(rows, qs[i], k, "euclidean", threads=8), f"append q{i}")
    rows = rows.copy()
    rows[123] = qs[2]
    idx.update(123, qs[2])
    moved = idx.swap_remove(77)
    rows[77] = rows[moved]
    rows = rows[:n - 1]
    res, used, _, _ = tc_search(idx, qs, k, "cosine")
    assert used == 4
    for i in range(4):
        assert_same(res[i], o.search(rows, qs[i], k, "cosine", threads=8), f"mutated q{i}")
    assert res[2][0][0] == 123
    _, used, _, _ = tc_search(idx, qs[0], k, "cosine")
    assert used == 0, "single queries use the 1-query pre-filter"
    idx.close()


_SCREEN_AB = r"""
import sys, json
sys.path.insert(0, sys.argv[1])
import numpy as np
from neumann_b200 import DeviceIndex
from neumann_b200.synth import synth_rows
out = {}
for metric in ("cosine", "euclidean", "dot"):
    idx = DeviceIndex(160)
    idx.fill_synthetic(400_000, 0x5EED0001)
    idx.set_prefilter(1)
    qs = synth_rows(48, 160, 0x5EED1001)
    s0 = idx.stats()
    res = idx.search(qs, 25, metric)
    s1 = idx.stats()
    out[metric] = {"survivors": int(s1.tc_survivors - s0.tc_survivors),
                   "fallbacks": int(s1.tc_fallbacks - s0.tc_fallbacks),
                   "rows": [r[0].tolist() for r in res],
                   "scores": [r[1].view(np.uint32).tolist() for r in res]}
    idx.close()
print(json.dumps(out))
"""


def test_tc_screen_never_drops_a_keeper():
    """The f32 screen of the GEMM epilogue must be a superset of the rigorous interval test.
    With NM_TC_SCREEN=0 every (row, query) entry goes through the rigorous test; the screened
    run must keep exactly the same entries: identical results AND identical survivor counts."""
    import json, os, subprocess, sys
    from pathlib import Path
    root = str(Path(__file__).resolve().parent.parent)
    runs = {}
    for screen in ("1", "0"):
        env = dict(os.environ, NM_TC_SCREEN=screen)
        p = subprocess.run([sys.executable, "-c", _SCREEN_AB, root], env=env, capture_output=True,
                           text=True, timeout=600)
        assert p.returncode == 0, p.stderr[-2000:]
        runs[screen] = json.loads(p.stdout.strip().splitlines()[-1])
    for metric in ("cosine", "euclidean", "dot"):
        a, b = runs["1"][metric], runs["0"][metric]
        assert a["fallbacks"] == 0 and b["fallbacks"] == 0
        assert a["rows"] == b["rows"] and a["scores"] == b["scores"], metric
        assert a["survivors"] == b["survivors"], (metric, a["survivors"], b["survivors"])


def test_tc_concurrent_batches():
    """Batches from several host threads at once (each call owns its workspace and stream; the
    GEMM kernels of different calls cannot share an SM) return what sequential calls return."""
    import threading
    n, d, k = 200_000, 128, 10
    idx = DeviceIndex(d)
    idx.fill_synthetic(n, 0x5EED0001)
    idx.set_prefilter(1)
    batches = [synth_rows(24, d, 0x5EED3000 + t) for t in range(6)]
    expect = [idx.search(b, k, "euclidean") for b in batches]
    got = [None] * len(batches)
    errs = []

    def work(t):
        try:
            for _ in range(5):
                got[t] = idx.search(batches[t], k, "euclidean")
        except Exception as e:  # noqa: BLE001
            errs.append(e)

    th = [threading.Thread(target=work, args=(t,)) for t in range(len(batches))]
    for x in th:
        x.start()
    for x in th:
        x.join()
    assert not errs, errs
    for t in range(len(batches)):
        for i in range(24):
            assert_same(got[t][i], expect[t][i], f"thread {t} q{i}")
    idx.close()


@pytest.mark.parametrize("n,dim,nq,k,metric", [(10_000_000, 1536, 256, 100, "euclidean")])
def test_config4_full_size_properties(n, dim, nq, k, metric):
    """BASELINE config 4 at full size through size-independent properties: scores re-derived by
    the oracle from rows fetched back from the device, total-order sortedness, the k-th hit is
    not beaten by sampled rows, planted neighbours are found, and the result is identical to
    the exact batched kernels for a sample of the queries."""
    idx = DeviceIndex(dim)
    idx.fill_synthetic(n, 0x5EED0001)
    qs = synth_rows(nq, dim, 0x5EED1001)
    planted = [123_456, 9_999_999, 0]
    for j, r in enumerate(planted):
        qs[j] = idx.get_row(r) + np.float32(1e-3) * qs[j]
    idx.set_prefilter(1)
    res, used, fell, _ = tc_search(idx, qs, k, metric)
    assert used == nq and fell == 0
    rng = np.random.default_rng(2)
    sample_rows = rng.integers(0, n, 64)
    fetched = np.stack([idx.get_row(int(r)) for r in sample_rows])
    for qi in list(range(len(planted))) + [17, 200, 255]:
        r, s = res[qi]
        assert len(r) == k and len(set(r.tolist())) == k
        hit_rows = np.stack([idx.get_row(int(x)) for x in r[:12]])
        exp = o.score_rows(hit_rows, qs[qi], metric)
        assert np.array_equal(exp.view(np.uint32), s[:12].view(np.uint32))
        keys = [(-float(sc), int(rw)) for sc, rw in zip(s, r)]
        assert keys == sorted(keys)
        assert (o.score_rows(fetched, qs[qi], metric) <= s[-1]).all() or \
            set(sample_rows[o.score_rows(fetched, qs[qi], metric) > s[-1]]).issubset(set(r.tolist()))
    for j, rw in enumerate(planted):
        assert res[j][0][0] == rw
    idx.set_tensor_core(False)
    exact = idx.search(qs[:4], k, metric)
    for i in range(4):
        assert_same(res[i], exact[i], f"config 4 q{i}")
    idx.close()


@pytest.mark.parametrize("metric", METRICS)
def test_tc_device_resident_batches_with_conditional_redo(metric):
    """nm_search_device (query + outputs in HBM, asynchronous on a caller stream) takes the
    tensor-core pre-filter for batches too.  Nothing is read back to the host there: the queries
    the device flags (zero / non-finite / overflowing lists) are redone by CONDITIONAL launches of
    the exact kernels that check the flags on the device.  Results must equal nm_search."""
    import torch
    n, d, k = 90_000, 64, 12
    rows = o.fill_synthetic(n, d, 0x5EED0001)
    rows[500:900] = rows[7]                      # ties
    idx = DeviceIndex(d)
    idx.load(rows)
    qs = o.fill_synthetic(70, d, 0x5EED1001)
    qs[3] = 0.0                                   # unusable for the int8 path
    qs[40, 5] = np.inf                            # not finite
    qs[69] *= np.float32(1e-30)                   # tiny scale
    want = idx.search(qs, k, metric)              # host path (auto mode: tensor cores + host-side redo)
    assert idx.stats().tc_queries == 70
    stream = torch.cuda.Stream()
    dq = torch.from_numpy(qs).cuda()
    d_rows = torch.zeros((70, k), dtype=torch.int64, device="cuda")
    d_scores = torch.zeros((70, k), dtype=torch.float32, device="cuda")
    d_counts = torch.zeros(70, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    for _ in range(2):                            # twice: scratch and flags are reused correctly
        idx.search_device(dq.data_ptr(), 70, k, metric, d_rows.data_ptr(), d_scores.data_ptr(),
                          d_counts.data_ptr(), stream.cuda_stream)
    assert idx.stats().tc_queries == 3 * 70
    torch.cuda.synchronize()
    gr, gs, gc = d_rows.cpu().numpy().astype(np.uint64), d_scores.cpu().numpy(), d_counts.cpu().numpy()
    for i in range(70):
        assert_same((gr[i, :gc[i]], gs[i, :gc[i]]), want[i], f"device batch {metric} q{i}")
    for i in (0, 3, 40, 69):
        assert_same(want[i], o.search(rows, qs[i], k, metric, threads=8), f"oracle {metric} q{i}")
    # all-identical rows: every list overflows, every query is redone on the device
    idx.load(np.ones((100_000, d), np.float32))
    q1 = np.ones((3, d), np.float32)
    dq1 = torch.from_numpy(q1).cuda()
    torch.cuda.synchronize()
    idx.search_device(dq1.data_ptr(), 3, k, metric, d_rows.data_ptr(), d_scores.data_ptr(),
                      d_counts.data_ptr(), stream.cuda_stream)
    torch.cuda.synchronize()
    assert np.array_equal(d_rows[:3].cpu().numpy(), np.tile(np.arange(k), (3, 1)))
    idx.release_stream(stream.cuda_stream)
    idx.close()


@pytest.mark.parametrize("metric", METRICS)
def test_tc_masked_and_filtered_batches(metric):
    """Batches with a row mask (host bitmask or device-evaluated filter) take the tensor-core
    pre-filter too: ineligible rows never enter the kept lists; the result is the oracle's on the
    eligible subset, bit for bit."""
    from neumann_b200._ffi import NM_C_LT, NM_F_CMP, NM_V_INT, NmFilterOp
    n, d, k, nq = 120_000, 72, 15, 9
    rows = o.fill_synthetic(n, d, 0x5EED0001)
    idx = DeviceIndex(d)
    idx.load(rows)
    bucket = (np.arange(n) * 2654435761 % 10).astype(np.uint64)
    idx.column_set(1, 0, np.full(n, 3, np.uint8), bucket)
    qs = o.fill_synthetic(nq, d, 0x5EED1001)
    for lim in (3, 1):                                       # 30 % and 10 % of the rows eligible
        keep = bucket < lim
        sub = np.nonzero(keep)[0]
        prog = [NmFilterOp(kind=NM_F_CMP, cmp=NM_C_LT, lit_tag=NM_V_INT, column=1, lit=lim)]
        s0 = idx.stats().tc_queries
        res_f = idx.search_filtered(qs, k, metric, prog)
        res_m = idx.search_masked(qs, k, metric, keep)
        assert idx.stats().tc_queries - s0 == 2 * nq
        for i in range(nq):
            er, es = o.search(rows[sub], qs[i], k, metric, threads=8)
            want = (sub[er.astype(np.int64)].astype(np.uint64), es)
            assert_same(res_f[i], want, f"filtered batch {metric} lim={lim} q{i}")
            assert_same(res_m[i], want, f"masked batch {metric} lim={lim} q{i}")
    # fewer eligible rows than k: everything eligible comes back, in order
    keep = np.zeros(n, bool)
    keep[[5, 70_000, 119_999]] = True
    res = idx.search_masked(qs[:2], k, metric, keep)
    for i in range(2):
        er, es = o.search(rows[keep], qs[i], k, metric)
        assert_same(res[i], (np.nonzero(keep)[0][er.astype(np.int64)].astype(np.uint64), es), "tiny subset")
    idx.close()
