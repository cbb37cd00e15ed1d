"""Mixed workload on ONE index from many host threads (model: the reference's
stress_tests/tests/mixed_workload_stress.rs:291-307 and VE:5615-5711): single queries (coalesced
into shared passes), batches (tensor-core pre-filter, int8 copy built on the fly and kept up to
date), device-evaluated filters, pipelined device-resident searches — while another thread keeps
appending rows and their metadata.  Nothing may crash or tear: every hit must carry the exact
score of the row it names, lists must be sorted, filtered hits must pass the filter."""
import threading

import numpy as np
import pytest

import np_ref
import oracle_ffi as o
from neumann_b200 import DeviceIndex
from neumann_b200._ffi import NM_C_LT, NM_F_CMP, NM_V_INT, NmFilterOp

pytestmark = pytest.mark.gpu


def test_mixed_workload_stress():
    import torch
    d, k = 64, 8
    total, start = 260_000, 80_000
    rows = o.fill_synthetic(total, d, 0x5EED0001)
    bucket = (np.arange(total) * 7919 % 10).astype(np.uint64)
    idx = DeviceIndex(d)
    idx.load(rows[:start])
    idx.column_set(1, 0, np.full(start, NM_V_INT, np.uint8), bucket[:start])
    qs = o.fill_synthetic(64, d, 0x5EED1001)
    prog = [NmFilterOp(kind=NM_F_CMP, cmp=NM_C_LT, lit_tag=NM_V_INT, column=1, lit=3)]
    stop = threading.Event()
    errors, counts = [], {"single": 0, "batch": 0, "filtered": 0, "device": 0, "appends": 0}
    lock = threading.Lock()

    def check(q, r, s, metric, what, must_pass_filter=False):
        assert len(r) == len(s) and len(r) <= k, what
        rr = r.astype(np.int64)
        assert (rr < total).all(), what
        want = o.score_rows(rows[rr], q, metric) if len(rr) else np.zeros(0, np.float32)
        assert np.array_equal(s.view(np.uint32), want.view(np.uint32)), what
        key = np_ref.orderable(s).astype(np.int64)
        assert all(key[i] > key[i + 1] or (key[i] == key[i + 1] and rr[i] < rr[i + 1])
                   for i in range(len(rr) - 1)), what
        if must_pass_filter:
            assert (bucket[rr] < 3).all(), what

    def guarded(fn):
        def run():
            try:
                fn()
            except Exception as e:  # noqa: BLE001
                import traceback
                errors.append(traceback.format_exc()[-600:] + repr(e))
                stop.set()
        return run

    def appender():
        n = start
        rng = np.random.default_rng(3)
        while n < total and not stop.is_set():
            m = int(min(total - n, rng.integers(1, 12_000)))
            idx.append(rows[n:n + m])
            idx.column_set(1, n, np.full(m, NM_V_INT, np.uint8), bucket[n:n + m])
            n += m
            with lock:
                counts["appends"] += 1
        stop.set()

    def single(seed):
        def run():
            i = seed
            while not stop.is_set():
                m = ("cosine", "euclidean", "dot")[i % 3]
                ((r, s),) = idx.search(qs[i % 64], k, m)
                check(qs[i % 64], r, s, m, f"single {m}")
                i += 7
                with lock:
                    counts["single"] += 1
        return run

    def batch():
        i = 0
        while not stop.is_set():
            m = ("euclidean", "cosine")[i % 2]
            sel = [(i + j) % 64 for j in range(9)]
            for j, (r, s) in zip(sel, idx.search(qs[sel], k, m)):
                check(qs[j], r, s, m, f"batch {m}")
            i += 1
            with lock:
                counts["batch"] += 1

    def filtered():
        i = 0
        while not stop.is_set():
            ((r, s),) = idx.search_filtered(qs[i % 64], k, "cosine", prog)
            check(qs[i % 64], r, s, "cosine", "filtered", must_pass_filter=True)
            i += 1
            with lock:
                counts["filtered"] += 1

    def device():
        stream = torch.cuda.Stream()
        dq = torch.from_numpy(qs).cuda()
        d_rows = torch.zeros((16, k), dtype=torch.int64, device="cuda")
        d_scores = torch.zeros((16, k), dtype=torch.float32, device="cuda")
        d_counts = torch.zeros(16, dtype=torch.int32, device="cuda")
        torch.cuda.synchronize()
        idx.set_pipelining(True)
        i = 0
        while not stop.is_set():
            for j in range(16):
                idx.search_device(dq[(i + j) % 64].data_ptr(), 1, k, "dot", d_rows[j].data_ptr(),
                                  d_scores[j].data_ptr(), d_counts[j].data_ptr(), stream.cuda_stream)
            stream.synchronize()
            gr, gs, gc = d_rows.cpu().numpy(), d_scores.cpu().numpy(), d_counts.cpu().numpy()
            for j in range(16):
                check(qs[(i + j) % 64], gr[j, :gc[j]].astype(np.uint64), gs[j, :gc[j]], "dot", "device pipelined")
            i += 16
            with lock:
                counts["device"] += 16
        idx.release_stream(stream.cuda_stream)

    ts = [threading.Thread(target=guarded(f)) for f in
          (appender, single(0), single(1), single(2), batch, filtered, device)]
    for t in ts:
        t.start()
    for t in ts:
        t.join(timeout=300)
    assert not errors, errors[0]
    assert idx.rows == total and all(v > 0 for v in counts.values()), counts
    st = idx.stats()
    assert st.tc_queries > 0 and st.filter_masks_built > 0 and st.coalesced_queries > 0
    # the final state is exactly the corpus
    for m in ("cosine", "euclidean", "dot"):
        ((r, s),) = idx.search(qs[5], k, m)
        er, es = o.search(rows, qs[5], k, m, threads=8)
        assert np.array_equal(r, er) and np.array_equal(s.view(np.uint32), es.view(np.uint32))
    keep = bucket < 3
    sub = np.nonzero(keep)[0]
    ((r, s),) = idx.search_filtered(qs[5], k, "cosine", prog)
    er, es = o.search(rows[sub], qs[5], k, "cosine", threads=8)
    assert np.array_equal(r, sub[er.astype(np.int64)].astype(np.uint64))
    idx.close()
