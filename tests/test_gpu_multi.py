"""Multi-GPU parity (skipped when fewer than 2 devices are visible): in-process row-range
shards with the host-side merge, and one-process-per-GPU shards with the NCCL all-gather +
device merge."""
import os
import socket

import numpy as np
import pytest

import oracle_ffi as o
from neumann_b200 import DeviceIndex, device_count

pytestmark = pytest.mark.gpu


def _need(n):
    if device_count() < n:
        pytest.skip(f"needs {n} GPUs, {device_count()} visible")


@pytest.mark.parametrize("metric", ["cosine", "euclidean", "dot"])
def test_in_process_two_device_shards(metric):
    _need(2)
    n, d = 30011, 72
    rows = o.fill_synthetic(n, d, 0x5EED0001)
    rows[5] = rows[29000]  # tie across shards
    idx = DeviceIndex(d, devices=[0, 1])
    idx.load(rows)
    qs = np.stack([rows[29000], o.fill_synthetic(1, d, 3)[0]])
    res = idx.search(qs, 20, metric)
    for i in range(2):
        er, es = o.search(rows, qs[i], 20, metric, threads=4)
        assert np.array_equal(res[i][0], er)
        assert np.array_equal(res[i][1].view(np.uint32), es.view(np.uint32))
    idx.close()


@pytest.mark.parametrize("metric", ["cosine", "euclidean"])
def test_in_process_two_device_shards_tensor_core(metric):
    """Batches on a pre-filtered two-device index: each shard's hits come from the tensor-core
    pre-filter, the host merges them (merge_top_k semantics); identical to the oracle."""
    _need(2)
    n, d, k = 150_001, 72, 20
    rows = o.fill_synthetic(n, d, 0x5EED0001)
    rows[5] = rows[140_000]  # tie across shards
    idx = DeviceIndex(d, devices=[0, 1])
    idx.load(rows)
    idx.set_prefilter(1)
    qs = np.concatenate([rows[140_000:140_001], o.fill_synthetic(6, d, 3)])
    s0 = idx.stats()
    res = idx.search(qs, k, metric)
    assert idx.stats().tc_queries - s0.tc_queries == len(qs)
    for i in range(len(qs)):
        er, es = o.search(rows, qs[i], k, metric, threads=4)
        assert np.array_equal(res[i][0], er), (metric, i)
        assert np.array_equal(res[i][1].view(np.uint32), es.view(np.uint32))
    idx.close()


def test_two_device_mutations_keep_row_ids_dense():
    """Advisor finding (round 1): after swap-removes across the shard boundary an append must not
    leave a gap in the global row ids.  Load 4 rows on 2 devices, remove rows 3, 2 and 0, append:
    the new row is row 1 of 2, every id a search returns is < rows, and update/get_row find it."""
    _need(2)
    d = 8
    rows = o.fill_synthetic(6, d, 77)
    idx = DeviceIndex(d, devices=[0, 1])
    idx.load(rows[:4])
    assert idx.swap_remove(3) == 3
    assert idx.swap_remove(2) == 2
    assert idx.swap_remove(0) == 1          # row 1 moved into slot 0
    assert idx.rows == 1
    idx.append(rows[4:5])
    assert idx.rows == 2
    host = np.stack([rows[1], rows[4]])
    assert np.array_equal(idx.get_row(1), rows[4])
    for metric in ("cosine", "euclidean", "dot"):
        ((r, s),) = idx.search(rows[4], 5, metric)
        er, es = o.search(host, rows[4], 5, metric)
        assert np.array_equal(r, er) and np.array_equal(s.view(np.uint32), es.view(np.uint32))
        assert all(int(x) < idx.rows for x in r)
    idx.update(1, rows[5])
    assert np.array_equal(idx.get_row(1), rows[5])
    bases = [idx.shard_info(i).row_base for i in range(2)]
    sizes = [idx.shard_info(i).rows for i in range(2)]
    assert bases[0] == 0 and bases[1] == sizes[0] and sum(sizes) == 2
    idx.close()


def test_two_device_appends_are_rebalanced():
    """Appends land on the last shard; once it holds more than 1.5x its share the rows are re-split
    into equal contiguous ranges.  Global row ids (== the caller's key table) never change."""
    _need(2)
    d, n0, step = 48, 10_000, 9_000
    rows = o.fill_synthetic(n0 + 6 * step, d, 0x5EED0001)
    idx = DeviceIndex(d, devices=[0, 1])
    idx.load(rows[:n0])
    n = n0
    q = o.fill_synthetic(1, d, 3)[0]
    for i in range(6):
        idx.append(rows[n:n + step])
        n += step
        sizes = [idx.shard_info(s).rows for s in range(2)]
        assert sum(sizes) == n and max(sizes) <= 1.5 * ((n + 1) // 2) + 4096
        assert idx.shard_info(1).row_base == sizes[0]
        ((r, s),) = idx.search(q, 25, "cosine")
        er, es = o.search(rows[:n], q, 25, "cosine", threads=4)
        assert np.array_equal(r, er) and np.array_equal(s.view(np.uint32), es.view(np.uint32))
    sizes = [idx.shard_info(s).rows for s in range(2)]
    assert min(sizes) > n // 4, sizes                       # re-split happened at least once
    probe = [0, sizes[0] - 1, sizes[0], n - 1]
    for g in probe:
        assert np.array_equal(idx.get_row(g), rows[g])
    idx.close()


def test_two_device_filtered_and_masked_search():
    """Filtered (device-evaluated program) and masked (host bitmask) searches on an in-process
    two-device index: each shard evaluates / takes its own slice; results equal the oracle on the
    eligible subset, with global row ids."""
    _need(2)
    from neumann_b200._ffi import NM_C_LT, NM_F_CMP, NM_V_INT, NmFilterOp
    n, d, k = 150_003, 40, 9               # 75k rows per shard: batches take the tensor cores
    rows = o.fill_synthetic(n, d, 0x5EED0001)
    idx = DeviceIndex(d, devices=[0, 1])
    idx.load(rows)
    bucket = (np.arange(n) * 7919 % 101).astype(np.uint64)
    idx.column_set(3, 0, np.full(n, 3, np.uint8), bucket)            # NM_V_INT
    prog = [NmFilterOp(kind=NM_F_CMP, cmp=NM_C_LT, lit_tag=NM_V_INT, column=3, lit=4)]
    keep = bucket < 4
    assert np.array_equal(idx.filter_mask(prog), keep)
    sub = np.nonzero(keep)[0]
    qs = o.fill_synthetic(2, d, 3)
    for metric in ("cosine", "euclidean"):
        t0 = idx.stats().tc_queries
        res = idx.search_filtered(qs, k, metric, prog)              # a batch of 2: tensor-core
        resm = idx.search_masked(qs, k, metric, keep)               # pre-filter per shard, masked
        assert idx.stats().tc_queries - t0 == 4
        single = idx.search_filtered(qs[0], k, metric, prog)        # one query: masked f32 scans
        for i in range(2):
            er, es = o.search(rows[sub], qs[i], k, metric, threads=4)
            want = sub[er.astype(np.int64)].astype(np.uint64)
            for got in (res[i], resm[i]) + ((single[0],) if i == 0 else ()):
                assert np.array_equal(got[0], want), (metric, i)
                assert np.array_equal(got[1].view(np.uint32), es.view(np.uint32))
    # rows move between the shards when they are re-split: the columns follow
    idx.append(np.tile(rows[:1], (250_000, 1)))         # last shard > 1.5x its share -> re-split
    sizes = [idx.shard_info(s).rows for s in range(2)]
    assert abs(sizes[0] - sizes[1]) <= 1
    m = idx.filter_mask(prog)
    assert np.array_equal(m[:n], keep) and not m[n:].any()
    idx.close()


def _filtered_collective_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    from neumann_b200 import dist as nd
    from neumann_b200._ffi import NM_C_GE, NM_F_CMP, NM_V_INT, NmFilterOp
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n, d, k = 140_001, 64, 10               # 70k rows per rank: batches take the tensor cores
        idx = DeviceIndex(d, devices=[rank])
        lo, hi = nd.attach_index(idx, n)
        idx.fill_synthetic(hi - lo, 0x5EED0001, row_offset=lo)
        rows = o.fill_synthetic(n, d, 0x5EED0001)
        bucket = (np.arange(n) * 31 % 53).astype(np.uint64)
        idx.column_set(1, 0, np.full(hi - lo, 3, np.uint8), bucket[lo:hi])   # this rank's rows
        prog = [NmFilterOp(kind=NM_F_CMP, cmp=NM_C_GE, lit_tag=NM_V_INT, column=1, lit=50)]
        keep = bucket >= 50
        sub = np.nonzero(keep)[0]
        qs = o.fill_synthetic(3, d, 0x5EED1001)
        for metric in ("cosine", "dot"):
            t0 = idx.stats().tc_queries
            res = idx.search_filtered(qs, k, metric, prog)            # batch: per-shard tensor-core
            resm = idx.search_masked(qs, k, metric, keep[lo:hi])      # pre-filter + ONE all-gather
            assert idx.stats().tc_queries - t0 == 6
            one = idx.search_filtered(qs[1], k, metric, prog)         # single: fused exchange, masked
            for i in range(3):
                er, es = o.search(rows[sub], qs[i], k, metric, threads=4)
                want = sub[er.astype(np.int64)].astype(np.uint64)
                for got in (res[i], resm[i]) + ((one[0],) if i == 1 else ()):
                    assert np.array_equal(got[0], want), (metric, i, got[0], want)
                    assert np.array_equal(got[1].view(np.uint32), es.view(np.uint32))
        idx.detach_comm()
        idx.close()
        q.put((rank, "ok"))
    except Exception:  # noqa: BLE001
        import traceback
        q.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


def test_collective_filtered_and_masked_search():
    _need(2)
    _spawn(_filtered_collective_worker, 2)


def test_collective_filtered_search_nccl_fallback(monkeypatch):
    _need(2)
    monkeypatch.setenv("NM_DISABLE_PEER_EXCHANGE", "1")
    _spawn(_filtered_collective_worker, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _nccl_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    from neumann_b200 import dist as nd
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n, d, k = 200_003, 96, 16
        idx = DeviceIndex(d, devices=[rank])
        idx.set_prefilter(0)      # exact kernels first; the tensor-core pass is switched on below
        lo, hi = nd.attach_index(idx, n)
        idx.fill_synthetic(hi - lo, 0x5EED0001, row_offset=lo)
        qs = o.fill_synthetic(3, d, 0x5EED1001)
        rows = o.fill_synthetic(n, d, 0x5EED0001)
        for metric in ("cosine", "euclidean", "dot"):
            for i in range(3):                      # single-query calls: the fused path
                (res,) = idx.search(qs[i], k, metric)
                er, es = o.search(rows, qs[i], k, metric, threads=4)
                assert np.array_equal(res[0], er), (metric, i, res[0], er)
                assert np.array_equal(res[1].view(np.uint32), es.view(np.uint32))
        fused = os.environ.get("NM_DISABLE_PEER_EXCHANGE") != "1"
        # single queries: fused scan+exchange+merge launches (no separate merge kernel);
        # with the exchange disabled: one merge_shards_kernel per search call
        assert idx.stats().merge_launches == (0 if fused else 9)
        # a batch of 9 queries takes the batched kernels + ONE ncclAllGather + merge kernel
        qb = o.fill_synthetic(9, d, 0xBA7C)
        res = idx.search(qb, k, "euclidean")
        for i in range(9):
            er, es = o.search(rows, qb[i], k, "euclidean", threads=4)
            assert np.array_equal(res[i][0], er), ("batched", i)
            assert np.array_equal(res[i][1].view(np.uint32), es.view(np.uint32))
        assert idx.stats().merge_launches == (1 if fused else 10)
        # the same batch with the pre-filter on: shards of >= 65536 rows compute their hits with
        # the tensor-core pre-filter (a rank-local choice), same all-gather + merge
        idx.set_prefilter(1)
        t0 = idx.stats().tc_queries
        res = idx.search(qb, k, "euclidean")
        for i in range(9):
            er, es = o.search(rows, qb[i], k, "euclidean", threads=4)
            assert np.array_equal(res[i][0], er), ("tensor core", i)
            assert np.array_equal(res[i][1].view(np.uint32), es.view(np.uint32))
        assert (idx.stats().tc_queries - t0 == 9) == (hi - lo >= 65536)
        idx.set_prefilter(0)
        # k larger than a shard (and than the fast limit): chained passes + NCCL path
        res = idx.search(qs[:1], 1500, "cosine")
        er, es = o.search(rows, qs[0], 1500, "cosine", threads=4)
        assert np.array_equal(res[0][0], er) and np.array_equal(res[0][1].view(np.uint32), es.view(np.uint32))
        # gathers too large for the shared-memory sort (ranks x k > 4096): rank-based merge kernel,
        # two queries at once, with exact ties across the shard boundary
        res = idx.search(qs[:2], 5000, "dot")
        for i in range(2):
            er, es = o.search(rows, qs[i], 5000, "dot", threads=4)
            assert np.array_equal(res[i][0], er), ("rank merge", i)
            assert np.array_equal(res[i][1].view(np.uint32), es.view(np.uint32))
        idx.detach_comm()
        idx.close()
        q.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        import traceback
        q.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


def _empty_shard_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    from neumann_b200 import dist as nd
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # 3 rows over 2 ranks -> rank 0 holds 1 row, rank 1 holds 2; then rank 0 is emptied
        d = 16
        rows = o.fill_synthetic(2, d, 5, row_offset=0)
        idx = DeviceIndex(d, devices=[rank])
        nd.attach_index(idx, 2)      # bounds: rank0 [0,1) rank1 [1,2)
        if rank == 1:
            idx.load(rows[1:2])
        res = idx.search(rows[1], 5, "cosine")      # rank 0's shard is EMPTY
        assert list(res[0][0]) == [1] and abs(float(res[0][1][0]) - 1.0) < 1e-6
        idx.detach_comm()
        idx.close()
        q.put((rank, "ok"))
    except Exception:  # noqa: BLE001
        import traceback
        q.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


def _spawn(target, world):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=target, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=120)
    assert sorted(res) == [(r, "ok") for r in range(world)], res


def test_collective_with_an_empty_shard():
    _need(2)
    _spawn(_empty_shard_worker, 2)


def test_nccl_fallback_when_peer_exchange_disabled(monkeypatch):
    _need(2)
    monkeypatch.setenv("NM_DISABLE_PEER_EXCHANGE", "1")
    _spawn(_nccl_worker, 2)


@pytest.mark.parametrize("world", [2, 4, 8])
def test_nccl_allgather_shards(world):
    _need(world)
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_nccl_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=120)
    assert sorted(res) == [(r, "ok") for r in range(world)], res
