"""world_size-2 gloo tests (CPU) of the N>1 host logic: shard bounds, id broadcast, and that
per-shard top-k + ResultMerger::merge_top_k semantics reproduce the global answer."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle_ffi as o
from neumann_b200 import dist as nd


def test_shard_bounds_cover_rows_contiguously():
    for n in (0, 1, 7, 1000, 10_000_001):
        for w in (1, 2, 3, 8):
            b = [nd.shard_bounds(n, w, r) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            assert max(hi - lo for lo, hi in b) - min(hi - lo for lo, hi in b) <= 1
    with pytest.raises(ValueError):
        nd.shard_bounds(10, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n, d, k = 5001, 48, 10
        rows = o.fill_synthetic(n, d, 0x5EED0001)
        rows[10] = rows[4000]  # exact tie across the two shards
        query = rows[4000]
        lo, hi = nd.shard_bounds(n, world, rank)
        token = nd.broadcast_bytes(bytes(range(128)) if rank == 0 else None, 128)
        assert token == bytes(range(128))
        r, s = o.search(rows[lo:hi], query, k, "cosine")
        gathered = [None] * world
        dist.all_gather_object(gathered, ((r + lo).tolist(), s.tolist()))
        mr, ms = o.merge_top_k([np.asarray(g[0], np.uint64) for g in gathered],
                               [np.asarray(g[1], np.float32) for g in gathered], k)
        er, es = o.search(rows, query, k, "cosine")
        assert np.array_equal(mr, er) and np.array_equal(ms.view(np.uint32), es.view(np.uint32))
        assert list(er[:2]) == [10, 4000]
        slow = nd.max_over_ranks(1.0 + rank)
        assert slow == float(world)
        q.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_shard_and_merge():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res


# ---- merge_shards_rank_kernel (scan_kernels.cuh) restated in numpy -----------------------------
def _rank_merge(lists, k):
    """What merge_shards_rank_kernel computes: every shard list is sorted (score desc, row asc) with
    empty slots last; a hit's final rank is its position in its own list plus, per other list, the
    number of hits that precede it — strictly better score, or equal score in an EARLIER shard
    (binary search).  Hits with rank < k land at out[rank]."""
    import np_ref
    ords = [np_ref.orderable(np.asarray(s, np.float32)).astype(np.int64) for _, s in lists]
    out_r = np.zeros(k, np.uint64)
    out_s = np.zeros(k, np.float32)
    written = np.zeros(k, bool)
    for s, (rows, scores) in enumerate(lists):
        for i in range(len(rows)):
            o_i = ords[s][i]
            rank = i
            for t in range(len(lists)):
                if t == s:
                    continue
                lo, hi = 0, len(ords[t])                 # first index that does NOT precede the hit
                while lo < hi:
                    mid = (lo + hi) // 2
                    before = ords[t][mid] > o_i or (ords[t][mid] == o_i and t < s)
                    lo, hi = (mid + 1, hi) if before else (lo, mid)
                rank += lo
            if rank < k:
                assert not written[rank]                  # ranks are a permutation: no collisions
                out_r[rank], out_s[rank], written[rank] = rows[i], scores[i], True
    n = int(min(k, sum(len(r) for r, _ in lists)))
    assert written[:n].all() and not written[n:].any()
    return out_r[:n], out_s[:n]


def test_rank_based_merge_equals_merge_top_k():
    """The large-gather merge kernel's algorithm against the oracle's merge_top_k
    (distributed.rs:413-433: concatenate in shard order, stable sort, truncate) on tie-heavy shard
    lists of ragged lengths, with NaN scores (rank last) and -0.0 / +0.0 (tie)."""
    rng = np.random.default_rng(7)
    for trial in range(60):
        n_shards = int(rng.integers(1, 9))
        k = int(rng.integers(1, 40))
        pool = np.array([0.5, 0.25, -0.0, 0.0, 1.0, -1.0, np.nan, 0.125, 3.0], np.float32)
        lists, base = [], 0
        for s in range(n_shards):
            m = int(rng.integers(0, k + 1))              # a shard returns at most k hits
            scores = rng.choice(pool, m) if trial % 2 else rng.normal(0, 1, m).astype(np.float32)
            rows = base + np.sort(rng.choice(1000, m, replace=False)).astype(np.uint64)
            # each shard list arrives sorted by (score desc, NaN last, row asc): the scan's order
            import np_ref
            order = np.lexsort((rows, -np_ref.orderable(scores).astype(np.int64)))
            lists.append((rows[order], np.asarray(scores, np.float32)[order]))
            base += 1000
        er, es = o.merge_top_k([r for r, _ in lists], [s for _, s in lists], k) if sum(len(r) for r, _ in lists) \
            else (np.zeros(0, np.uint64), np.zeros(0, np.float32))
        gr, gs = _rank_merge(lists, k)
        assert np.array_equal(gr, er), (trial, gr, er)
        nan = np.isnan(es)
        assert np.array_equal(np.isnan(gs), nan)
        assert np.array_equal(gs.view(np.uint32)[~nan], es.view(np.uint32)[~nan])
