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
