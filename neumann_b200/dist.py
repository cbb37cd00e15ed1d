"""Host-side plumbing for row-range sharding across processes (one process per GPU).

torch.distributed is used for rendezvous only: rank 0 creates the NCCL id through the C ABI
(nm_comm_create_id) and broadcasts it; the data-path exchange is the single ncclAllGather issued
inside nm_search (see include/neumann_b200.h).  Mirrors the reference's scatter/gather shape:
QueryPlanner -> ScatterGather{all shards, TopK(k)} (query_router/src/distributed.rs:173-179,
269-272) with contiguous ascending row ranges per shard.
"""
from __future__ import annotations

import os


def shard_bounds(n_rows: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous row range [lo, hi) of `rank`; same split as nm_index_load uses per device."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank {rank} of {world}")
    return n_rows * rank // world, n_rows * (rank + 1) // world


def env_rank_world() -> tuple[int, int, int]:
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def broadcast_bytes(payload: bytes | None, nbytes: int, src: int = 0, device=None) -> bytes:
    """Broadcast a fixed-size byte string from `src` over the default process group."""
    import torch
    import torch.distributed as dist
    if dist.get_rank() == src:
        assert payload is not None and len(payload) == nbytes
        t = torch.tensor(list(payload), dtype=torch.uint8)
    else:
        t = torch.zeros(nbytes, dtype=torch.uint8)
    if device is not None:
        t = t.to(device)
    dist.broadcast(t, src=src)
    return bytes(t.cpu().tolist())


def attach_index(index, n_rows_total: int) -> tuple[int, int]:
    """Create + broadcast the communicator id and attach `index` as this rank's shard.
    Returns this rank's [lo, hi).  No-op for world size 1."""
    import torch
    import torch.distributed as dist
    from .index import comm_create_id
    from ._ffi import NM_COMM_ID_BYTES
    rank, world = dist.get_rank(), dist.get_world_size()
    lo, hi = shard_bounds(n_rows_total, world, rank)
    if world == 1:
        return lo, hi
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else None
    cid = broadcast_bytes(comm_create_id() if rank == 0 else None, NM_COMM_ID_BYTES, 0, dev)
    index.attach_comm(cid, world, rank, lo)
    return lo, hi


def max_over_ranks(value: float, device=None) -> float:
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
