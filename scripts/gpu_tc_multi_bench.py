"""Sharded batch through the tensor-core pre-filter (one process per GPU, torchrun):
   torchrun --nproc-per-node N scripts/gpu_tc_multi_bench.py rows dim nq k metric
Every rank holds rows/N rows, computes its shard's hits with the tcgen05 pre-filter, ONE
ncclAllGather of the hits, merge kernel; rank 0 prints ms per batch (max over ranks)."""
import sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np
import torch
import torch.distributed as dist
from neumann_b200 import DeviceIndex
from neumann_b200 import dist as nd
from neumann_b200.synth import synth_rows

n, d, nq, k = (int(x) for x in sys.argv[1:5])
metric = sys.argv[5]
rank, world, local = nd.env_rank_world()
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
idx = DeviceIndex(d, devices=[local])
lo, hi = nd.attach_index(idx, n)
idx.fill_synthetic(hi - lo, 0x5EED0001, row_offset=lo)
qs = synth_rows(nq, d, 0x5EED1001)
exact = idx.search(qs[:4], k, metric)
idx.set_prefilter(1)
idx.search(qs, k, metric)
ts = []
for _ in range(8):
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter(); res = idx.search(qs, k, metric); ts.append(time.perf_counter() - t0)
t = nd.max_over_ranks(sorted(ts)[len(ts) // 2], dev)
same = all(np.array_equal(a[0], b[0]) and np.array_equal(a[1].view(np.uint32), b[1].view(np.uint32))
           for a, b in zip(res[:4], exact))
st = idx.stats()
if rank == 0:
    print(f"{world} GPUs, {n}x{d} {metric} k={k} nq={nq}: {t*1e3:.3f} ms per batch ({nq/t:.0f} QPS), "
          f"identical to the exact sharded path: {same}, tc_queries {st.tc_queries}, fallbacks {st.tc_fallbacks}", flush=True)
idx.detach_comm(); idx.close()
dist.destroy_process_group()
