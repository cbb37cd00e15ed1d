"""Config 4 (10M x 1536 L2 top-100, 256-query batches) with T concurrent callers: calls of different
host threads run on their own workspaces / streams, so the latency-bound phases of one call (the
refines, the final exact re-score) overlap the tensor-core phases of another.
    python gpu_batch_overlap.py [rows] [dim] [nq] [k]"""
import sys, time, threading
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np
from neumann_b200 import DeviceIndex
from neumann_b200.synth import synth_rows

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
d = int(sys.argv[2]) if len(sys.argv) > 2 else 1536
nq = int(sys.argv[3]) if len(sys.argv) > 3 else 256
k = int(sys.argv[4]) if len(sys.argv) > 4 else 100
metric = sys.argv[5] if len(sys.argv) > 5 else "euclidean"
idx = DeviceIndex(d); idx.fill_synthetic(n, 0x5EED0004)
batches = [synth_rows(nq, d, 0x5EED1004 + 977 * b) for b in range(4)]
ref = [idx.search(b, k, metric) for b in batches]          # builds the int8 copy (auto mode), warms up
for b in batches: idx.search(b, k, metric)
for T in (1, 2, 3, 4):
    per = 12
    ok = [True]
    def worker(t):
        for j in range(per):
            b = (t + j) % 4
            res = idx.search(batches[b], k, metric)
            if j == per - 1:
                ok[0] &= all(np.array_equal(res[i][0], ref[b][i][0]) and
                             np.array_equal(res[i][1].view(np.uint32), ref[b][i][1].view(np.uint32)) for i in range(nq))
    ts = [threading.Thread(target=worker, args=(t,)) for t in range(T)]
    t0 = time.perf_counter()
    for t in ts: t.start()
    for t in ts: t.join()
    dt = time.perf_counter() - t0
    print(f"callers={T}: {T * per * nq / dt:9.0f} QPS, {dt / (T * per) * 1e3:6.3f} ms per {nq}-query batch "
          f"(per-caller latency {dt / per * 1e3:6.3f} ms), identical {ok[0]}", flush=True)
st = idx.stats()
print("tc batches", getattr(st, "tc_batches", None), "fallbacks", getattr(st, "tc_fallback_queries", None))
