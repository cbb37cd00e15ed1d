"""Serving-style throughput: T host threads, each issuing single-query nm_search calls against
the same 10M x 768 mirror, with and without coalescing."""
import sys, time, threading
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np
from neumann_b200 import DeviceIndex
from neumann_b200.synth import synth_rows

n, d, k = 10_000_000, 768, 10
idx = DeviceIndex(d); idx.fill_synthetic(n, 0x5EED0001)
qs = synth_rows(256, d, 0x5EED1001)
modes = [("f32", 64), ("f32", 1)]
if len(sys.argv) > 1 and sys.argv[1] == "prefilter":
    idx.set_prefilter(1)   # single calls: int8 dp4a pre-filter; coalesced rounds: tcgen05 pre-filter
    modes = [("int8", 64), ("int8", 1)]
for T in (1, 4, 16, 64):
    for tag, co in modes:
        idx.set_coalescing(co)
        per = max(8, 256 // T)
        lat = []
        def worker(t):
            for j in range(per):
                t0 = time.perf_counter()
                idx.search(qs[(t * per + j) % 256], k, "cosine")
                lat.append(time.perf_counter() - t0)
        for i in range(3): idx.search(qs[i], k, "cosine")
        ts = [threading.Thread(target=worker, args=(t,)) for t in range(T)]
        t0 = time.perf_counter()
        for t in ts: t.start()
        for t in ts: t.join()
        dt = time.perf_counter() - t0
        print(f"[{tag}] threads={T:3d} coalescing={'on ' if co > 1 else 'off'}: {T * per / dt:8.1f} QPS  "
              f"latency p50 {np.median(lat) * 1e3:7.2f} ms p99 {np.percentile(lat, 99) * 1e3:7.2f} ms", flush=True)
