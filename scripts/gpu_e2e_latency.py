"""End-to-end latency of nm_search with host buffers (what bench.py's e2e measures), per corpus size:
    python gpu_e2e_latency.py            (NM_ZERO_COPY_RESULTS=0 for the copied result block)"""
import os, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np
from neumann_b200 import DeviceIndex
from neumann_b200.synth import synth_rows

for n, d, k, reps in ((10_000, 128, 5, 3000), (1_000_000, 768, 10, 500), (10_000_000, 768, 10, 100)):
    idx = DeviceIndex(d); idx.fill_synthetic(n, 0x5EED0001)
    qs = synth_rows(16, d, 0x5EED1001)
    for i in range(20): idx.search(qs[i % 16], k, "cosine")
    ts = []
    for i in range(reps):
        t0 = time.perf_counter(); idx.search(qs[i % 16], k, "cosine"); ts.append(time.perf_counter() - t0)
    ts = np.array(ts) * 1e6
    print(f"zero_copy={os.environ.get('NM_ZERO_COPY_RESULTS', '1')} {n}x{d} top-{k}: mean {ts.mean():8.2f} us  p50 {np.median(ts):8.2f}  p10 {np.percentile(ts, 10):8.2f}  p90 {np.percentile(ts, 90):8.2f}", flush=True)
    idx.close()
