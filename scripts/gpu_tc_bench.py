"""Timing of the tensor-core batch pre-filter: python gpu_tc_bench.py n dim nq k metric [steps]"""
import sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np
from neumann_b200 import DeviceIndex
from neumann_b200.synth import synth_rows

n, d, nq, k = (int(x) for x in sys.argv[1:5])
metric = sys.argv[5]
steps = int(sys.argv[6]) if len(sys.argv) > 6 else 5
idx = DeviceIndex(d); idx.fill_synthetic(n, 0x5EED0001); idx.set_prefilter(1); idx.set_profiling(True)
qs = synth_rows(nq, d, 0x5EED1001)
idx.search(qs, k, metric)
s0 = idx.stats()
ts = []
for i in range(steps):
    t0 = time.perf_counter(); r = idx.search(qs, k, metric); ts.append(time.perf_counter() - t0)
s1 = idx.stats()
nqs = max(s1.tc_queries - s0.tc_queries, 1)
print(f"{n}x{d} {metric} k={k} nq={nq}: e2e best {min(ts)*1e3:.3f} ms median {sorted(ts)[len(ts)//2]*1e3:.3f} ms "
      f"({nq/min(ts):.0f} QPS) device {s1.last_scan_ms:.3f} ms fallbacks {s1.tc_fallbacks-s0.tc_fallbacks} "
      f"survivors/query {(s1.tc_survivors-s0.tc_survivors)/nqs:.0f} launches/call {(s1.scan_launches-s0.scan_launches)/steps:.0f}")
