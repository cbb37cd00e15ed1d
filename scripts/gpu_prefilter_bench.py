"""int8 pre-filter vs exact f32 scan, end to end through nm_search (host buffers)."""
import sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np
from neumann_b200 import DeviceIndex
from neumann_b200.synth import synth_rows

for (n, d, k, metric) in [(10_000_000, 768, 10, "cosine"), (10_000_000, 768, 100, "dot"), (1_000_000, 768, 10, "cosine")]:
    idx = DeviceIndex(d); idx.fill_synthetic(n, 0x5EED0001)
    qs = synth_rows(16, d, 0x5EED1001)
    def run(steps=100):
        for i in range(10): idx.search(qs[i % 16], k, metric)
        t = time.perf_counter()
        for i in range(steps): r = idx.search(qs[i % 16], k, metric)
        return (time.perf_counter() - t) / steps, r
    t_exact, r_exact = run()
    t0 = time.perf_counter(); idx.set_prefilter(1); t_build = time.perf_counter() - t0
    s0 = idx.stats()
    t_pf, r_pf = run()
    s1 = idx.stats()
    same = np.array_equal(r_exact[0][0], r_pf[0][0]) and np.array_equal(r_exact[0][1].view(np.uint32), r_pf[0][1].view(np.uint32))
    nqs = s1.prefilter_queries - s0.prefilter_queries
    print(f"{n}x{d} {metric} k={k}: exact {t_exact*1e3:.3f} ms ({1/t_exact:.1f} QPS) | int8 pre-filter {t_pf*1e3:.3f} ms ({1/t_pf:.1f} QPS) "
          f"= {t_exact/t_pf:.2f}x | identical={same} | kept/query {(s1.prefilter_kept-s0.prefilter_kept)/max(nqs,1):.0f} "
          f"fallbacks {s1.prefilter_fallbacks-s0.prefilter_fallbacks}/{nqs} | quantise {t_build*1e3:.0f} ms", flush=True)
    idx.close()
