"""The reference's own benchmark shapes (vector_engine/benches/vector_engine_bench.rs:39-59,
docs/book/src/benchmarks/vector-engine.md:31-33): search_similar top-10 on 1k x 128, 1k x 768,
10k x 128 — through the VectorEngine host mirror (keys in, keys out) and through nm_search."""
import sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np
import oracle_ffi as o
from neumann_b200 import DeviceIndex
from neumann_b200.engine import VectorEngine

def med_us(fn, n=300, warm=30):
    for _ in range(warm): fn()
    ts = []
    for _ in range(n):
        t = time.perf_counter(); fn(); ts.append(time.perf_counter() - t)
    return float(np.median(ts) * 1e6), float(np.percentile(ts, 90) * 1e6)

published = {(1000, 128): 242, (1000, 768): 367, (10000, 128): 1930}
for (n, d) in [(1000, 128), (1000, 768), (10000, 128), (100000, 128), (1000000, 128)]:
    rng = np.random.default_rng(42)
    rows = rng.uniform(-1, 1, (n, d)).astype(np.float32)      # bench recipe: uniform(-1,1)
    q = rng.uniform(-1, 1, d).astype(np.float32)
    idx = DeviceIndex(d); idx.load(rows)
    abi = med_us(lambda: idx.search(q, 10, "cosine"))
    idx.set_profiling(True); idx.search(q, 10, "cosine"); k_ms = idx.stats().last_scan_ms * 1e3; idx.set_profiling(False)
    eng_t = None
    if n <= 100000:
        e = VectorEngine()
        for i in range(n): e.store_embedding(f"v{i}", rows[i])
        eng_t = med_us(lambda: e.search_similar(q, 10))
        r = e.search_similar(q, 10); er, es = o.search(rows, q, 10, "cosine")
        assert [x.key for x in r] == [f"v{int(i)}" for i in er]
    cpu1 = med_us(lambda: o.search(rows, q, 10, "cosine"), n=5, warm=1)
    cpumt = med_us(lambda: o.search(rows, q, 10, "cosine", threads=16), n=5, warm=1)
    print(f"{n}x{d}: nm_search {abi[0]:.1f} us (p90 {abi[1]:.1f}), kernel {k_ms:.1f} us, "
          f"VectorEngine.search_similar {eng_t[0] if eng_t else float('nan'):.1f} us | oracle 1 thread {cpu1[0]:.0f} us, "
          f"16 threads {cpumt[0]:.0f} us | reference published {published.get((n, d), '-')} us", flush=True)
    idx.close()
