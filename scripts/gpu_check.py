"""First-contact GPU check: parity vs the oracle on small shapes + raw scan timing."""
import sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np
import oracle_ffi as o
from neumann_b200 import DeviceIndex

def parity(n, d, k, metric, seed=0x5EED0001, via_load=False):
    idx = DeviceIndex(d)
    rows = o.fill_synthetic(n, d, seed)
    if via_load:
        idx.load(rows)
    else:
        idx.fill_synthetic(n, seed)
    got = idx.get_row(n - 1)
    assert np.array_equal(got.view(np.uint32), rows[n - 1].view(np.uint32)), "fill mismatch"
    bad = 0
    for qi in range(3):
        q = o.fill_synthetic(1, d, 0x5EED1001 + qi)[0]
        er, es = o.search(rows, q, k, metric)
        (gr, gs), = idx.search(q, k, metric)
        ok = np.array_equal(er, gr) and np.array_equal(es.view(np.uint32), gs.view(np.uint32))
        if not ok:
            bad += 1
            print("  MISMATCH", metric, n, d, k, "exp", er[:5], es[:5], "got", gr[:5], gs[:5])
    print(f"parity n={n} d={d} k={k} {metric} load={via_load}: {'OK' if not bad else 'FAIL'}", flush=True)
    idx.close()
    return bad == 0

ok = True
for (n, d, k) in [(1000, 128, 5), (10000, 128, 5), (5000, 768, 10), (3, 3, 3), (300, 100, 7),
                  (257, 33, 10), (70000, 64, 100), (4096, 1536, 100), (50000, 96, 1000), (1000, 7, 4)]:
    for m in ("cosine", "euclidean", "dot"):
        ok &= parity(n, d, k, m)
ok &= parity(20000, 768, 10, "cosine", via_load=True)
ok &= parity(1000, 13, 10, "euclidean", via_load=True)
print("ALL PARITY", "OK" if ok else "FAIL", flush=True)

for (n, d, k, metric) in [(1_000_000, 768, 10, "cosine"), (10_000_000, 768, 10, "cosine"),
                          (10_000_000, 768, 10, "euclidean"), (10_000_000, 768, 10, "dot"),
                          (5_000_000, 1536, 100, "euclidean")]:
    idx = DeviceIndex(d); idx.set_profiling(True)   # last_scan_ms
    t0 = time.time(); idx.fill_synthetic(n, 0x5EED0001); t1 = time.time()
    q = o.fill_synthetic(1, d, 0x5EED1001)[0]
    for _ in range(3): idx.search(q, k, metric)
    ts = []; ks = []
    for _ in range(10):
        t = time.perf_counter(); r = idx.search(q, k, metric); ts.append(time.perf_counter() - t)
        ks.append(idx.stats().last_scan_ms)
    gb = n * d * 4 / 1e9
    print(f"n={n} d={d} k={k} {metric}: fill {t1-t0:.2f}s  e2e med {np.median(ts)*1e3:.3f} ms  "
          f"kernel med {np.median(ks):.3f} ms min {np.min(ks):.3f}  -> {gb/np.median(ks)*1e3:.0f} GB/s "
          f"({gb/np.median(ks)*1e3/6549.1*100:.1f}% of 6549)  top={r[0][0][:3]} {r[0][1][:3]}", flush=True)
    idx.close()
