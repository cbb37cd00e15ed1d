"""Tiny end-to-end pass over every kernel, meant to run under compute-sanitizer."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np
import oracle_ffi as o
from neumann_b200 import DeviceIndex

def same(got, exp):
    return np.array_equal(got[0], exp[0]) and np.array_equal(got[1].view(np.uint32), exp[1].view(np.uint32))

ok = True
for (n, d) in [(3000, 100), (1500, 64), (700, 13)]:
    rows = o.fill_synthetic(n, d, 1)
    idx = DeviceIndex(d)
    idx.load(rows)
    qs = o.fill_synthetic(9, d, 2)
    for m in ("cosine", "euclidean", "dot"):
        for k in (10, 1500):
            (g,) = idx.search(qs[0], k, m)
            ok &= same(g, o.search(rows, qs[0], k, m))
        res = idx.search(qs, 7, m)                       # batched kernels (nq = 9)
        for i in range(9):
            ok &= same(res[i], o.search(rows, qs[i], 7, m))
        mask = (np.arange(n) % 3) == 0
        (g,) = idx.search_masked(qs[1], 5, m, mask)
        sub = np.nonzero(mask)[0]
        er, es = o.search(rows[sub], qs[1], 5, m)
        ok &= np.array_equal(g[0], sub[er.astype(np.int64)].astype(np.uint64))
    idx.set_prefilter(1)                                   # int8 pre-filter kernels
    for m in ("cosine", "dot"):
        (g,) = idx.search(qs[4], 10, m)
        ok &= same(g, o.search(rows, qs[4], 10, m))
    idx.set_prefilter(0)
    idx.update(5, qs[3]); idx.swap_remove(7); idx.append(rows[:10])
    (g,) = idx.search(qs[3], 3, "cosine")
    ok &= g[0][0] == 5
    idx.close()
print("SANITIZE_DRIVER", "OK" if ok else "MISMATCH")
sys.exit(0 if ok else 1)
