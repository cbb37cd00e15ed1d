"""Small pass over the tensor-core batch pre-filter kernels, meant to run under compute-sanitizer."""
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
for (n, d, nq, k) in [(66_000, 64, 5, 7), (70_001, 131, 18, 3)]:
    rows = o.fill_synthetic(n, d, 1)
    idx = DeviceIndex(d)
    idx.load(rows)
    idx.set_prefilter(1)
    qs = o.fill_synthetic(nq, d, 2)
    for m in ("euclidean", "cosine", "dot"):
        s0 = idx.stats().tc_queries
        res = idx.search(qs, k, m)
        ok &= idx.stats().tc_queries - s0 == nq
        for i in range(nq):
            ok &= same(res[i], o.search(rows, qs[i], k, m, threads=8))
    idx.close()
print("SANITIZE_TC_DRIVER", "OK" if ok else "MISMATCH")
sys.exit(0 if ok else 1)
