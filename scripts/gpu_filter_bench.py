"""Filtered SIMILAR timing (device-evaluated filter): python gpu_filter_bench.py [rows] [dim]"""
import sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np
from neumann_b200 import DeviceIndex
from neumann_b200._ffi import NM_C_LT, NM_F_CMP, NM_V_INT, NmFilterOp
from neumann_b200.synth import synth_rows

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
d = int(sys.argv[2]) if len(sys.argv) > 2 else 768
idx = DeviceIndex(d); idx.fill_synthetic(n, 0x5EED0001)
qs = synth_rows(16, d, 0x5EED1001)
idx.set_prefilter(1); idx.search(qs[0], 10, "cosine"); idx.set_prefilter(0)   # build and drop an int8 copy first
bucket = (np.arange(n, dtype=np.uint64) * np.uint64(2654435761)) % np.uint64(100)
t0 = time.perf_counter(); idx.column_set(1, 0, np.full(n, NM_V_INT, np.uint8), bucket); t_col = time.perf_counter() - t0
for i in range(5): idx.search(qs[i], 10, "cosine")
t0 = time.perf_counter()
for i in range(20): idx.search(qs[i % 16], 10, "cosine")
t_plain = (time.perf_counter() - t0) / 20
for lim in (50, 1):
    first = []
    for rep in range(6):
        prog = [NmFilterOp(kind=NM_F_CMP, cmp=NM_C_LT, lit_tag=NM_V_INT, column=1, lit=lim + 100 * rep)]  # a NEW filter each time
        prog = [NmFilterOp(kind=NM_F_CMP, cmp=NM_C_LT, lit_tag=NM_V_INT, column=1, lit=lim)] if rep == 0 else prog
        t0 = time.perf_counter(); idx.search_filtered(qs[rep], 10, "cosine", prog); first.append(time.perf_counter() - t0)
    prog = [NmFilterOp(kind=NM_F_CMP, cmp=NM_C_LT, lit_tag=NM_V_INT, column=1, lit=lim)]
    t0 = time.perf_counter()
    for i in range(20): idx.search_filtered(qs[i % 16], 10, "cosine", prog)
    t_cached = (time.perf_counter() - t0) / 20
    print(f"lim {lim}: unfiltered {t_plain*1e3:.3f} ms, new filter calls {[round(x*1e3, 3) for x in first]} ms, cached {t_cached*1e3:.3f} ms; "
          f"column upload {t_col*1e3:.1f} ms; masks built {idx.stats().filter_masks_built}")
