"""Bring-up check of the tensor-core batch pre-filter: integer dots vs numpy, then results vs
the exact batched kernels, then timing."""
import sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np
from neumann_b200 import DeviceIndex
from neumann_b200.synth import synth_rows


def quantise(q):
    q = q.astype(np.float32)
    s = np.float32(np.abs(q).max()) / np.float32(127.0)
    return np.clip(np.rint(q / s), -127, 127).astype(np.int32)


def check_dots(n, d, nq):
    idx = DeviceIndex(d); idx.fill_synthetic(n, 0x5EED0001); idx.set_prefilter(1); idx.set_profiling(True)
    qs = synth_rows(nq, d, 0x5EED1001)
    dots = idx.debug_tc_dots(qs)
    rows = sorted(set([0, 1, 31, 32, 127, 128, 129, n - 1, n - 2, n // 2] +
                      list(np.random.default_rng(1).integers(0, n, 200))))
    q8 = np.stack([quantise(q) for q in qs])
    bad = 0
    for r in rows:
        x8, _ = idx.debug_q8_row(int(r))
        ref = q8 @ x8.astype(np.int32)
        if not np.array_equal(ref, dots[:, r]):
            if bad < 5:
                print(f"  MISMATCH row {r}: ref {ref[:4]} got {dots[:4, r]}")
            bad += 1
    print(f"dots {n}x{d} nq={nq}: {len(rows) - bad}/{len(rows)} sampled rows exact", flush=True)
    idx.close()
    return bad == 0


def check_search(n, d, nq, k, metric, timing=False):
    idx = DeviceIndex(d); idx.fill_synthetic(n, 0x5EED0001)
    qs = synth_rows(nq, d, 0x5EED1001)
    t0 = time.perf_counter(); exact = idx.search(qs, k, metric); t_exact = time.perf_counter() - t0
    if timing:
        t0 = time.perf_counter(); exact = idx.search(qs, k, metric); t_exact = time.perf_counter() - t0
    idx.set_prefilter(1)
    s0 = idx.stats()
    tc = idx.search(qs, k, metric)
    t0 = time.perf_counter(); tc = idx.search(qs, k, metric); t_tc = time.perf_counter() - t0
    s1 = idx.stats()
    same = all(np.array_equal(a[0], b[0]) and np.array_equal(a[1].view(np.uint32), b[1].view(np.uint32))
               for a, b in zip(exact, tc))
    nqs = s1.tc_queries - s0.tc_queries
    print(f"search {n}x{d} {metric} k={k} nq={nq}: identical={same} exact {t_exact*1e3:.2f} ms tc {t_tc*1e3:.2f} ms "
          f"({nq/t_tc:.0f} QPS, {t_exact/t_tc:.1f}x) tc_queries={nqs} fallbacks={s1.tc_fallbacks-s0.tc_fallbacks} "
          f"survivors/query={(s1.tc_survivors-s0.tc_survivors)/max(nqs,1):.0f} last_scan_ms={s1.last_scan_ms:.3f}", flush=True)
    if not same:
        for i, (a, b) in enumerate(zip(exact, tc)):
            if not (np.array_equal(a[0], b[0]) and np.array_equal(a[1].view(np.uint32), b[1].view(np.uint32))):
                print("  first differing query", i, a[0][:5], b[0][:5], a[1][:5], b[1][:5]); break
    idx.close()
    return same


stage = sys.argv[1] if len(sys.argv) > 1 else "all"
ok = True
if stage in ("all", "dots"):
    ok &= check_dots(40_000, 128, 16)
    ok &= check_dots(40_000, 200, 20)
    ok &= check_dots(70_000, 768, 256)
if ok and stage in ("all", "search"):
    for metric in ("euclidean", "cosine", "dot"):
        ok &= check_search(100_000, 256, 16, 10, metric)
        ok &= check_search(300_000, 200, 37, 100, metric)
        ok &= check_search(70_000, 131, 5, 7, metric)
    ok &= check_search(1_000_000, 768, 256, 10, "cosine", timing=True)
if ok and stage in ("all", "big"):
    ok &= check_search(2_000_000, 1536, 256, 100, "euclidean", timing=True)
    ok &= check_search(10_000_000, 768, 256, 10, "cosine", timing=True)
    ok &= check_search(10_000_000, 1536, 256, 100, "euclidean", timing=True)
print("TC CHECK", "OK" if ok else "FAILED")
sys.exit(0 if ok else 1)
