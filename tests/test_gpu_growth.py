"""The mirror grows IN PLACE (CUDA virtual memory management, nm_vmm.hpp): appends map more
physical memory behind the existing rows instead of realloc-and-copy, so a mirror can grow to the
free memory of the device, and searches keep running while it grows (reference behaviour pin:
concurrent store + search, vector_engine/src/lib.rs:5615-5711)."""
import threading

import numpy as np
import pytest

import oracle_ffi as o
from neumann_b200 import DeviceIndex

pytestmark = pytest.mark.gpu


def test_appends_grow_in_place_and_match_the_oracle():
    d = 100
    rows = o.fill_synthetic(260_000, d, 0x5EED0001)
    idx = DeviceIndex(d)
    idx.load(rows[:1000])
    info = idx.shard_info()
    assert info.grows_in_place == 1 and info.rows == 1000 and info.capacity_rows >= 1000
    n, mapped, remaps = 1000, info.mapped_bytes, info.remaps
    q = o.fill_synthetic(1, d, 9)[0]
    for step in (1, 37, 5_000, 60_000, 193_962):
        idx.append(rows[n:n + step])
        n += step
        info = idx.shard_info()
        assert info.rows == n and info.capacity_rows >= n
        assert info.mapped_bytes >= mapped and info.mapped_bytes >= n * d * 4
        assert info.reserved_bytes >= info.mapped_bytes
        mapped = info.mapped_bytes
        for metric in ("cosine", "euclidean"):
            ((r, s),) = idx.search(q, 10, metric)
            er, es = o.search(rows[:n], q, 10, metric, threads=8)
            assert np.array_equal(r, er) and np.array_equal(s.view(np.uint32), es.view(np.uint32))
    assert np.array_equal(idx.get_rows(0, n).view(np.uint32), rows[:n].view(np.uint32))
    # geometric growth: far fewer physical chunks than appends would need one by one, and the
    # virtual range was replaced (re-mapped, not copied) only a few times
    assert info.chunks <= 40 and info.remaps - remaps <= 8
    idx.close()


def test_append_past_60_percent_of_device_memory_without_oom():
    """Realloc-and-copy growth needs old + new buffer at once and dies long before the device is
    full.  Append 1 GiB slabs until the mirror holds > 0.6 x the device's memory; the batch path's
    int8 copy is off so that the f32 mirror alone crosses the line."""
    import torch
    free_b, total_b = torch.cuda.mem_get_info()
    target = int(0.62 * total_b)
    if free_b < target + (8 << 30):
        pytest.skip(f"only {free_b >> 30} GiB free of {total_b >> 30} GiB")
    d = 1024
    slab_rows = (1 << 30) // (d * 4)
    slab = torch.empty((slab_rows, d), dtype=torch.float32).pin_memory().numpy()
    slab[:] = o.fill_synthetic(slab_rows, d, 0xA11CE)
    q = o.fill_synthetic(1, d, 0xB0B)[0]
    idx = DeviceIndex(d)
    idx.set_prefilter(0)
    n = 0
    while n * d * 4 <= target:
        idx.append(slab)
        n += slab_rows
    info = idx.shard_info()
    assert info.rows == n and info.mapped_bytes > 0.6 * total_b and info.grows_in_place == 1
    # every slab is the same block of rows: the best row of the slab wins in every copy, lowest id first
    er, es = o.search(slab, q, 1, "cosine", threads=16)
    ((r, s),) = idx.search(q, 3, "cosine")
    assert [int(x) for x in r] == [int(er[0]) + i * slab_rows for i in range(3)]
    assert all(x.view(np.uint32) == es[0].view(np.uint32) for x in s)
    # the last row really is where it should be
    assert np.array_equal(idx.get_row(n - 1), slab[-1])
    idx.close()


def test_search_during_growth():
    """One thread keeps appending, four keep searching (VE:5615-5711).  Every result must be the
    exact top-k of SOME prefix of the final corpus that is at least as long as the rows visible when
    the call started: appends are atomic with respect to searches, growth never tears a scan."""
    d, k = 64, 5
    total = 400_000
    rows = o.fill_synthetic(total, d, 0x5EED0001)
    q = o.fill_synthetic(1, d, 0x51)[0]
    scores = o.score_rows(rows, q, "dot")
    idx = DeviceIndex(d)
    idx.load(rows[:10_000])
    sizes = [10_000]
    stop = threading.Event()
    errors, seen = [], []

    def appender():
        try:
            n = 10_000
            rng = np.random.default_rng(5)
            while n < total:
                step = int(min(total - n, rng.integers(1, 30_000)))
                idx.append(rows[n:n + step])
                n += step
                sizes.append(n)
        except Exception as e:  # noqa: BLE001
            errors.append(repr(e))
        finally:
            stop.set()

    def searcher():
        try:
            while not stop.is_set():
                n_before = idx.rows
                ((r, s),) = idx.search(q, k, "dot")
                seen.append((n_before, r.copy(), s.copy()))
        except Exception as e:  # noqa: BLE001
            errors.append(repr(e))

    ts = [threading.Thread(target=appender)] + [threading.Thread(target=searcher) for _ in range(4)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errors, errors[:2]
    assert len(seen) > 20 and idx.rows == total
    # exact top-k of every committed prefix: score desc, ties by ascending row
    import np_ref
    key = np_ref.orderable(scores).astype(np.int64)
    order = np.lexsort((np.arange(total), -key))
    topk_of = {m: [int(x) for x in order[order < m][:k]] for m in set(sizes)}
    for n_before, r, s in seen:
        assert np.array_equal(s.view(np.uint32), scores[r.astype(np.int64)].view(np.uint32))
        got = [int(x) for x in r]
        assert any(got == topk_of[m] for m in topk_of if m >= n_before), (n_before, got)
    idx.close()


def test_realloc_fallback_without_vmm():
    """NM_NO_VMM=1 (or a driver without virtual memory management) keeps the r01 growth strategy —
    a larger cudaMalloc and one copy.  Same results; shard_info says which strategy is active."""
    import subprocess
    import sys
    from pathlib import Path
    root = Path(__file__).resolve().parent.parent
    code = r'''
import sys
sys.path.insert(0, "ROOT"); sys.path.insert(0, "ROOT/tests")
import numpy as np
import oracle_ffi as o
from neumann_b200 import DeviceIndex
d = 40
rows = o.fill_synthetic(90_000, d, 7)
idx = DeviceIndex(d)
idx.load(rows[:10])
n = 10
for step in (1, 500, 20_000, 69_489):
    idx.append(rows[n:n + step]); n += step
info = idx.shard_info()
assert info.grows_in_place == 0 and info.rows == n == 90_000, (info.grows_in_place, info.rows)
idx.column_set(2, 0, np.full(n, 3, np.uint8), np.arange(n, dtype=np.uint64))
q = o.fill_synthetic(1, d, 8)[0]
for metric in ("cosine", "euclidean", "dot"):
    ((r, s),) = idx.search(q, 10, metric)
    er, es = o.search(rows, q, 10, metric, threads=4)
    assert np.array_equal(r, er) and np.array_equal(s.view(np.uint32), es.view(np.uint32)), metric
res = idx.search(rows[:4], 5, "euclidean")          # a batch: int8 copy in plain buffers too
assert all(int(res[i][0][0]) == i for i in range(4))
print("FALLBACK_OK")
'''.replace("ROOT", str(root))
    env = dict(__import__("os").environ, NM_NO_VMM="1")
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "FALLBACK_OK" in r.stdout, (r.stdout[-400:], r.stderr[-800:])
