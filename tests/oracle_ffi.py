"""ctypes binding of oracle/libnm_oracle.so — the CPU checker (test infrastructure only)."""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
_LIB = ROOT / "oracle" / "libnm_oracle.so"
COSINE, EUCLIDEAN, DOT = 0, 1, 2
METRICS = {"cosine": COSINE, "euclidean": EUCLIDEAN, "dot": DOT}

_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not _LIB.exists():
            subprocess.run(["make", "-C", str(ROOT / "oracle")], check=True, capture_output=True)
        l = C.CDLL(str(_LIB))
        vp, u64, u32, f32, i32 = C.c_void_p, C.c_uint64, C.c_uint32, C.c_float, C.c_int
        sig = {
            "nmo_dot_product": (f32, [vp, vp, u64]),
            "nmo_sum_of_squares": (f32, [vp, u64]),
            "nmo_magnitude": (f32, [vp, u64]),
            "nmo_euclidean_distance": (f32, [vp, vp, u64]),
            "nmo_cosine_similarity": (f32, [vp, vp, u64, f32]),
            "nmo_compute_score": (f32, [vp, vp, u64, f32, i32]),
            "nmo_compute_similarity": (f32, [vp, vp, u64]),
            "nmo_score_rows": (None, [vp, u64, u32, vp, i32, vp]),
            "nmo_search": (u64, [vp, u64, u32, vp, u64, i32, vp, vp]),
            "nmo_search_mt": (u64, [vp, u64, u32, vp, u64, i32, i32, vp, vp]),
            "nmo_merge_top_k": (u64, [vp, vp, vp, u64, u64, vp, vp]),
            "nmo_synth_value": (f32, [u64, u64]),
            "nmo_fill_synthetic": (None, [vp, u64, u32, u64, u64]),
            "nmo_fill_synthetic_mt": (None, [vp, u64, u32, u64, u64, i32]),
        }
        for n, (r, a) in sig.items():
            fn = getattr(l, n)
            fn.restype, fn.argtypes = r, a
        _lib = l
    return _lib


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


def dot_product(a, b) -> np.float32:
    a, b = _f32(a), _f32(b)
    return np.float32(lib().nmo_dot_product(a.ctypes.data, b.ctypes.data, a.size))


def magnitude(a) -> np.float32:
    a = _f32(a)
    return np.float32(lib().nmo_magnitude(a.ctypes.data, a.size))


def euclidean_distance(a, b) -> np.float32:
    a, b = _f32(a), _f32(b)
    return np.float32(lib().nmo_euclidean_distance(a.ctypes.data, b.ctypes.data, a.size))


def compute_similarity(a, b) -> np.float32:
    a, b = _f32(a), _f32(b)
    return np.float32(lib().nmo_compute_similarity(a.ctypes.data, b.ctypes.data, a.size))


def compute_score(q, x, metric) -> np.float32:
    q, x = _f32(q), _f32(x)
    m = METRICS[metric] if isinstance(metric, str) else metric
    return np.float32(lib().nmo_compute_score(q.ctypes.data, x.ctypes.data, q.size,
                                              magnitude(q), m))


def score_rows(rows, q, metric) -> np.ndarray:
    rows, q = _f32(rows), _f32(q)
    m = METRICS[metric] if isinstance(metric, str) else metric
    out = np.empty(rows.shape[0], np.float32)
    lib().nmo_score_rows(rows.ctypes.data, rows.shape[0], rows.shape[1], q.ctypes.data, m,
                         out.ctypes.data)
    return out


def search(rows, q, k, metric, threads: int = 0):
    rows, q = _f32(rows), _f32(q)
    m = METRICS[metric] if isinstance(metric, str) else metric
    n, d = rows.shape
    kk = max(min(k, n), 1)
    out_r = np.zeros(kk, np.uint64)
    out_s = np.zeros(kk, np.float32)
    if threads and threads > 0:
        c = lib().nmo_search_mt(rows.ctypes.data, n, d, q.ctypes.data, k, m, threads,
                                out_r.ctypes.data, out_s.ctypes.data)
    else:
        c = lib().nmo_search(rows.ctypes.data, n, d, q.ctypes.data, k, m, out_r.ctypes.data,
                             out_s.ctypes.data)
    return out_r[:c].copy(), out_s[:c].copy()


def merge_top_k(shard_rows: list, shard_scores: list, k: int):
    rows = np.concatenate([np.asarray(r, np.uint64) for r in shard_rows]) if shard_rows else \
        np.zeros(0, np.uint64)
    scores = np.concatenate([np.asarray(s, np.float32) for s in shard_scores]) if shard_scores \
        else np.zeros(0, np.float32)
    counts = np.array([len(r) for r in shard_rows], np.uint64)
    kk = max(min(k, rows.size), 1)
    out_r = np.zeros(kk, np.uint64)
    out_s = np.zeros(kk, np.float32)
    c = lib().nmo_merge_top_k(rows.ctypes.data, scores.ctypes.data, counts.ctypes.data,
                              len(shard_rows), k, out_r.ctypes.data, out_s.ctypes.data)
    return out_r[:c].copy(), out_s[:c].copy()


def fill_synthetic(n: int, dim: int, seed: int, row_offset: int = 0, threads: int = 8) -> np.ndarray:
    out = np.empty((n, dim), np.float32)
    lib().nmo_fill_synthetic_mt(out.ctypes.data, n, dim, seed, row_offset, threads)
    return out
