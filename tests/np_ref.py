"""Second, independent restatement of the reference arithmetic in numpy float32 (no C, no
oracle code shared): used to cross-check oracle/nm_oracle.c, never the product."""
from __future__ import annotations

import numpy as np

f32 = np.float32


def dot_product(a, b) -> np.float32:
    a = np.asarray(a, f32)
    b = np.asarray(b, f32)
    n = a.size
    chunks = n // 8
    lanes = np.zeros(8, f32)
    for c in range(chunks):  # f32x8: sum += va * vb  (hnsw.rs:172-178)
        lanes = (lanes + (a[c * 8:c * 8 + 8] * b[c * 8:c * 8 + 8]).astype(f32)).astype(f32)
    r = f32(0.0)
    for j in range(8):  # arr.iter().sum()
        r = f32(r + lanes[j])
    for i in range(chunks * 8, n):
        r = f32(r + f32(a[i] * b[i]))
    return r


def magnitude(v) -> np.float32:
    return f32(np.sqrt(dot_product(v, v)))


def euclidean_distance(a, b) -> np.float32:
    a = np.asarray(a, f32)
    b = np.asarray(b, f32)
    s = f32(0.0)
    for i in range(a.size):
        d = f32(a[i] - b[i])
        s = f32(s + f32(d * d))
    return f32(np.sqrt(s))


def score(q, x, metric: str) -> np.float32:
    if metric == "cosine":
        qm, xm = magnitude(q), magnitude(x)
        if qm == 0 or xm == 0:
            return f32(0.0)
        return f32(dot_product(q, x) / f32(qm * xm))
    if metric == "dot":
        return dot_product(q, x)
    return f32(f32(1.0) / f32(f32(1.0) + euclidean_distance(q, x)))


def score_rows_vectorised(rows, q, metric: str) -> np.ndarray:
    """Same arithmetic, vectorised over rows (each row keeps its own sequential order)."""
    rows = np.asarray(rows, f32)
    q = np.asarray(q, f32)
    n, d = rows.shape
    chunks = d // 8
    if metric == "euclidean":
        s = np.zeros(n, f32)
        for i in range(d):
            df = (q[i] - rows[:, i]).astype(f32)
            s = (s + (df * df).astype(f32)).astype(f32)
        dist = np.sqrt(s).astype(f32)
        return (f32(1.0) / (f32(1.0) + dist).astype(f32)).astype(f32)

    def tree(a2, b2):
        lanes = np.zeros((n, 8), f32)
        for c in range(chunks):
            lanes = (lanes + (a2[:, c * 8:c * 8 + 8] * b2[:, c * 8:c * 8 + 8]).astype(f32)).astype(f32)
        r = np.zeros(n, f32)
        for j in range(8):
            r = (r + lanes[:, j]).astype(f32)
        for i in range(chunks * 8, d):
            r = (r + (a2[:, i] * b2[:, i]).astype(f32)).astype(f32)
        return r

    qb = np.broadcast_to(q, rows.shape)
    dot = tree(qb, rows)
    if metric == "dot":
        return dot
    xm = np.sqrt(tree(rows, rows)).astype(f32)
    qm = magnitude(q)
    den = (qm * xm).astype(f32)
    with np.errstate(divide="ignore", invalid="ignore"):
        out = (dot / den).astype(f32)
    out[(xm == 0) | (qm == 0)] = 0.0
    return out


def orderable(scores: np.ndarray) -> np.ndarray:
    u = np.asarray(scores, f32).view(np.uint32).astype(np.uint64)
    nan = (u & 0x7fffffff) > 0x7f800000
    u = np.where(u == 0x80000000, 0, u)
    o = np.where(u & 0x80000000, (~u) & 0xffffffff, u | 0x80000000)
    return np.where(nan, 0, o).astype(np.uint64)


def topk(scores: np.ndarray, k: int):
    """(score desc, -0.0 == +0.0, NaN last, row asc) — the tightened total order."""
    o = orderable(scores)
    order = np.lexsort((np.arange(scores.size), -o.astype(np.int64)))
    order = order[:k]
    return order.astype(np.uint64), np.asarray(scores, f32)[order]
