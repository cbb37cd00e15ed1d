"""CPU check of the error model behind the tensor-core batch pre-filter.

`tc_interval` (neumann_b200/csrc/tc_prefilter_kernels.cuh) turns the exact integer dot product
of an int8 row and an int8 query into an interval [lb, ub] that must contain the score the
reference arithmetic produces (vector_engine/src/lib.rs:2231-2266).  This file restates that
function in numpy, constant for constant, and checks the containment against the oracle on
friendly and hostile data, so the rigor claim does not rest on GPU runs alone.  (Keep the
constants in step with the CUDA source; tests/test_gpu_tc.py checks the kernels themselves.)"""
import numpy as np
import pytest

import oracle_ffi as o

f32 = np.float32
U = 2.0 ** -24


def quantise_rows(x):
    """quantize_rows_kernel: per-row scale, int8 values, sum |xt|, reference |x|, flags."""
    mx = np.abs(x).max(axis=1).astype(f32)
    bad = ~np.isfinite(x).all(axis=1)
    scale = (mx / f32(127.0)).astype(f32)
    bad |= (mx > 0) & (scale < f32(1.17549435e-38))
    scale = np.where(bad, f32(0), scale)
    with np.errstate(divide="ignore", invalid="ignore"):
        xt = np.where(scale[:, None] > 0, np.rint((x / scale[:, None]).astype(f32)), 0.0)
    xt = np.clip(np.nan_to_num(xt), -127, 127).astype(np.int64)
    rmag = np.array([o.magnitude(r) for r in x], f32)
    # (>= ||xt||_2, >= ||x / s - xt||_2) as the kernel stores them: f32 quotient, slack for its
    # rounding, rounded up
    with np.errstate(divide="ignore", invalid="ignore"):
        qf = np.where(scale[:, None] > 0, (x / scale[:, None]).astype(f32), 0.0).astype(np.float64)
    d = np.nan_to_num(qf - xt)
    xnorm = np.sqrt((xt.astype(np.float64) ** 2).sum(axis=1)) * (1 + 1e-7)
    dnorm = (np.sqrt((d ** 2).sum(axis=1)) + 7.7e-6 * np.sqrt(x.shape[1])) * (1 + 1e-6)
    return xt, scale, np.abs(xt).sum(axis=1), rmag, bad, xnorm, dnorm


def quantise_query(q):
    mx = f32(np.abs(q).max())
    s = f32(mx / f32(127.0))
    bad = (not np.isfinite(q).all()) or not (s >= f32(1e-15))
    qt = np.zeros(q.shape, np.int64) if bad else np.clip(np.rint((q / s).astype(f32)), -127, 127).astype(np.int64)
    c = float(np.sum(q.astype(np.float64) ** 2))
    e = np.zeros(q.shape) if bad else (q / s).astype(f32).astype(np.float64) - qt
    qnorm = float(np.sqrt((qt.astype(np.float64) ** 2).sum())) * (1 + 1e-9)
    enorm = (float(np.sqrt((e ** 2).sum())) + 7.7e-6 * np.sqrt(q.size)) * (1 + 1e-9)
    return qt, (f32(0) if bad else s), int(np.abs(qt).sum()), o.magnitude(q), bad, \
        max(c * (1 - 1e-12) - 1e-40, 0.0), c * (1 + 1e-12) + 1e-40, qnorm, enorm


def rd(x):
    """__double2float_rd"""
    with np.errstate(over="ignore"):
        y = f32(x)
    return y if float(y) <= x else np.nextafter(y, f32(-np.inf))


def ru(x):
    """__double2float_ru"""
    with np.errstate(over="ignore"):
        y = f32(x)
    return y if float(y) >= x else np.nextafter(y, f32(np.inf))


def l2_score(s):
    with np.errstate(over="ignore", invalid="ignore"):
        return f32(1.0) / (f32(1.0) + np.sqrt(f32(s)))


def interval(metric, I, scale, x1, rmag, bad_row, s_q, q1, qmag, c_lo, c_hi, dim, xnorm, dnorm,
             qnorm, enorm):
    """-> (lb, ub, wild) as in tc_interval; wild == always a candidate."""
    wild = bool(bad_row)
    g = 2.0 * (dim + 16.0) * U
    ss = float(s_q) * float(scale)
    BL = 0.5001 * (q1 + x1) + 0.2502 * dim
    B = min(0.5001 * q1, qnorm * dnorm) + min(0.5001 * x1, xnorm * enorm) + \
        min(0.2502 * dim, enorm * dnorm)
    Dt = ss * float(I)
    if metric == "euclidean":
        E = ss * B * 1.000001 + 1e-37
        a = float(rmag) * float(rmag)
        rel = (dim // 8 + 24.0) * U * 1.01
        a_lo, a_hi = max(a * (1 - rel) - 1e-37, 0.0), a * (1 + rel) + 1e-37
        gc = (dim + 4.0) * U * 1.01
        lo = (a_lo + c_lo - 2.0 * (Dt + E)) * (1 - gc) * (1 - 1e-12) - 1e-36
        hi = (a_hi + c_hi - 2.0 * (Dt - E)) * (1 + gc) * (1 + 1e-12) + 1e-36
        if not (hi < 1e37) or not (a_hi < 1e37) or not (c_hi < 1e37):
            wild = True
        slo = rd(lo) if lo > 0 else f32(0)
        shi = ru(hi) if hi > 0 else f32(0)
        return l2_score(shi), l2_score(slo), wild
    S = 127.51 * x1 + BL
    E = ss * (B + g * S) * 1.000001 + 1e-37
    if not (ss * S < 1e37) or not (abs(Dt) + E < 1e37):
        return f32(0), f32(0), True
    lo, hi = rd(Dt - E), ru(Dt + E)
    if metric == "dot":
        return lo, hi, wild
    if qmag == 0 or rmag == 0:
        return f32(0), f32(0), wild
    den = f32(qmag) * f32(rmag)
    if not (den > f32(1.17549435e-38)) or not (den < f32(3.0e38)):
        return f32(0), f32(0), True
    return f32(lo / den), f32(hi / den), wild


def check(rows, queries, metric):
    n, dim = rows.shape
    xt, scale, x1, rmag, bad, xnorm, dnorm = quantise_rows(rows)
    width = []
    for q in queries:
        qt, s_q, q1, qmag, qbad, c_lo, c_hi, qnorm, enorm = quantise_query(q)
        if qbad:
            continue                      # such queries are redone by the exact path
        exact = o.score_rows(rows, q, metric)
        dots = xt @ qt
        for r in range(n):
            lb, ub, wild = interval(metric, int(dots[r]), scale[r], int(x1[r]), rmag[r], bad[r],
                                    s_q, q1, qmag, c_lo, c_hi, dim, xnorm[r], dnorm[r], qnorm, enorm)
            if wild:
                continue
            e = exact[r]
            assert not np.isnan(e), (metric, r)
            assert lb <= e <= ub, (metric, r, float(lb), float(e), float(ub))
            width.append(float(ub) - float(lb))
    return width


@pytest.mark.parametrize("metric", ["cosine", "euclidean", "dot"])
@pytest.mark.parametrize("dim", [8, 100, 768, 1536])
def test_interval_contains_reference_score_uniform(metric, dim):
    rows = o.fill_synthetic(300, dim, 0x5EED0001)
    qs = o.fill_synthetic(3, dim, 0x5EED1001)
    qs[1] = rows[7]
    w = check(rows, qs, metric)
    assert len(w) == 900 and max(w) < (1.0 if metric != "dot" else 0.05 * dim)


@pytest.mark.parametrize("metric", ["cosine", "euclidean", "dot"])
def test_interval_contains_reference_score_hostile(metric):
    dim = 96
    rng = np.random.default_rng(3)
    rows = o.fill_synthetic(400, dim, 21)
    q = o.fill_synthetic(1, dim, 22)[0]
    rows[0:40] = q + rng.normal(0, 1e-4, (40, dim)).astype(f32)       # inside the int8 error
    rows[40:60, 0] = 1000.0                                             # outlier: coarse scale
    rows[60:80] *= f32(1e-20)
    rows[80:100] *= f32(1e15)
    rows[100:110] = 0.0
    rows[110:130] = -rows[0:20]
    rows[130:140] *= f32(1e-38)
    rows[140, 3] = np.nan
    rows[141, 5] = np.inf
    rows[142] *= f32(3e38)
    rows[150:200] = rng.normal(0, 1, (50, dim)).astype(f32) * f32(1e5)
    rows[200:250] = np.sign(rows[200:250]) * f32(0.5)                  # every element on a rounding tie
    queries = [q, rows[45], rows[70], rows[90], np.abs(q), q * f32(1e30), q * f32(1e-10),
               np.full(dim, 0.5, f32)]
    check(rows, np.stack(queries), metric)
