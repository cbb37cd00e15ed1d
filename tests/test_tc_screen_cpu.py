"""CPU check that the GEMM epilogue's f32 screen never drops an entry the rigorous interval test
keeps ("three levels").

Restates, constant for constant, `tc_make_coef`, `tc_row_consts` / `tc_row_coef` and the fine
screen of `tc_gemm_filter_kernel` (neumann_b200/csrc/tc_prefilter_kernels.cuh) in numpy —
directed roundings emulated exactly through float64 — and checks on uniform and hostile data, for
thresholds taken from the data itself:  ub_ord >= tau_ord  ==>  the screen passes.
tests/test_gpu_tc.py::test_tc_screen_never_drops_a_keeper checks the kernels the same way."""
import numpy as np
import pytest

import oracle_ffi as o
from test_tc_model_cpu import (U, f32, interval, l2_score, quantise_query, quantise_rows, rd, ru)

EPS_Q, EPS_R = 2.0 ** -16, 2.0 ** -20
CB, CD = 0.5001 * 1.000002, 0.2502 * 1.000002


def score_to_ord(x):
    u = int(np.array(x, f32).view(np.uint32))
    if (u & 0x7fffffff) > 0x7f800000:
        return 0
    if u == 0x80000000:
        u = 0
    return (~u & 0xffffffff) if (u & 0x80000000) else (u | 0x80000000)


def ord_to_float(o_):
    if o_ == 0:
        return f32(np.nan)
    bits = (o_ ^ 0x80000000) if (o_ & 0x80000000) else (~o_ & 0xffffffff)
    return np.array(bits, np.uint32).view(f32)[()]


def make_coef(metric, tau_ord, s_q, q1, qmag, c_lo, dim):
    """tc_make_coef -> (w, u, v) as f32, or None for pass-all"""
    if tau_ord == 0:
        return None
    tau = ord_to_float(tau_ord)
    g = 2.0 * (dim + 16.0) * U
    sq = float(s_q)
    w = 0.0
    if metric == "euclidean":
        if not (tau > 0) or not (l2_score(f32(0)) >= tau):
            return None
        lo_b, hi_b = 0, 0x7f800000
        while hi_b - lo_b > 1:
            mid = lo_b + ((hi_b - lo_b) >> 1)
            if l2_score(np.array(mid, np.uint32).view(f32)[()]) >= tau:
                lo_b = mid
            else:
                hi_b = mid
        Tn = float(np.array(lo_b + 1, np.uint32).view(f32)[()])
        if not (Tn < 1e37):
            return None
        gc = (dim + 4.0) * U * 1.01
        Tp = (Tn + 2e-36) / ((1.0 - gc) * (1.0 - 2e-12))
        w = 1.0 / sq
        u = (c_lo - Tp) / sq
        v = -(CB * q1 + CD * dim)
    else:
        tp = float(tau)
        if metric == "cosine":
            if abs(tp) < 1e-30:
                tp = min(tp, 0.0) - 1e-30
            tp *= float(qmag)
        else:
            tp = float(ord_to_float(tau_ord - 1))
            if not (abs(tp) < 1e38):
                return None
            tp -= 1e-37
        u = tp / sq
        v = -((1.0 + g) * 1.000002 * (0.5001 * q1 + 0.2502 * dim))
    w *= (1.0 - EPS_Q)
    u -= EPS_Q * abs(u)
    v -= EPS_Q * abs(v)
    return rd(w), rd(u), rd(v)


def row_coef(metric, scale, x1, rmag, bad, dim):
    """tc_row_consts + tc_row_coef -> (alpha, beta, br) f32; br = inf: always rigorous"""
    if bad or not (scale >= f32(1e-15)) or not (rmag < f32(3.0e38)):
        return f32(0), f32(0), f32(np.inf)
    g = 2.0 * (dim + 16.0) * U
    rel = (dim // 8 + 24.0) * U * 1.01
    cb = CB if metric == "euclidean" else 1.000002 * (0.5001 * (1.0 + g) + 127.51 * g)
    one_minus_rel, cbf, addc = rd(1.0 - rel), ru(cb * (1.0 + EPS_R)), ru(EPS_R * (16129.0 * dim + 1.0))
    if metric == "euclidean":
        a_lo = rd(float(rd(float(rmag) * float(rmag))) * float(one_minus_rel))
        a_lo = max(rd(float(a_lo) - 1e-37), f32(0))
        two_s = f32(2.0) * f32(scale)
        rcp_rd = rd(1.0 / float(two_s))
        alpha = rd(float(rd(float(a_lo) * float(rcp_rd))) * float(f32(0.99999904)))
        beta = f32(1.0 / float(two_s))
    else:
        alpha = f32(0)
        beta = f32(float(rmag) / float(scale)) if metric == "cosine" else f32(1.0 / float(scale))
    br = ru(float(cbf) * float(ru(float(x1))) + float(addc))
    return alpha, beta, br


def fma32(a, b, c):
    return f32(float(a) * float(b) + float(c))     # exact product, one rounding (ties aside)


def screen_passes(I, alpha, beta, br, coef, shift):
    """the fine test of the epilogue: !(lhs < rhs)"""
    if coef is None:
        return True
    sc = 2.0 ** -shift
    w, u = f32(float(coef[0]) * sc), f32(float(coef[1]) * sc)
    v = rd(float(coef[2]) * sc + 12582912.0)
    brs = ru(float(br) * sc)
    if brs < f32(2000000.0):
        add_r = 0x4B400000 + int(np.ceil(float(brs))) + 5
    else:
        add_r = 0x7f000000
    lhs = np.array((I >> shift) + add_r, np.int64).astype(np.uint32).view(f32)[()]
    rhs = fma32(alpha, w, fma32(beta, u, v))
    return not (lhs < rhs)


def check(rows, queries, metric, k):
    n, dim = rows.shape
    shift = 0
    while ((16300 * dim) >> shift) >= (1 << 22) - 64:
        shift += 1
    xt, scale, x1, rmag, bad, xnorm, dnorm = quantise_rows(rows)
    kept = dropped = 0
    for q in queries:
        qt, s_q, q1, qmag, qbad, c_lo, c_hi, qnorm, enorm = quantise_query(q)
        if qbad:
            continue
        dots = xt @ qt
        iv = [interval(metric, int(dots[r]), scale[r], int(x1[r]), rmag[r], bad[r], s_q, q1, qmag,
                       c_lo, c_hi, dim, xnorm[r], dnorm[r], qnorm, enorm) for r in range(n)]
        lb_ord = np.array([0 if w else score_to_ord(lb) for lb, ub, w in iv], np.int64)
        ub_ord = np.array([0xffffffff if w else score_to_ord(ub) for lb, ub, w in iv], np.int64)
        # thresholds a refine step could produce: the k-th best lower bound of a prefix, and the
        # k-th best exact score (tighter)
        exact = o.score_rows(rows, q, metric)
        ex_ord = np.array([score_to_ord(e) for e in exact], np.int64)
        taus = {int(np.sort(lb_ord[:m])[-k]) for m in (n // 4, n // 2, n)} | \
               {int(np.sort(ex_ord[:m])[-k]) for m in (n // 2, n)}
        rc = [row_coef(metric, scale[r], int(x1[r]), rmag[r], bad[r], dim) for r in range(n)]
        for tau_ord in taus:
            coef = make_coef(metric, tau_ord, s_q, q1, qmag, c_lo, dim)
            for r in range(n):
                if ub_ord[r] >= tau_ord:
                    kept += 1
                    assert screen_passes(int(dots[r]), *rc[r], coef, shift), \
                        (metric, r, tau_ord, int(dots[r]))
                elif not screen_passes(int(dots[r]), *rc[r], coef, shift):
                    dropped += 1
    return kept, dropped


@pytest.mark.parametrize("metric", ["cosine", "euclidean", "dot"])
@pytest.mark.parametrize("dim", [24, 768, 1536])
def test_screen_is_a_superset_uniform(metric, dim):
    rows = o.fill_synthetic(400, dim, 0x5EED0001)
    qs = o.fill_synthetic(3, dim, 0x5EED1001)
    qs[1] = rows[7]
    kept, dropped = check(rows, qs, metric, 10)
    assert kept > 0 and dropped > 0          # the screen does filter


@pytest.mark.parametrize("metric", ["cosine", "euclidean", "dot"])
def test_screen_is_a_superset_hostile(metric):
    dim = 96
    rng = np.random.default_rng(3)
    rows = o.fill_synthetic(400, dim, 21)
    q = o.fill_synthetic(1, dim, 22)[0]
    rows[0:40] = q + rng.normal(0, 1e-4, (40, dim)).astype(f32)
    rows[40:60, 0] = 1000.0
    rows[60:80] *= f32(1e-12)
    rows[80:100] *= f32(1e12)
    rows[100:110] = 0.0
    rows[110:130] = -rows[0:20]
    rows[130:140] *= f32(1e-38)
    rows[140, 3] = np.nan
    rows[141, 5] = np.inf
    rows[150:200] = rng.normal(0, 1, (50, dim)).astype(f32) * f32(1e5)
    rows[200:250] = np.sign(rows[200:250]) * f32(0.5)
    queries = [q, rows[45], rows[70], rows[90], np.abs(q), q * f32(1e20), q * f32(1e-10),
               np.full(dim, 0.5, f32), -q]
    kept, _ = check(rows, np.stack(queries), metric, 7)
    assert kept > 0
