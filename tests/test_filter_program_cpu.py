"""Host logic of the device-side filter path, without a device: VectorEngine stages the metadata
of its rows as typed columns (dictionary-encoded strings) and compiles a FilterCondition to the
postfix nm_filter_op program that filter_mask_kernel runs.  Here the program is executed by a
Python restatement of the kernel (same type rules: vector_engine/src/lib.rs:3592-3684) over the
staged columns and must give, row by row, what the host's evaluate_filter gives — and what an
independent Python evaluation of the WHERE expression gives."""
import struct

import numpy as np

from neumann_b200 import engine as eng
from neumann_b200._ffi import (NM_C_EQ, NM_C_GE, NM_C_GT, NM_C_LE, NM_C_LT, NM_C_NE, NM_F_AND, NM_F_CMP,
                               NM_F_EXISTS, NM_F_FALSE, NM_F_OR, NM_F_STR_TABLE, NM_F_TRUE, NM_V_BOOL,
                               NM_V_FLOAT, NM_V_INT, NM_V_MISSING, NM_V_NULL, NM_V_STRING)
from test_gpu_filter import WHERE_CASES, _meta, _py_eval


def _f64(bits: int) -> float:
    return struct.unpack("<d", struct.pack("<Q", bits))[0]


def _i64(bits: int) -> int:
    return bits - (1 << 64) if bits >= (1 << 63) else bits


def _cmp_holds(c, cmp):
    if c is None:
        return False
    return {NM_C_EQ: c == 0, NM_C_NE: c != 0, NM_C_LT: c < 0, NM_C_LE: c <= 0, NM_C_GT: c > 0,
            NM_C_GE: c >= 0}[cmp]


def _ordf(a, b):
    if a != a or b != b:
        return None
    return (a > b) - (a < b)


def run_program(prog: dict, row: int) -> bool:
    """filter_mask_kernel for one row (filter_kernels.cuh)."""
    stack = []
    for kind, cmp, lit_tag, column, lit, table_off, table_bits in prog["ops"]:
        if kind == NM_F_AND:
            b, a = stack.pop(), stack.pop()
            stack.append(a and b)
            continue
        if kind == NM_F_OR:
            b, a = stack.pop(), stack.pop()
            stack.append(a or b)
            continue
        if kind == NM_F_TRUE:
            stack.append(True)
            continue
        if kind == NM_F_FALSE:
            stack.append(False)
            continue
        col = prog["columns"].get(str(column))
        tag = col["tags"][row] if col else NM_V_MISSING
        val = col["vals"][row] if col else 0
        if kind == NM_F_EXISTS:
            stack.append(tag != NM_V_MISSING)
        elif tag == NM_V_MISSING:
            stack.append(False)
        elif kind == NM_F_STR_TABLE:
            ok = tag == NM_V_STRING and val < table_bits and \
                (prog["tables"][table_off + val // 32] >> (val % 32)) & 1
            stack.append(bool(ok))
        else:
            assert kind == NM_F_CMP
            c = None
            if lit_tag == NM_V_INT:
                if tag == NM_V_INT:
                    a, b = _i64(val), _i64(lit)
                    c = (a > b) - (a < b)
                elif tag == NM_V_FLOAT:
                    c = _ordf(_f64(val), float(_i64(lit)))
            elif lit_tag == NM_V_FLOAT:
                if tag == NM_V_FLOAT:
                    c = _ordf(_f64(val), _f64(lit))
                elif tag == NM_V_INT:
                    c = _ordf(float(_i64(val)), _f64(lit))
            elif lit_tag == NM_V_BOOL:
                if tag == NM_V_BOOL:
                    c = int(val != 0) - int(lit != 0)
            elif lit_tag == NM_V_NULL:
                if tag == NM_V_NULL:
                    c = 0
            stack.append(_cmp_holds(c, cmp))
    assert len(stack) == 1
    return bool(stack[0])


EXTRA_CASES = [
    "price >= -1", "price < 0", "bucket = 7.0", "bucket > 48.5 OR bucket < 0.5", "opt = false",
    "opt = null", "name != 'n0007'", "name <= 'n0010'", "name > 'n2990' AND STARTS_WITH(name, 'n29')",
    "CONTAINS(name, 'n')", "CONTAINS(bucket, '1')", "STARTS_WITH(nothing, 'x')", "EXISTS(nothing)",
    "mixed != 3", "mixed < 'zzz'", "mixed = true", "mixed IN (3, 3.5, 'three', true)", "mixed IN ()",
    "((bucket < 10 AND price > 3) OR (bucket > 40 AND price < 3)) AND NOT_A_FIELD = 1",
    "bucket < 25 AND (opt = true OR name < 'n1000')",
]


def test_compiled_program_agrees_with_evaluate_filter_row_by_row():
    n, d = 3000, 4
    e = eng.VectorEngine()
    metas = [_meta(i) for i in range(n)]
    metas[17] = {}                                      # a row without any metadata
    metas[18] = {"bucket": float("nan"), "price": -0.0, "name": ""}
    metas[19] = {"bucket": 2**62, "price": 1e308, "name": "é"}
    vec = np.ones(d, np.float32)
    for i in range(n):
        e.store_embedding_with_metadata(f"k{i:04d}", vec, metas[i])
    for w in WHERE_CASES + EXTRA_CASES:
        prog = e.debug_filter_program(d, w)
        assert prog["rows"] == n
        depth = 0
        for op in prog["ops"]:
            depth += -1 if op[0] in (NM_F_AND, NM_F_OR) else 1
            assert depth >= 1
        assert depth == 1 and len(prog["ops"]) <= 128, w
        got = np.array([run_program(prog, r) for r in range(n)])
        host = np.array(prog["host"], bool)
        assert np.array_equal(got, host), (w, np.nonzero(got != host)[0][:5])
        if w in WHERE_CASES:                            # ... and with the independent evaluation
            want = np.array([bool(_py_eval(metas[r], w)) for r in range(n)])
            want[18] = host[18]                         # (NaN / -0.0 row: covered by `host` only)
            assert np.array_equal(got, want), w
        assert int(host.sum()) == e.count_matching(w), w
    # string columns are dictionary-encoded once per DISTINCT string
    prog = e.debug_filter_program(d, "name = 'n0042'")
    (op,) = prog["ops"]
    assert op[0] == NM_F_STR_TABLE and op[6] >= n - 3 and sum(bin(x).count("1") for x in prog["tables"]) == 1
    prog = e.debug_filter_program(d, "mixed = 'three'")
    assert prog["ops"][0][6] == 1                        # one distinct string in that column
    e.close()
