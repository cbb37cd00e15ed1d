"""ctypes binding of include/neumann_b200.h (the drop-in C ABI).

Python is only the test/bench harness here; the product is the shared library.  Loading fails
loudly when the library has not been built: there is no Python or CPU fallback for the scan.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

_LIB_PATH = Path(__file__).resolve().parent / "libneumann_b200.so"

NM_COSINE, NM_EUCLIDEAN, NM_DOT_PRODUCT = 0, 1, 2
NM_OK = 0
NM_ERR_EMPTY_VECTOR = 1
NM_ERR_INVALID_TOP_K = 2
NM_ERR_DIMENSION_MISMATCH = 3
NM_ERR_STORAGE = 4
NM_ERR_SEARCH_TIMEOUT = 5
NM_ERR_INVALID_ARGUMENT = 6
NM_ERR_NOT_FOUND = 7
NM_ERR_CONFIGURATION = 8
NM_ERR_COLLECTION_EXISTS = 9
NM_ERR_COLLECTION_NOT_FOUND = 10
NM_COMM_ID_BYTES = 128
NM_TOPK_FAST_MAX = 1024


class NmStats(C.Structure):
    _fields_ = [
        ("searches", C.c_uint64), ("rows_scanned", C.c_uint64), ("bytes_streamed", C.c_uint64),
        ("scan_launches", C.c_uint64), ("merge_launches", C.c_uint64),
        ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64), ("last_scan_ms", C.c_double),
        ("profiled_scan_ms", C.c_double), ("profiled_scans", C.c_uint64),
        ("prefilter_queries", C.c_uint64), ("prefilter_fallbacks", C.c_uint64),
        ("prefilter_kept", C.c_uint64),
        ("coalesced_batches", C.c_uint64), ("coalesced_queries", C.c_uint64),
        ("tc_queries", C.c_uint64), ("tc_fallbacks", C.c_uint64), ("tc_survivors", C.c_uint64),
        ("filter_masks_built", C.c_uint64), ("filter_mask_hits", C.c_uint64),
    ]


class NmShardInfo(C.Structure):
    _fields_ = [
        ("device", C.c_int), ("rows", C.c_uint64), ("row_base", C.c_uint64),
        ("capacity_rows", C.c_uint64), ("mapped_bytes", C.c_uint64), ("reserved_bytes", C.c_uint64),
        ("chunks", C.c_uint64), ("remaps", C.c_uint64), ("q8_rows", C.c_uint64),
        ("grows_in_place", C.c_int),
    ]


class NmFilterOp(C.Structure):
    _fields_ = [
        ("kind", C.c_uint8), ("cmp", C.c_uint8), ("lit_tag", C.c_uint8), ("reserved", C.c_uint8),
        ("column", C.c_uint32), ("lit", C.c_uint64), ("table_off", C.c_uint32), ("table_bits", C.c_uint32),
    ]


NM_V_MISSING, NM_V_NULL, NM_V_BOOL, NM_V_INT, NM_V_FLOAT, NM_V_STRING = range(6)
NM_F_TRUE, NM_F_FALSE, NM_F_AND, NM_F_OR, NM_F_EXISTS, NM_F_CMP, NM_F_STR_TABLE = range(7)
NM_C_EQ, NM_C_NE, NM_C_LT, NM_C_LE, NM_C_GT, NM_C_GE = range(6)

_f32p = C.POINTER(C.c_float)
_u64p = C.POINTER(C.c_uint64)
_u32p = C.POINTER(C.c_uint32)
_vp = C.c_void_p

# name -> (restype, argtypes): exactly the declarations of include/neumann_b200.h
SIGNATURES = {
    "nm_abi_version": (C.c_int, []),
    "nm_last_error": (C.c_char_p, []),
    "nm_device_count": (C.c_int, []),
    "nm_index_create": (C.c_int, [C.c_uint32, C.POINTER(C.c_int), C.c_int, C.POINTER(_vp)]),
    "nm_index_destroy": (None, [_vp]),
    "nm_index_load": (C.c_int, [_vp, _vp, C.c_uint64]),
    "nm_index_append": (C.c_int, [_vp, _vp, C.c_uint64]),
    "nm_index_update": (C.c_int, [_vp, C.c_uint64, _vp]),
    "nm_index_swap_remove": (C.c_int, [_vp, C.c_uint64, _u64p]),
    "nm_index_clear": (C.c_int, [_vp]),
    "nm_index_get_row": (C.c_int, [_vp, C.c_uint64, _vp]),
    "nm_index_get_rows": (C.c_int, [_vp, C.c_uint64, C.c_uint64, _vp]),
    "nm_index_rows": (C.c_uint64, [_vp]),
    "nm_index_dim": (C.c_uint32, [_vp]),
    "nm_index_device_count": (C.c_int, [_vp]),
    "nm_index_shard_info": (C.c_int, [_vp, C.c_int, C.POINTER(NmShardInfo)]),
    "nm_index_fill_synthetic": (C.c_int, [_vp, C.c_uint64, C.c_uint64, C.c_uint64]),
    "nm_search": (C.c_int, [_vp, _vp, C.c_uint32, C.c_uint32, C.c_int, _vp, _vp, _vp]),
    "nm_search_masked": (C.c_int, [_vp, _vp, C.c_uint32, C.c_uint32, C.c_int, _vp, _vp, _vp, _vp]),
    "nm_search_device": (C.c_int, [_vp, _vp, C.c_uint32, C.c_uint32, C.c_int, _vp, _vp, _vp, _vp]),
    "nm_index_column_set": (C.c_int, [_vp, C.c_uint32, C.c_uint64, C.c_uint64, _vp, _vp]),
    "nm_search_filtered": (C.c_int, [_vp, _vp, C.c_uint32, C.c_uint32, C.c_int, _vp, C.c_uint32, _vp,
                                     C.c_uint32, _vp, _vp, _vp]),
    "nm_index_filter_mask": (C.c_int, [_vp, _vp, C.c_uint32, _vp, C.c_uint32, _vp, _u64p]),
    "nm_comm_create_id": (C.c_int, [_vp]),
    "nm_index_attach_comm": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_uint64]),
    "nm_index_detach_comm": (C.c_int, [_vp]),
    "nm_index_stats": (C.c_int, [_vp, C.POINTER(NmStats)]),
    "nm_index_set_profiling": (C.c_int, [_vp, C.c_int]),
    "nm_index_set_batching": (C.c_int, [_vp, C.c_int]),
    "nm_index_set_pipelining": (C.c_int, [_vp, C.c_int]),
    "nm_index_release_stream": (C.c_int, [_vp, _vp]),
    "nm_index_set_prefilter": (C.c_int, [_vp, C.c_int]),
    "nm_index_set_coalescing": (C.c_int, [_vp, C.c_int]),
    "nm_index_set_tensor_core": (C.c_int, [_vp, C.c_int]),
    "nm_debug_tc_dots": (C.c_int, [_vp, _vp, C.c_uint32, _vp]),
    "nm_debug_q8_row": (C.c_int, [_vp, C.c_uint64, _vp, _f32p]),
}

_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not _LIB_PATH.exists():
            raise RuntimeError(
                f"{_LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU fallback for the SIMILAR scan)")
        l = C.CDLL(str(_LIB_PATH))
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


class NmError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"nm_status {code}: {msg}")
        self.code = code
        self.msg = msg


def check(code: int) -> None:
    if code != NM_OK:
        raise NmError(code, lib().nm_last_error().decode("utf-8", "replace"))
