"""Thin object wrapper over the nm_index_* / nm_search C ABI (test + bench harness)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _ffi
from ._ffi import NM_COSINE, NM_DOT_PRODUCT, NM_EUCLIDEAN, NmError, NmStats, check  # noqa: F401

METRICS = {"cosine": NM_COSINE, "euclidean": NM_EUCLIDEAN, "l2": NM_EUCLIDEAN,
           "dot": NM_DOT_PRODUCT, "dot_product": NM_DOT_PRODUCT}


def _metric(m) -> int:
    return METRICS[m.lower()] if isinstance(m, str) else int(m)


class DeviceIndex:
    """Device mirror of the `emb:` rows; see include/neumann_b200.h."""

    def __init__(self, dim: int, devices: list[int] | None = None):
        self._h = C.c_void_p()
        lib = _ffi.lib()
        if devices:
            arr = (C.c_int * len(devices))(*devices)
            check(lib.nm_index_create(dim, arr, len(devices), C.byref(self._h)))
        else:
            check(lib.nm_index_create(dim, None, 0, C.byref(self._h)))
        self.dim = dim

    def close(self) -> None:
        if self._h:
            _ffi.lib().nm_index_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        return self._h

    @property
    def rows(self) -> int:
        return int(_ffi.lib().nm_index_rows(self._h))

    @staticmethod
    def _rows_arg(rows: np.ndarray, dim: int) -> np.ndarray:
        a = np.ascontiguousarray(rows, dtype=np.float32)
        if a.ndim == 1:
            a = a.reshape(1, -1)
        if a.shape[1] != dim:
            raise ValueError(f"rows have dim {a.shape[1]}, index has {dim}")
        return a

    def load(self, rows: np.ndarray) -> None:
        a = self._rows_arg(rows, self.dim) if len(rows) else np.zeros((0, self.dim), np.float32)
        check(_ffi.lib().nm_index_load(self._h, a.ctypes.data, a.shape[0]))

    def append(self, rows: np.ndarray) -> None:
        a = self._rows_arg(rows, self.dim)
        check(_ffi.lib().nm_index_append(self._h, a.ctypes.data, a.shape[0]))

    def update(self, row: int, vec: np.ndarray) -> None:
        a = self._rows_arg(vec, self.dim)
        check(_ffi.lib().nm_index_update(self._h, row, a.ctypes.data))

    def swap_remove(self, row: int) -> int:
        moved = C.c_uint64()
        check(_ffi.lib().nm_index_swap_remove(self._h, row, C.byref(moved)))
        return int(moved.value)

    def clear(self) -> None:
        check(_ffi.lib().nm_index_clear(self._h))

    def get_row(self, row: int) -> np.ndarray:
        out = np.empty(self.dim, np.float32)
        check(_ffi.lib().nm_index_get_row(self._h, row, out.ctypes.data))
        return out

    def get_rows(self, first: int, n: int, out: np.ndarray | None = None) -> np.ndarray:
        """Rows [first, first+n) as a dense float32 [n, dim] host array (into `out` if given)."""
        if out is None:
            out = np.empty((n, self.dim), np.float32)
        assert out.dtype == np.float32 and out.flags.c_contiguous and out.size >= n * self.dim
        check(_ffi.lib().nm_index_get_rows(self._h, first, n, out.ctypes.data))
        return out.reshape(-1)[:n * self.dim].reshape(n, self.dim)

    def shard_info(self, shard: int = 0) -> "_ffi.NmShardInfo":
        info = _ffi.NmShardInfo()
        check(_ffi.lib().nm_index_shard_info(self._h, shard, C.byref(info)))
        return info

    @property
    def n_shards(self) -> int:
        return int(_ffi.lib().nm_index_device_count(self._h))

    def fill_synthetic(self, n: int, seed: int, row_offset: int = 0) -> None:
        check(_ffi.lib().nm_index_fill_synthetic(self._h, n, seed, row_offset))

    def search(self, queries: np.ndarray, k: int, metric="cosine"):
        """-> list of (rows uint64[m], scores float32[m]) per query."""
        q = np.ascontiguousarray(queries, dtype=np.float32)
        if q.ndim == 1:
            q = q.reshape(1, -1)
        nq = q.shape[0]
        if q.shape[1] != self.dim:
            raise NmError(_ffi.NM_ERR_DIMENSION_MISMATCH,
                          f"query dim {q.shape[1]} != index dim {self.dim}")
        kk = max(int(k), 1)
        rows = np.zeros((nq, kk), np.uint64)
        scores = np.zeros((nq, kk), np.float32)
        counts = np.zeros(nq, np.uint32)
        check(_ffi.lib().nm_search(self._h, q.ctypes.data, nq, int(k), _metric(metric),
                                   rows.ctypes.data, scores.ctypes.data, counts.ctypes.data))
        return [(rows[i, :counts[i]].copy(), scores[i, :counts[i]].copy()) for i in range(nq)]

    def search_masked(self, queries: np.ndarray, k: int, metric, mask: np.ndarray):
        """mask: bool array over rows (True = eligible).  -> like search()."""
        q = np.ascontiguousarray(queries, dtype=np.float32)
        if q.ndim == 1:
            q = q.reshape(1, -1)
        nq = q.shape[0]
        bits = np.packbits(np.asarray(mask, dtype=bool), bitorder="little")
        words = np.zeros((bits.size + 7) // 8 + 1, np.uint64)
        words.view(np.uint8)[:bits.size] = bits
        kk = max(int(k), 1)
        rows = np.zeros((nq, kk), np.uint64)
        scores = np.zeros((nq, kk), np.float32)
        counts = np.zeros(nq, np.uint32)
        check(_ffi.lib().nm_search_masked(self._h, q.ctypes.data, nq, int(k), _metric(metric),
                                          words.ctypes.data, rows.ctypes.data, scores.ctypes.data,
                                          counts.ctypes.data))
        return [(rows[i, :counts[i]].copy(), scores[i, :counts[i]].copy()) for i in range(nq)]

    # ---- metadata columns + device-side filters (include/neumann_b200.h) ----
    def column_set(self, column: int, first_row: int, tags, values) -> None:
        t = np.ascontiguousarray(tags, dtype=np.uint8)
        v = np.ascontiguousarray(values, dtype=np.uint64)
        assert t.shape == v.shape and t.ndim == 1
        check(_ffi.lib().nm_index_column_set(self._h, column, first_row, t.size, t.ctypes.data,
                                             v.ctypes.data))

    @staticmethod
    def _program(ops, tables):
        arr = (_ffi.NmFilterOp * len(ops))(*ops)
        tab = np.ascontiguousarray(tables if tables is not None else np.zeros(0), dtype=np.uint32)
        return arr, tab

    def search_filtered(self, queries: np.ndarray, k: int, metric, ops, tables=None):
        """ops: list of _ffi.NmFilterOp in postfix order; tables: uint32 words.  -> like search()."""
        q = np.ascontiguousarray(queries, dtype=np.float32)
        if q.ndim == 1:
            q = q.reshape(1, -1)
        nq = q.shape[0]
        arr, tab = self._program(ops, tables)
        kk = max(int(k), 1)
        rows = np.zeros((nq, kk), np.uint64)
        scores = np.zeros((nq, kk), np.float32)
        counts = np.zeros(nq, np.uint32)
        check(_ffi.lib().nm_search_filtered(self._h, q.ctypes.data, nq, int(k), _metric(metric), arr,
                                            len(ops), tab.ctypes.data if tab.size else None, tab.size,
                                            rows.ctypes.data, scores.ctypes.data, counts.ctypes.data))
        return [(rows[i, :counts[i]].copy(), scores[i, :counts[i]].copy()) for i in range(nq)]

    def filter_mask(self, ops, tables=None) -> np.ndarray:
        """Evaluate a filter program on the device: bool array over this process's rows."""
        arr, tab = self._program(ops, tables)
        n = self.rows
        words = np.zeros((n + 63) // 64 + 1, np.uint64)
        elig = C.c_uint64()
        check(_ffi.lib().nm_index_filter_mask(self._h, arr, len(ops), tab.ctypes.data if tab.size else None,
                                              tab.size, words.ctypes.data, C.byref(elig)))
        bits = np.unpackbits(words.view(np.uint8), bitorder="little")[:n].astype(bool)
        assert int(bits.sum()) == int(elig.value)
        return bits

    def search_device(self, d_queries_ptr: int, nq: int, k: int, metric, d_rows_ptr: int,
                      d_scores_ptr: int, d_counts_ptr: int, stream_ptr: int = 0) -> None:
        check(_ffi.lib().nm_search_device(self._h, d_queries_ptr, nq, k, _metric(metric),
                                          d_rows_ptr, d_scores_ptr, d_counts_ptr, stream_ptr))

    def attach_comm(self, comm_id: bytes, n_ranks: int, rank: int, row_base: int) -> None:
        buf = C.create_string_buffer(comm_id, _ffi.NM_COMM_ID_BYTES)
        check(_ffi.lib().nm_index_attach_comm(self._h, buf, n_ranks, rank, row_base))

    def detach_comm(self) -> None:
        check(_ffi.lib().nm_index_detach_comm(self._h))

    def set_prefilter(self, mode: int) -> None:
        """0 = off, 1 = int8 copy for single queries and batches, 2 = auto (default): batches only,
        copy built by the first eligible batch (see include/neumann_b200.h)."""
        check(_ffi.lib().nm_index_set_prefilter(self._h, int(mode)))

    def set_tensor_core(self, enable: bool) -> None:
        """Batches on a pre-filtered index: tcgen05 int8 GEMM pre-filter (default on) or not."""
        check(_ffi.lib().nm_index_set_tensor_core(self._h, 1 if enable else 0))

    def debug_tc_dots(self, queries: np.ndarray) -> np.ndarray:
        """Exact integer dot products of the tensor-core pass: int32 [nq, rows]."""
        q = np.ascontiguousarray(queries, dtype=np.float32).reshape(-1, self.dim)
        out = np.zeros((min(q.shape[0], 256), self.rows), np.int32)
        check(_ffi.lib().nm_debug_tc_dots(self._h, q.ctypes.data, q.shape[0], out.ctypes.data))
        return out

    def debug_q8_row(self, row: int):
        """(int8[dim], scale) of one row of the int8 copy."""
        out = np.zeros(self.dim, np.int8)
        scale = C.c_float()
        check(_ffi.lib().nm_debug_q8_row(self._h, row, out.ctypes.data, C.byref(scale)))
        return out, float(scale.value)

    def set_coalescing(self, max_batch: int) -> None:
        check(_ffi.lib().nm_index_set_coalescing(self._h, int(max_batch)))

    def set_batching(self, enable: bool) -> None:
        check(_ffi.lib().nm_index_set_batching(self._h, 1 if enable else 0))

    def set_pipelining(self, enable: bool) -> None:
        """Consecutive async single-query search_device calls on one stream overlap (see header)."""
        check(_ffi.lib().nm_index_set_pipelining(self._h, 1 if enable else 0))

    def release_stream(self, stream_ptr: int) -> None:
        check(_ffi.lib().nm_index_release_stream(self._h, stream_ptr))

    def set_profiling(self, enable: bool) -> None:
        check(_ffi.lib().nm_index_set_profiling(self._h, 1 if enable else 0))

    def stats(self) -> NmStats:
        s = NmStats()
        check(_ffi.lib().nm_index_stats(self._h, C.byref(s)))
        return s


def comm_create_id() -> bytes:
    buf = C.create_string_buffer(_ffi.NM_COMM_ID_BYTES)
    check(_ffi.lib().nm_comm_create_id(buf))
    return buf.raw


def device_count() -> int:
    return int(_ffi.lib().nm_device_count())
