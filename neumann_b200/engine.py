"""Python view of the C++ host mirror (VectorEngine / QueryRouter) through
include/neumann_b200_engine.h.  Test + bench harness only: method names, argument meaning and
error behaviour follow vector_engine/src/lib.rs so the parity tests read like the reference's.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _ffi

COSINE, EUCLIDEAN, DOT_PRODUCT = 0, 1, 2
AUTO, PRE_FILTER, POST_FILTER = 0, 1, 2


def encode_metadata(meta: dict | None) -> bytes:
    """dict -> the typed wire string of include/neumann_b200_engine.h."""
    recs = []
    for name, v in (meta or {}).items():
        if v is None:
            t, val = "n", ""
        elif isinstance(v, bool):
            t, val = "b", "1" if v else "0"
        elif isinstance(v, int):
            t, val = "i", str(v)
        elif isinstance(v, float):
            t, val = "f", repr(v)
        else:
            t, val = "s", str(v)
        recs.append(f"{name}\x1e{t}\x1e{val}")
    return "\x1f".join(recs).encode()


class NmEngineConfig(C.Structure):
    _fields_ = [
        ("default_dimension", C.c_uint64), ("sparse_threshold", C.c_float),
        ("parallel_threshold", C.c_uint64), ("default_metric", C.c_int),
        ("max_dimension", C.c_uint64), ("search_timeout_ms", C.c_int64),
        ("n_devices", C.c_int), ("devices", C.c_int * 8),
        ("device_prefilter", C.c_int), ("max_keys_per_scan", C.c_uint64),
    ]


_vp, _cp, _sz, _u64 = C.c_void_p, C.c_char_p, C.c_size_t, C.c_uint64
_pvp = C.POINTER(C.c_void_p)
ENGINE_SIGNATURES = {
    "nm_engine_config_default": (None, [C.POINTER(NmEngineConfig)]),
    "nm_engine_create": (C.c_int, [C.POINTER(NmEngineConfig), _pvp]),
    "nm_engine_destroy": (None, [_vp]),
    "nm_engine_last_error": (_cp, []),
    "nm_engine_store_embedding": (C.c_int, [_vp, _cp, _vp, _sz]),
    "nm_engine_get_embedding": (C.c_int, [_vp, _cp, _vp, _sz, C.POINTER(_sz)]),
    "nm_engine_delete_embedding": (C.c_int, [_vp, _cp]),
    "nm_engine_exists": (C.c_int, [_vp, _cp]),
    "nm_engine_count": (_u64, [_vp]),
    "nm_engine_search_similar": (C.c_int, [_vp, _vp, _sz, _sz, _pvp]),
    "nm_engine_search_similar_with_metric": (C.c_int, [_vp, _vp, _sz, _sz, C.c_int, _pvp]),
    "nm_engine_search_similar_batch": (C.c_int, [_vp, _vp, _sz, _sz, _sz, C.c_int, _pvp]),
    "nm_engine_compute_similarity": (C.c_int, [_vp, _sz, _vp, _sz, C.POINTER(C.c_float)]),
    "nm_engine_create_collection": (C.c_int, [_vp, _cp, _u64, C.c_int]),
    "nm_engine_delete_collection": (C.c_int, [_vp, _cp]),
    "nm_engine_collection_exists": (C.c_int, [_vp, _cp]),
    "nm_engine_store_in_collection": (C.c_int, [_vp, _cp, _cp, _vp, _sz]),
    "nm_engine_delete_from_collection": (C.c_int, [_vp, _cp, _cp]),
    "nm_engine_collection_count": (_u64, [_vp, _cp]),
    "nm_engine_search_in_collection": (C.c_int, [_vp, _cp, _vp, _sz, _sz, _pvp]),
    "nm_engine_store_embedding_with_metadata": (C.c_int, [_vp, _cp, _vp, _sz, _cp]),
    "nm_engine_store_in_collection_with_metadata": (C.c_int, [_vp, _cp, _cp, _vp, _sz, _cp]),
    "nm_engine_search_similar_filtered": (C.c_int, [_vp, _vp, _sz, _sz, _cp, C.c_int, _sz, _pvp]),
    "nm_engine_search_filtered_in_collection": (C.c_int, [_vp, _cp, _vp, _sz, _sz, _cp, C.c_int, _sz, _pvp]),
    "nm_engine_count_matching": (C.c_int, [_vp, _cp, C.POINTER(_u64)]),
    "nm_engine_update_metadata": (C.c_int, [_vp, _cp, _cp]),
    "nm_engine_remove_metadata_field": (C.c_int, [_vp, _cp, _cp]),
    "nm_engine_has_metadata_field": (C.c_int, [_vp, _cp, _cp]),
    "nm_engine_clear": (C.c_int, [_vp, C.POINTER(_u64)]),
    "nm_engine_batch_delete_embeddings": (C.c_int, [_vp, _cp, C.POINTER(_u64)]),
    "nm_engine_search_paginated": (C.c_int, [_vp, C.c_int, _vp, _sz, _sz, _sz, C.c_int64, C.c_int, _pvp,
                                             C.POINTER(_u64), C.POINTER(C.c_int)]),
    "nm_engine_debug_filter_program": (C.c_int, [_vp, C.c_uint32, _cp, _vp, _sz, C.POINTER(_sz)]),
    "nm_engine_query_points": (C.c_int, [_vp, _cp, _vp, _sz, _sz, _sz, C.c_int, C.c_float, _pvp]),
    "nm_engine_set_entity_embedding": (C.c_int, [_vp, _cp, _vp, _sz]),
    "nm_engine_remove_entity_embedding": (C.c_int, [_vp, _cp]),
    "nm_engine_entity_has_embedding": (C.c_int, [_vp, _cp]),
    "nm_engine_search_entities": (C.c_int, [_vp, _vp, _sz, _sz, _pvp]),
    "nm_engine_execute": (C.c_int, [_vp, _cp, _pvp]),
    "nm_engine_execute_parsed": (C.c_int, [_vp, _cp, _pvp]),
    "nm_engine_mirror_rows": (C.c_int, [_vp, C.c_uint32, C.POINTER(_u64), C.POINTER(_u64)]),
    "nm_results_len": (_sz, [_vp]),
    "nm_results_key": (_cp, [_vp, _sz]),
    "nm_results_score": (C.c_float, [_vp, _sz]),
    "nm_results_free": (None, [_vp]),
}

_bound = False


def _lib():
    global _bound
    l = _ffi.lib()
    if not _bound:
        for name, (res, args) in ENGINE_SIGNATURES.items():
            fn = getattr(l, name)
            fn.restype, fn.argtypes = res, args
        _bound = True
    return l


_KINDS = {1: "EmptyVector", 2: "InvalidTopK", 3: "DimensionMismatch", 4: "StorageError",
          5: "SearchTimeout", 6: "InvalidArgument", 7: "NotFound", 8: "ConfigurationError",
          9: "CollectionExists", 10: "CollectionNotFound"}


class VectorError(Exception):
    """VectorError (vector_engine/src/lib.rs:102-149); `.kind` is the variant name."""

    def __init__(self, code: int, msg: str):
        super().__init__(msg)
        self.code = code
        self.kind = _KINDS.get(code, "Unknown")


def _check(code: int) -> None:
    if code != 0:
        raise VectorError(code, _lib().nm_engine_last_error().decode("utf-8", "replace"))


@dataclass
class SearchResult:
    key: str
    score: float


def _f32(v) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(v, dtype=np.float32).reshape(-1))


def _take(handle: C.c_void_p) -> list[SearchResult]:
    l = _lib()
    if not handle:
        return []
    try:
        n = l.nm_results_len(handle)
        return [SearchResult(l.nm_results_key(handle, i).decode(), float(np.float32(l.nm_results_score(handle, i))))
                for i in range(n)]
    finally:
        l.nm_results_free(handle)


class VectorEngine:
    def __init__(self, *, sparse_threshold: float | None = None, parallel_threshold: int | None = None,
                 max_dimension: int | None = None, search_timeout_ms: int | None = None,
                 devices: list[int] | None = None, device_prefilter: bool = False,
                 max_keys_per_scan: int | None = None):
        l = _lib()
        cfg = NmEngineConfig()
        l.nm_engine_config_default(C.byref(cfg))
        cfg.device_prefilter = 1 if device_prefilter else 0
        if sparse_threshold is not None:
            cfg.sparse_threshold = sparse_threshold
        if parallel_threshold is not None:
            cfg.parallel_threshold = parallel_threshold
        if max_dimension is not None:
            cfg.max_dimension = max_dimension
        if max_keys_per_scan is not None:
            cfg.max_keys_per_scan = max_keys_per_scan
        if search_timeout_ms is not None:
            cfg.search_timeout_ms = search_timeout_ms
        if devices:
            cfg.n_devices = len(devices)
            for i, d in enumerate(devices[:8]):
                cfg.devices[i] = d
        self._h = C.c_void_p()
        _check(l.nm_engine_create(C.byref(cfg), C.byref(self._h)))

    def close(self):
        if self._h:
            _lib().nm_engine_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- store side ----
    def store_embedding(self, key: str, vector) -> None:
        v = _f32(vector)
        _check(_lib().nm_engine_store_embedding(self._h, key.encode(), v.ctypes.data, v.size))

    def get_embedding(self, key: str) -> np.ndarray:
        n = C.c_size_t()
        _check(_lib().nm_engine_get_embedding(self._h, key.encode(), None, 0, C.byref(n)))
        out = np.empty(n.value, np.float32)
        _check(_lib().nm_engine_get_embedding(self._h, key.encode(), out.ctypes.data, out.size, C.byref(n)))
        return out

    def delete_embedding(self, key: str) -> None:
        _check(_lib().nm_engine_delete_embedding(self._h, key.encode()))

    def exists(self, key: str) -> bool:
        return bool(_lib().nm_engine_exists(self._h, key.encode()))

    def count(self) -> int:
        return int(_lib().nm_engine_count(self._h))

    # ---- search side ----
    def search_similar_batch(self, queries, top_k: int, metric: int = 0) -> list[list[SearchResult]]:
        """queries: [nq, dim]; element i == search_similar_with_metric(queries[i], top_k, metric)."""
        q = np.ascontiguousarray(np.asarray(queries, dtype=np.float32))
        if q.ndim != 2:
            raise ValueError("queries must be [nq, dim]")
        nq = q.shape[0]
        hs = (C.c_void_p * max(nq, 1))()
        _check(_lib().nm_engine_search_similar_batch(self._h, q.ctypes.data, nq, q.shape[1], top_k,
                                                     int(metric), hs))
        return [_take(C.c_void_p(hs[i])) for i in range(nq)]

    def search_similar(self, query, top_k: int) -> list[SearchResult]:
        q = _f32(query)
        h = C.c_void_p()
        _check(_lib().nm_engine_search_similar(self._h, q.ctypes.data, q.size, top_k, C.byref(h)))
        return _take(h)

    def search_similar_with_metric(self, query, top_k: int, metric: int) -> list[SearchResult]:
        q = _f32(query)
        h = C.c_void_p()
        _check(_lib().nm_engine_search_similar_with_metric(self._h, q.ctypes.data, q.size, top_k,
                                                           metric, C.byref(h)))
        return _take(h)

    @staticmethod
    def compute_similarity(a, b) -> float:
        a, b = _f32(a), _f32(b)
        out = C.c_float()
        _check(_lib().nm_engine_compute_similarity(a.ctypes.data, a.size, b.ctypes.data, b.size,
                                                   C.byref(out)))
        return float(np.float32(out.value))

    # ---- collections ----
    def create_collection(self, name: str, dimension: int | None = None, metric: int = COSINE):
        _check(_lib().nm_engine_create_collection(self._h, name.encode(), dimension or 0, metric))

    def delete_collection(self, name: str):
        _check(_lib().nm_engine_delete_collection(self._h, name.encode()))

    def collection_exists(self, name: str) -> bool:
        return bool(_lib().nm_engine_collection_exists(self._h, name.encode()))

    def store_in_collection(self, collection: str, key: str, vector):
        v = _f32(vector)
        _check(_lib().nm_engine_store_in_collection(self._h, collection.encode(), key.encode(),
                                                    v.ctypes.data, v.size))

    def delete_from_collection(self, collection: str, key: str):
        _check(_lib().nm_engine_delete_from_collection(self._h, collection.encode(), key.encode()))

    def collection_count(self, collection: str) -> int:
        return int(_lib().nm_engine_collection_count(self._h, collection.encode()))

    def search_in_collection(self, collection: str, query, top_k: int) -> list[SearchResult]:
        q = _f32(query)
        h = C.c_void_p()
        _check(_lib().nm_engine_search_in_collection(self._h, collection.encode(), q.ctypes.data,
                                                     q.size, top_k, C.byref(h)))
        return _take(h)

    # ---- metadata + filtered search ----
    def store_embedding_with_metadata(self, key: str, vector, metadata: dict) -> None:
        v = _f32(vector)
        _check(_lib().nm_engine_store_embedding_with_metadata(self._h, key.encode(), v.ctypes.data,
                                                              v.size, encode_metadata(metadata)))

    def store_in_collection_with_metadata(self, collection: str, key: str, vector, metadata: dict):
        v = _f32(vector)
        _check(_lib().nm_engine_store_in_collection_with_metadata(
            self._h, collection.encode(), key.encode(), v.ctypes.data, v.size, encode_metadata(metadata)))

    def search_similar_filtered(self, query, top_k: int, where: str, strategy: int = AUTO,
                                oversample_factor: int = 0) -> list[SearchResult]:
        q = _f32(query)
        h = C.c_void_p()
        _check(_lib().nm_engine_search_similar_filtered(self._h, q.ctypes.data, q.size, top_k,
                                                        where.encode(), strategy, oversample_factor,
                                                        C.byref(h)))
        return _take(h)

    def search_filtered_in_collection(self, collection: str, query, top_k: int, where: str,
                                      strategy: int = AUTO, oversample_factor: int = 0):
        q = _f32(query)
        h = C.c_void_p()
        _check(_lib().nm_engine_search_filtered_in_collection(
            self._h, collection.encode(), q.ctypes.data, q.size, top_k, where.encode(), strategy,
            oversample_factor, C.byref(h)))
        return _take(h)

    def update_metadata(self, key: str, metadata: dict) -> None:
        _check(_lib().nm_engine_update_metadata(self._h, key.encode(), encode_metadata(metadata)))

    def remove_metadata_field(self, key: str, field: str) -> None:
        _check(_lib().nm_engine_remove_metadata_field(self._h, key.encode(), field.encode()))

    def has_metadata_field(self, key: str, field: str) -> bool:
        return bool(_lib().nm_engine_has_metadata_field(self._h, key.encode(), field.encode()))

    def clear(self) -> int:
        n = _u64(0)
        _check(_lib().nm_engine_clear(self._h, C.byref(n)))
        return int(n.value)

    def batch_delete_embeddings(self, keys) -> int:
        n = _u64(0)
        _check(_lib().nm_engine_batch_delete_embeddings(self._h, "\x1f".join(keys).encode(), C.byref(n)))
        return int(n.value)

    def search_paginated(self, query, top_k: int, skip: int = 0, limit: int | None = None,
                         count_total: bool = False, entities: bool = False):
        """-> (items, total_count or None, has_more): search_similar_paginated / search_entities_paginated."""
        q = _f32(query)
        h = C.c_void_p()
        total = _u64(0)
        more = C.c_int(0)
        _check(_lib().nm_engine_search_paginated(self._h, 1 if entities else 0, q.ctypes.data, q.size, top_k,
                                                 skip, -1 if limit is None else limit, 1 if count_total else 0,
                                                 C.byref(h), C.byref(total), C.byref(more)))
        items = _take(h)
        return items, (None if total.value == 0xFFFFFFFFFFFFFFFF else int(total.value)), bool(more.value)

    def debug_filter_program(self, dim: int, where: str) -> dict:
        """Columns + compiled postfix program + host verdict per row (JSON; no device involved)."""
        import json
        n = _sz(0)
        _check(_lib().nm_engine_debug_filter_program(self._h, dim, where.encode(), None, 0, C.byref(n)))
        buf = C.create_string_buffer(n.value + 1)
        _check(_lib().nm_engine_debug_filter_program(self._h, dim, where.encode(), buf, n.value + 1, C.byref(n)))
        return json.loads(buf.value.decode())

    def count_matching(self, where: str) -> int:
        n = C.c_uint64()
        _check(_lib().nm_engine_count_matching(self._h, where.encode(), C.byref(n)))
        return int(n.value)

    def query_points(self, collection: str, vector, limit: int, offset: int = 0,
                     score_threshold: float | None = None) -> list[SearchResult]:
        q = _f32(vector)
        h = C.c_void_p()
        _check(_lib().nm_engine_query_points(self._h, collection.encode(), q.ctypes.data, q.size,
                                             limit, offset, 0 if score_threshold is None else 1,
                                             score_threshold or 0.0, C.byref(h)))
        return _take(h)

    # ---- unified entity mode ----
    def set_entity_embedding(self, entity_key: str, vector) -> None:
        v = _f32(vector)
        _check(_lib().nm_engine_set_entity_embedding(self._h, entity_key.encode(), v.ctypes.data, v.size))

    def remove_entity_embedding(self, entity_key: str) -> None:
        _check(_lib().nm_engine_remove_entity_embedding(self._h, entity_key.encode()))

    def entity_has_embedding(self, entity_key: str) -> bool:
        return bool(_lib().nm_engine_entity_has_embedding(self._h, entity_key.encode()))

    def search_entities(self, query, top_k: int) -> list[SearchResult]:
        q = _f32(query)
        h = C.c_void_p()
        _check(_lib().nm_engine_search_entities(self._h, q.ctypes.data, q.size, top_k, C.byref(h)))
        return _take(h)

    # ---- router ----
    def execute(self, command: str) -> list[SearchResult] | None:
        """QueryRouter::execute (legacy string path).  None == QueryResult::Empty."""
        h = C.c_void_p()
        _check(_lib().nm_engine_execute(self._h, command.encode(), C.byref(h)))
        return _take(h) if h else None

    def execute_parsed(self, command: str) -> list[SearchResult] | None:
        h = C.c_void_p()
        _check(_lib().nm_engine_execute_parsed(self._h, command.encode(), C.byref(h)))
        return _take(h) if h else None

    def mirror_rows(self, dim: int) -> tuple[int, int]:
        a, b = C.c_uint64(), C.c_uint64()
        _check(_lib().nm_engine_mirror_rows(self._h, dim, C.byref(a), C.byref(b)))
        return int(a.value), int(b.value)
