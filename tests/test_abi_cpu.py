"""CPU-side tests of the drop-in boundary: the library loads, exports every symbol the headers
declare, and fails loudly (no CPU fallback) when asked to compute without a GPU."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

from neumann_b200 import _ffi, engine as eng

ROOT = Path(__file__).resolve().parent.parent


def declared_functions(header: Path) -> set[str]:
    text = re.sub(r"/\*.*?\*/", "", header.read_text(), flags=re.S)
    text = re.sub(r"//[^\n]*", "", text)
    return set(re.findall(r"\b(nm_[a-z0-9_]+)\s*\(", text))


def test_library_exports_every_declared_symbol():
    lib = _ffi.lib()
    for header in sorted((ROOT / "include").glob("*.h")):
        names = declared_functions(header)
        assert names, header
        for n in names:
            assert hasattr(lib, n), f"{n} declared in {header.name} but not exported"


def test_ctypes_signatures_cover_the_headers():
    assert declared_functions(ROOT / "include" / "neumann_b200.h") == set(_ffi.SIGNATURES)
    assert declared_functions(ROOT / "include" / "neumann_b200_engine.h") == set(eng.ENGINE_SIGNATURES)


def test_abi_version_and_device_count():
    lib = _ffi.lib()
    assert lib.nm_abi_version() == 1
    assert lib.nm_device_count() >= 0


def test_oracle_is_not_linked_into_the_product():
    """The product library must not depend on, or contain, the oracle."""
    import subprocess
    so = ROOT / "neumann_b200" / "libneumann_b200.so"
    deps = subprocess.run(["ldd", str(so)], capture_output=True, text=True).stdout
    assert "nm_oracle" not in deps
    syms = subprocess.run(["nm", "-D", str(so)], capture_output=True, text=True).stdout
    assert "nmo_" not in syms
    for src in (ROOT / "neumann_b200").rglob("*"):
        if src.suffix in {".cu", ".cuh", ".cpp", ".hpp", ".py"}:
            assert "nm_oracle" not in src.read_text() or src.name == "build.py", src


def test_product_sass_uses_tma_and_no_fma_in_lane_arithmetic():
    """Cheap static evidence (cuobjdump works without a GPU): the scan kernels stream with TMA
    (UTMALDG) and accumulate with separate FMUL/FADD; the few FFMA left belong to the IEEE
    sqrt/div sequences only."""
    import shutil, subprocess
    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not on PATH")
    so = ROOT / "neumann_b200" / "libneumann_b200.so"
    sass = subprocess.run(["cuobjdump", "-sass", str(so)], capture_output=True, text=True).stdout
    assert "UTMALDG" in sass
    assert "sm_100a" in sass or "SM100a" in sass.upper() or "EF_CUDA_SM100" in sass
    chunks = sass.split("Function : ")
    scans = [c for c in chunks if "scan_topk_kernel" in c.split("\n", 1)[0]]
    assert len(scans) == 3
    for c in scans:
        assert c.count("FMUL") > 30 and c.count("FADD") > 30
        assert c.count("FFMA") < 40  # sqrt/div expansions only


def test_product_sass_tensor_core_prefilter():
    """Static evidence for the batch pre-filter: the GEMM is tcgen05 (UTCIMMA = kind::i8) on CTA
    pairs with TMA-staged operands, TMEM reads in the epilogue and multicast commits; the exact
    re-score kernels accumulate with separate FMUL/FADD like the scan."""
    import shutil, subprocess
    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not on PATH")
    so = ROOT / "neumann_b200" / "libneumann_b200.so"
    sass = subprocess.run(["cuobjdump", "-sass", str(so)], capture_output=True, text=True).stdout
    chunks = sass.split("Function : ")
    gemm = [c for c in chunks if "tc_gemm_filter_kernel" in c.split("\n", 1)[0]]
    assert len(gemm) == 2                                   # <1> and <2>
    pair = [c for c in gemm if "UTCIMMA.2CTA" in c]
    assert len(pair) == 1
    for mnemonic in ("UTCIMMA.2CTA", "UTMALDG.2D.2CTA", "UTCBAR.2CTA.MULTICAST", "LDTM.x16",
                     "UCGABAR_ARV"):
        assert mnemonic in pair[0], mnemonic
    single = [c for c in gemm if c is not pair[0]][0]
    assert "UTCIMMA" in single and "UTMALDG.2D" in single and "LDTM.x16" in single
    for name in ("tc_score_sorted_kernel", "tc_refine_kernel"):
        ks = [c for c in chunks if name in c.split("\n", 1)[0]]
        assert len(ks) == 1
        assert ks[0].count("FMUL") > 30 and ks[0].count("FADD") > 30
        assert "HMMA" not in ks[0]
    assert "HMMA" not in sass and "HGMMA" not in sass      # no legacy tensor path anywhere


@pytest.mark.skipif(_ffi.lib().nm_device_count() > 0, reason="CPU-only behaviour")
def test_no_gpu_fails_loudly():
    lib = _ffi.lib()
    h = C.c_void_p()
    rc = lib.nm_index_create(8, None, 0, C.byref(h))
    assert rc == _ffi.NM_ERR_STORAGE
    assert b"no CUDA device" in lib.nm_last_error()
    e = eng.VectorEngine()
    e.store_embedding("a", [1.0, 0.0])
    with pytest.raises(eng.VectorError) as ei:
        e.search_similar([1.0, 0.0], 1)
    assert ei.value.kind == "StorageError"


def test_create_rejects_bad_arguments():
    lib = _ffi.lib()
    h = C.c_void_p()
    assert lib.nm_index_create(0, None, 0, C.byref(h)) == _ffi.NM_ERR_EMPTY_VECTOR
    assert lib.nm_index_create(8, None, 0, None) == _ffi.NM_ERR_INVALID_ARGUMENT
    assert lib.nm_search(None, None, 1, 1, 0, None, None, None) == _ffi.NM_ERR_INVALID_ARGUMENT


# ---- host mirror of VectorEngine: everything that does not need the device ------------------
def test_engine_store_get_delete_count():
    e = eng.VectorEngine()
    e.store_embedding("a", [1.0, 2.0, 3.0])
    e.store_embedding("b", [4.0, 5.0])
    assert e.count() == 2 and e.exists("a") and not e.exists("zz")
    assert np.array_equal(e.get_embedding("a"), np.array([1, 2, 3], np.float32))
    e.store_embedding("a", [7.0, 8.0])  # overwrite with a different dimension
    assert np.array_equal(e.get_embedding("a"), np.array([7, 8], np.float32))
    assert e.count() == 2
    e.delete_embedding("a")
    assert e.count() == 1 and not e.exists("a")
    with pytest.raises(eng.VectorError) as ei:
        e.delete_embedding("a")
    assert ei.value.kind == "NotFound"
    with pytest.raises(eng.VectorError) as ei:
        e.get_embedding("nope")
    assert ei.value.kind == "NotFound" and "nope" in str(ei.value)


def test_engine_validation_order_and_short_circuits():
    # vector_engine/src/lib.rs:1953-1974, 2057-2068
    e = eng.VectorEngine(max_dimension=4)
    with pytest.raises(eng.VectorError) as ei:
        e.search_similar([], 5)
    assert ei.value.kind == "EmptyVector"
    with pytest.raises(eng.VectorError) as ei:
        e.search_similar([1.0], 0)
    assert ei.value.kind == "InvalidTopK"
    with pytest.raises(eng.VectorError) as ei:
        e.search_similar([1.0] * 5, 3)
    assert ei.value.kind == "DimensionMismatch" and "expected 4, got 5" in str(ei.value)
    with pytest.raises(eng.VectorError) as ei:
        e.store_embedding("k", [1.0] * 5)
    assert ei.value.kind == "DimensionMismatch"
    with pytest.raises(eng.VectorError) as ei:
        e.store_embedding("k", [])
    assert ei.value.kind == "EmptyVector"
    with pytest.raises(eng.VectorError) as ei:
        e.search_similar_with_metric([], 5, eng.COSINE)
    assert ei.value.kind == "EmptyVector"
    with pytest.raises(eng.VectorError) as ei:
        e.search_similar_with_metric([1.0], 0, eng.COSINE)
    assert ei.value.kind == "InvalidTopK"
    # zero query: empty result before any device work (cosine, dot), lib.rs:1970-1974, 2066
    e.store_embedding("a", [1.0, 0.0])
    assert e.search_similar([0.0, 0.0], 5) == []
    assert e.search_similar_with_metric([0.0, 0.0], 5, eng.DOT_PRODUCT) == []
    # no rows of the query's dimension: empty result, no device work
    assert e.search_similar([1.0, 0.0, 0.0], 5) == []
    # the batch form applies the same checks per query, in the same order
    assert e.search_similar_batch(np.zeros((3, 2), np.float32), 5, eng.COSINE) == [[], [], []]
    assert e.search_similar_batch(np.ones((2, 3), np.float32), 5, eng.EUCLIDEAN) == [[], []]
    assert e.search_similar_batch(np.zeros((0, 2), np.float32), 5, eng.COSINE) == []
    with pytest.raises(eng.VectorError) as ei:
        e.search_similar_batch(np.ones((2, 2), np.float32), 0, eng.COSINE)
    assert ei.value.kind == "InvalidTopK"
    with pytest.raises(eng.VectorError) as ei:
        e.search_similar_batch(np.zeros((2, 0), np.float32), 3, eng.COSINE)
    assert ei.value.kind == "EmptyVector"


def test_engine_config_validation():
    with pytest.raises(eng.VectorError) as ei:
        eng.VectorEngine(sparse_threshold=1.5)
    assert ei.value.kind == "ConfigurationError"
    with pytest.raises(eng.VectorError) as ei:
        eng.VectorEngine(parallel_threshold=0)
    assert ei.value.kind == "ConfigurationError"


def test_engine_sparse_roundtrip_drops_negative_zero():
    # vector_engine/src/lib.rs:1876-1885 + tensor_store/src/sparse_vector.rs:212-229
    e = eng.VectorEngine()
    e.store_embedding("s", [-0.0, 0.0, 0.0, 5.0])  # 75 % zeros -> sparse storage
    assert np.array_equal(e.get_embedding("s").view(np.uint32),
                          np.array([0.0, 0.0, 0.0, 5.0], np.float32).view(np.uint32))
    e.store_embedding("d", [-0.0, 1.0, 2.0, 3.0])  # dense storage keeps -0.0
    assert e.get_embedding("d").view(np.uint32)[0] == 0x80000000


def test_compute_similarity_matches_reference_rules():
    assert abs(eng.VectorEngine.compute_similarity([1, 0], [1, 1]) - 2 ** 0.5 / 2) < 1e-6
    assert eng.VectorEngine.compute_similarity([0, 0], [1, 0]) == 0.0
    with pytest.raises(eng.VectorError) as ei:
        eng.VectorEngine.compute_similarity([1, 2], [1, 2, 3])
    assert ei.value.kind == "DimensionMismatch"
    with pytest.raises(eng.VectorError) as ei:
        eng.VectorEngine.compute_similarity([], [1])
    assert ei.value.kind == "EmptyVector"


def test_host_simd_bits_match_oracle():
    import oracle_ffi as o
    rng = np.random.default_rng(1)
    for dim in (1, 7, 8, 33, 768):
        a = rng.standard_normal(dim).astype(np.float32)
        b = rng.standard_normal(dim).astype(np.float32)
        got = np.float32(eng.VectorEngine.compute_similarity(a, b))
        assert got.view(np.uint32) == o.compute_similarity(a, b).view(np.uint32)


def test_collections_host_side():
    e = eng.VectorEngine()
    e.create_collection("docs", dimension=3, metric=eng.EUCLIDEAN)
    with pytest.raises(eng.VectorError) as ei:
        e.create_collection("docs")
    assert ei.value.kind == "CollectionExists"
    with pytest.raises(eng.VectorError) as ei:
        e.store_in_collection("docs", "k", [1.0, 2.0])
    assert ei.value.kind == "DimensionMismatch"
    e.store_in_collection("docs", "k", [1.0, 2.0, 3.0])
    e.store_in_collection("loose", "k", [1.0])  # no config needed (lib.rs:1445-1500)
    assert e.collection_count("docs") == 1 and e.collection_count("loose") == 1
    assert e.count() == 0  # default space untouched
    with pytest.raises(eng.VectorError) as ei:
        e.search_in_collection("docs", [1.0, 2.0], 3)
    assert ei.value.kind == "DimensionMismatch"
    assert e.search_in_collection("nope", [1.0], 3) == []
    e.delete_collection("docs")
    assert not e.collection_exists("docs") and e.collection_count("docs") == 0
    with pytest.raises(eng.VectorError) as ei:
        e.delete_collection("docs")
    assert ei.value.kind == "CollectionNotFound"


def test_router_parsing_without_device():
    e = eng.VectorEngine()
    assert e.execute("EMBED doc1 [1.0, 0.0, 0.0]") is None
    assert e.execute("embed doc2 0.0, 1.0, 0.0") is None  # brackets optional (QR:6884-6885)
    assert e.count() == 2
    with pytest.raises(eng.VectorError):
        e.execute("EMBED key []")  # QR:8462-8463
    with pytest.raises(eng.VectorError):
        e.execute("SIMILAR")
    with pytest.raises(eng.VectorError):
        e.execute("SIMILAR doc1 TOP abc")
    with pytest.raises(eng.VectorError) as ei:
        e.execute("SIMILAR missing TOP 2")
    assert ei.value.kind == "NotFound"
    with pytest.raises(eng.VectorError):
        e.execute("SELECT * FROM t")
    assert e.execute("SIMILAR [0.0, 0.0, 0.0] TOP 2") == []  # zero query short-circuit
    assert e.execute_parsed("SIMILAR [0.0, 0.0, 0.0] LIMIT 2 COSINE") == []
    with pytest.raises(eng.VectorError):
        e.execute_parsed("SIMILAR 'doc1' LIMIT x")
    with pytest.raises(eng.VectorError):
        e.execute_parsed("SIMILAR 'doc1' CONNECTED TO 'n1' LIMIT 3")
    assert e.execute_parsed("EMBED STORE 'k9' [1.0, 2.0, 3.0] INTO c1") is None
    assert e.collection_count("c1") == 1


def test_entity_embeddings_host_side():
    # vector_engine/src/lib.rs:3060-3145
    e = eng.VectorEngine()
    e.set_entity_embedding("user:1", [1.0, 0.0])
    e.set_entity_embedding("user:2", [0.0, 1.0])
    assert e.entity_has_embedding("user:1") and not e.entity_has_embedding("user:9")
    assert e.count() == 0                      # entity embeddings live outside the emb: space
    e.remove_entity_embedding("user:1")
    assert not e.entity_has_embedding("user:1")
    with pytest.raises(eng.VectorError) as ei:
        e.remove_entity_embedding("user:1")
    assert ei.value.kind == "NotFound"
    with pytest.raises(eng.VectorError) as ei:
        e.set_entity_embedding("user:3", [])
    assert ei.value.kind == "EmptyVector"
    assert e.search_entities([0.0, 0.0], 3) == []
    with pytest.raises(eng.VectorError) as ei:
        e.search_entities([1.0, 0.0], 0)
    assert ei.value.kind == "InvalidTopK"


def test_metadata_maintenance_clear_and_batch_delete_host_side():
    """vector_engine/src/lib.rs:3346-3420 (update_metadata / remove_metadata_field /
    has_metadata_field), :2340-2354 (clear, bounded by max_keys_per_scan), :2924-2940
    (batch_delete_embeddings); reference KATs :5253-5283, :6700-6745.  Host logic only."""
    e = eng.VectorEngine()
    e.store_embedding_with_metadata("item", [1.0, 2.0], {"color": "red"})
    e.update_metadata("item", {"size": "large", "color": "blue"})            # :6700-6736
    assert e.count_matching("color = 'blue' AND size = 'large'") == 1
    assert e.count_matching("color = 'red'") == 0
    assert e.has_metadata_field("item", "size") and not e.has_metadata_field("item", "weight")
    assert not e.has_metadata_field("nobody", "size")
    with pytest.raises(eng.VectorError) as ei:
        e.update_metadata("nonexistent", {})                                  # :6739-6745
    assert ei.value.kind == "NotFound"
    e.remove_metadata_field("item", "size")
    assert not e.has_metadata_field("item", "size") and e.count_matching("EXISTS(size)") == 0
    with pytest.raises(eng.VectorError) as ei:
        e.remove_metadata_field("nonexistent", "size")
    assert ei.value.kind == "NotFound"
    # batch delete: missing keys are skipped (:5253-5283)
    for k, v in (("a", [1.0]), ("b", [2.0]), ("c", [3.0])):
        e.store_embedding(k, v)
    assert e.batch_delete_embeddings(["a", "b"]) == 2
    assert e.count() == 2 and e.exists("c") and e.exists("item")
    assert e.batch_delete_embeddings([]) == 0
    assert e.batch_delete_embeddings(["c", "nonexistent"]) == 1
    # clear: everything at once without a bound ...
    for i in range(7):
        e.store_embedding(f"k{i}", [float(i), 1.0])
    assert e.clear() == 8 and e.count() == 0 and e.clear() == 0
    e.close()
    # ... and at most max_keys_per_scan per call with one ("call again until 0 is returned")
    e = eng.VectorEngine(max_keys_per_scan=3)
    for i in range(8):
        e.store_embedding(f"k{i}", [float(i), 1.0])
    assert [e.clear(), e.clear(), e.clear(), e.clear()] == [3, 3, 2, 0]
    assert e.count() == 0
    e.close()
