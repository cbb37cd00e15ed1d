"""tests/cpp/reference_suite.cpp: the reference's own vector_engine unit tests (218 of vector_engine, 34 of query_router, 7 of the gRPC points service: store /
get / delete, search_similar*, metrics, sparse storage, entities, pagination, batch operations,
metadata, filtered search, collections, timeouts, concurrency, edge values) restated against the
C++ host mirror, each under the reference test's name with its lib.rs line.  The tests that never
search run here on the CPU; the whole suite (searches on the device) runs under -m gpu."""
import re
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def _build(tmp_path) -> Path:
    from neumann_b200 import build as nb
    nb.build_library()  # no-op when the in-tree library matches the sources
    exe = tmp_path / "reference_suite"
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-Wno-unused-function",
           str(ROOT / "tests" / "cpp" / "reference_suite.cpp"),
           "-I", str(ROOT / "neumann_b200" / "csrc"), "-I", str(ROOT / "include"),
           "-L", str(ROOT / "neumann_b200"), "-lneumann_b200", "-pthread",
           "-Wl,-rpath," + str(ROOT / "neumann_b200"), "-o", str(exe)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def _summary(stdout: str):
    m = re.search(r"reference_suite: (\d+) tests run, (\d+) skipped \(need a device\), (\d+) failed", stdout)
    assert m, stdout
    return tuple(int(x) for x in m.groups())


def test_reference_suite_host_side(tmp_path):
    r = subprocess.run([str(_build(tmp_path)), "--host"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (r.stdout, r.stderr)
    ran, skipped, failed = _summary(r.stdout)
    assert failed == 0 and ran >= 173 and ran + skipped >= 261


@pytest.mark.parametrize("sanitizer", ["address,undefined", "thread"])
def test_host_mirror_is_clean_under_sanitizers(tmp_path, sanitizer):
    """The host subset again, with the host mirror's own sources (vector_engine.cpp, similar_router.cpp,
    filter.cpp) compiled into the test binary under ASan + UBSan / TSan — the concurrency tests of
    the reference (same-key stores and deletes, concurrent batches) run under ThreadSanitizer."""
    from neumann_b200 import build as nb
    nb.build_library()
    exe = tmp_path / "reference_suite_san"
    csrc = ROOT / "neumann_b200" / "csrc"
    cmd = ["g++", "-std=c++17", "-O1", "-g", f"-fsanitize={sanitizer}", "-fno-omit-frame-pointer",
           str(ROOT / "tests" / "cpp" / "reference_suite.cpp"), str(csrc / "vector_engine.cpp"),
           str(csrc / "similar_router.cpp"), str(csrc / "filter.cpp"),
           "-I", str(csrc), "-I", str(ROOT / "include"), "-I", "/usr/local/cuda/include",
           "-L", str(ROOT / "neumann_b200"), "-lneumann_b200", "-pthread",
           "-Wl,-rpath," + str(ROOT / "neumann_b200"), "-o", str(exe)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0 and ("cannot find -l" in r.stderr or "unrecognized" in r.stderr):
        pytest.skip("no sanitizer runtime in this toolchain: " + r.stderr.strip()[-200:])
    assert r.returncode == 0, r.stderr[-3000:]
    import os
    env = dict(os.environ, ASAN_OPTIONS="detect_leaks=0:abort_on_error=0", UBSAN_OPTIONS="halt_on_error=1",
               TSAN_OPTIONS="halt_on_error=1")
    r = subprocess.run([str(exe), "--host"], capture_output=True, text=True, timeout=900, env=env)
    if r.returncode != 0 and "reference_suite:" not in r.stdout and "FATAL:" in r.stderr:
        pytest.skip("the sanitizer runtime cannot start in this environment: " + r.stderr.strip()[:200])
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-4000:])
    assert "ERROR: AddressSanitizer" not in r.stderr and "runtime error:" not in r.stderr
    assert "WARNING: ThreadSanitizer" not in r.stderr
    ran, skipped, failed = _summary(r.stdout)
    assert failed == 0 and ran >= 173


@pytest.mark.gpu
def test_reference_suite_on_the_device(tmp_path):
    r = subprocess.run([str(_build(tmp_path))], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, (r.stdout, r.stderr)
    ran, skipped, failed = _summary(r.stdout)
    assert failed == 0 and skipped == 0 and ran >= 261
