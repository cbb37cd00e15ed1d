"""The C ABI consumed from plain C11 (no ctypes): include/*.h compile as C with -Werror
-pedantic, the program links against libneumann_b200.so and runs create / load / append / search /
masked + filtered search / mutations / destroy.  On a CPU-only host it checks the documented
loud failure instead (exit code 10)."""
import os
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def _build(tmp_path) -> Path:
    exe = tmp_path / "c_consumer"
    cmd = ["gcc", "-std=c11", "-Wall", "-Wextra", "-Werror", "-pedantic", str(ROOT / "tests" / "c_consumer.c"),
           "-I", str(ROOT / "include"), "-L", str(ROOT / "neumann_b200"), "-lneumann_b200", "-lm",
           "-Wl,-rpath," + str(ROOT / "neumann_b200"), "-o", str(exe)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_headers_compile_as_c11_and_fail_loudly_without_a_gpu(tmp_path):
    from neumann_b200 import device_count
    if device_count() > 0:
        pytest.skip("GPU present: covered by the gpu-marked run")
    r = subprocess.run([str(_build(tmp_path))], capture_output=True, text=True)
    assert r.returncode == 10, (r.returncode, r.stdout, r.stderr)


@pytest.mark.gpu
def test_c_consumer_end_to_end(tmp_path):
    r = subprocess.run([str(_build(tmp_path))], capture_output=True, text=True)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    assert "c_consumer: ok" in r.stdout
