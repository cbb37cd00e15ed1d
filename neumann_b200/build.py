"""Builds the in-tree native libraries.

`libneumann_b200.so` is the product: hand-written sm_100a CUDA kernels + the C ABI declared in
include/neumann_b200.h + the C++ host mirror of the reference's VectorEngine / SIMILAR operator.
It is compiled with nvcc for sm_100a only (no multi-arch fatbin, no PTX fallback for older
parts).  The built .so stays in-tree (git-ignored) so it travels to the GPU box.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
CSRC = ROOT / "neumann_b200" / "csrc"
LIB = ROOT / "neumann_b200" / "libneumann_b200.so"

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")

NVCC_FLAGS = [
    "-std=c++17", "-O3",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    # bit-exact reference arithmetic: no FMA contraction, IEEE div/sqrt, keep denormals
    "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
    "-Xcompiler", "-fPIC,-ffp-contract=off,-fno-fast-math,-Wall",
    "-shared", "-cudart", "static",
    "--threads", "4",
]


def sources() -> list[Path]:
    return sorted(CSRC.glob("*.cu")) + sorted(CSRC.glob("*.cpp"))


def _content_hash(deps: list[Path], flags: list[str]) -> str:
    """sha256 over the build recipe and every source byte: the .so is git-ignored but ships to the
    GPU box, so "is it current" must not depend on mtimes (a checkout resets them)."""
    h = hashlib.sha256()
    h.update("\0".join(flags).encode())
    for d in sorted(deps):
        h.update(str(d.relative_to(ROOT)).encode() + b"\0")
        h.update(d.read_bytes())
        h.update(b"\0")
    return h.hexdigest()


def _stale(target: Path, digest: str) -> bool:
    side = target.with_name(target.name + ".srchash")
    return not target.exists() or not side.exists() or side.read_text().strip() != digest


def library_deps() -> list[Path]:
    return sources() + sorted(CSRC.glob("*.cuh")) + sorted(CSRC.glob("*.hpp")) + \
        sorted((ROOT / "include").glob("*.h"))


def library_is_current() -> bool:
    return not _stale(LIB, _content_hash(library_deps(), NVCC_FLAGS))


def build_library(force: bool = False, verbose: bool = False) -> Path:
    digest = _content_hash(library_deps(), NVCC_FLAGS)
    if not force and not _stale(LIB, digest):
        return LIB
    cmd = [NVCC, *NVCC_FLAGS, "-I", str(ROOT / "include"), "-o", str(LIB),
           *[str(s) for s in sources()], "-ldl", "-lpthread"]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
        print(" ".join(cmd), file=sys.stderr)
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed building libneumann_b200.so")
    LIB.with_name(LIB.name + ".srchash").write_text(digest + "\n")
    return LIB


def build_oracle(force: bool = False) -> Path:
    """Test infrastructure only (see oracle/nm_oracle.c)."""
    odir = ROOT / "oracle"
    lib = odir / "libnm_oracle.so"
    digest = _content_hash([odir / "nm_oracle.c", odir / "Makefile"], [])
    if force or _stale(lib, digest):
        subprocess.run(["make", "-C", str(odir), "-B"], check=True, capture_output=True)
        lib.with_name(lib.name + ".srchash").write_text(digest + "\n")
    return lib


if __name__ == "__main__":
    build_library(force="--force" in sys.argv, verbose=True)
    build_oracle(force="--force" in sys.argv)
    print(LIB)
