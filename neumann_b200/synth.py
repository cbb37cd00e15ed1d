"""Host-side (numpy) twin of the device synthetic-corpus generator (nm_index_fill_synthetic):
x[r,c] = u24(splitmix64(splitmix64(seed) ^ ((row_offset+r)*dim + c))) * 2^-23 - 1.
Used by bench.py to make queries; bit-identical to the CUDA kernel by construction (integer
hash, then two exact f32 operations)."""
from __future__ import annotations

import numpy as np

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix64(x: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        z = (x + np.uint64(0x9E3779B97F4A7C15)) & _M64
        z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
        z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
        return z ^ (z >> np.uint64(31))


def synth_rows(n: int, dim: int, seed: int, row_offset: int = 0) -> np.ndarray:
    s = _splitmix64(np.array([seed], dtype=np.uint64))[0]
    flat = (np.arange(n * dim, dtype=np.uint64) + np.uint64(row_offset * dim))
    u24 = (_splitmix64(s ^ flat) >> np.uint64(40)).astype(np.float32)
    return (u24 * np.float32(2.0 ** -23) - np.float32(1.0)).astype(np.float32).reshape(n, dim)
