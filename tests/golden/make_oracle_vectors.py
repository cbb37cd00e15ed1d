"""Regenerates tests/golden/oracle_vectors.npz: outputs of oracle/nm_oracle.c on seeded
synthetic inputs.  These are NOT reference outputs (the Rust reference cannot run here); they
freeze the oracle so that the CUDA path and any later oracle edit are compared with the same
committed bits.  Run from the repo root: python tests/golden/make_oracle_vectors.py"""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import oracle_ffi as o  # noqa: E402

SEED_ROWS, SEED_QUERY = 0x5EED0001, 0x5EED1001
out = {"seed_rows": np.uint64(SEED_ROWS), "seed_query": np.uint64(SEED_QUERY)}
for name, (n, dim, k) in {"cosine": (20000, 768, 10), "euclidean": (6000, 1536, 100),
                          "dot": (50000, 100, 16)}.items():
    rows = o.fill_synthetic(n, dim, SEED_ROWS)
    q = o.fill_synthetic(1, dim, SEED_QUERY)[0]
    r, s = o.search(rows, q, k, name)
    out[f"{name}_shape"] = np.array([n, dim, k], np.int64)
    out[f"{name}_rows"] = r
    out[f"{name}_score_bits"] = s.view(np.uint32)
np.savez(Path(__file__).parent / "oracle_vectors.npz", **out)
print("wrote oracle_vectors.npz")
