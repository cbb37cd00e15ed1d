"""neumann_b200 — B200-native SIMILAR brute-force scan behind Neumann's vector_engine API.

The product is `libneumann_b200.so` (sm_100a CUDA kernels + C ABI, include/neumann_b200.h).
This package only holds the build script and ctypes harness used by tests/ and bench.py.
"""
from .index import DeviceIndex, NmError, comm_create_id, device_count  # noqa: F401
