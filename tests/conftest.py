import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    # build (or reuse) the native libraries once per session
    from neumann_b200 import build
    build.build_library()
    build.build_oracle()


def _gpu_count() -> int:
    from neumann_b200 import device_count
    return device_count()


@pytest.fixture(scope="session")
def gpu_count():
    return _gpu_count()


def pytest_collection_modifyitems(config, items):
    # `-m gpu` on a box without a GPU is a configuration error, not a skip: the product has no
    # CPU fallback.  Only multi-GPU tests may skip (when fewer devices are visible).
    pass
