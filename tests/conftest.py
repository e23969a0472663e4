import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # the native artefacts are built in-tree and git-ignored: build them if this is a fresh checkout
    needed = [os.path.join(ROOT, "panacus_b200", "libpanacus_b200.so"), os.path.join(ROOT, "panacus_b200", "bin", "panacus"),
              os.path.join(ROOT, "oracle", "libpanacus_oracle.so")]
    if not all(os.path.exists(p) for p in needed):
        import __graft_entry__
        __graft_entry__.build()


def _cuda_devices() -> int:
    """Number of usable CUDA devices as the product library sees them (0 when the driver / library is unusable)."""
    try:
        import ctypes as C

        from panacus_b200 import _native
        n = C.c_int(0)
        return n.value if _native.lib().pgx_device_count(C.byref(n)) == 0 else 0
    except Exception:
        return 0


def pytest_collection_modifyitems(config, items):
    # plain `pytest tests` on a machine without a GPU: skip (not fail) everything marked gpu; tests that need
    # two devices carry @pytest.mark.gpu too and additionally check the count themselves
    if not any("gpu" in item.keywords for item in items):
        return
    if _cuda_devices() > 0:
        return
    skip = pytest.mark.skip(reason="no usable CUDA device (pgx_device_count)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def n_cuda_devices():
    return _cuda_devices()


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
