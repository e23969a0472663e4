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


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
