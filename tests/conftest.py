import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Make sure the in-tree shared libraries exist (nvcc cross-compiles without a GPU)."""
    import __graft_entry__
    __graft_entry__.build()


@pytest.fixture(scope="session")
def host_harness():
    import ctypes
    return ctypes.CDLL(os.path.join(ROOT, "tests", "host_harness", "libepnp_host.so"))
