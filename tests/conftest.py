import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")
if GOLDEN not in sys.path:
    sys.path.insert(0, GOLDEN)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session", autouse=True)
def _built_library():
    """The GPU tier needs the in-tree libslicq.so; build it (nvcc, sm_100a) if a fresh checkout lacks it."""
    lib = os.path.join(ROOT, "xumx_slicq_b200", "libslicq.so")
    if not os.path.exists(lib):
        import __graft_entry__
        __graft_entry__.build()
    yield
