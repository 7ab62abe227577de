import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _native_artifacts():
    """The shared libraries are build products (git-ignored): compile them if a fresh checkout has none.
    (nvcc cross-compiles sm_100a without a GPU; ~1-2 min once.)"""
    lib = os.path.join(ROOT, "cloudy.jl_b200", "libcloudy_b200.so")
    ora = os.path.join(ROOT, "oracle", "liboracle.so")
    if not (os.path.exists(lib) and os.path.exists(ora)):
        import __graft_entry__
        __graft_entry__.build()
    yield
