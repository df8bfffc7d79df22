import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu under gpurun)")


@pytest.fixture(scope="session")
def lib_built():
    """Build (or reuse) the sm_100a library; nvcc cross-compiles on a GPU-less box."""
    from openwork_b200 import build
    return build.build()
