import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def built_lib():
    """Path of libpmk_b200.so, building it if absent (nvcc cross-compiles without a GPU)."""
    from patchmixturekriging_b200 import LIB_PATH
    if not os.path.exists(LIB_PATH):
        from patchmixturekriging_b200.build import build
        build()
    return LIB_PATH
