import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def built_lib():
    """The in-tree CUDA library; built here if nvcc is present, otherwise it must have travelled."""
    from haploconduct_b200 import build as B, capi

    if os.path.exists(B.NVCC):
        B.build()
    assert os.path.exists(capi.LIB_PATH), "libhc_b200.so missing"
    return capi.lib()
