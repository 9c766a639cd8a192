import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def c_oracle_lib():
    from oracle import c_oracle
    c_oracle.build()
    return c_oracle.lib()


@pytest.fixture(scope="session")
def bp4_lib():
    """libbp4.so must exist (built in-tree by __graft_entry__.build / build.py)."""
    from mf_data_locality_b200 import capi
    if not os.path.exists(capi.LIB_PATH):
        from mf_data_locality_b200 import build
        build.build_cuda()
    return capi.lib()
