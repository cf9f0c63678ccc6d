import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle_backend():
    from oracle import pdlp_oracle
    return pdlp_oracle.backend()


@pytest.fixture(scope="session")
def b200_backend():
    from ortools_b200 import pdlp
    be = pdlp.backend()
    if be.device_count() < 1:
        pytest.fail("libpdlp_b200.so loaded but no CUDA device is usable (there is no CPU fallback)")
    return be


@pytest.fixture(params=["oracle", pytest.param("b200", marks=pytest.mark.gpu)])
def backend(request):
    """Both sides of every parity test: the CPU oracle (runs anywhere) and the
    CUDA product through its C ABI (GPU box only)."""
    return request.getfixturevalue(request.param + "_backend")
