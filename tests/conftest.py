import os
import sys

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(__file__))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _gpu_count():
    try:
        from pmpc_b200 import _lib

        return _lib.device_count()
    except Exception:
        return 0


@pytest.fixture(scope="session")
def gpu():
    n = _gpu_count()
    if n < 1:
        pytest.fail("test marked gpu but no CUDA device / libpmpc_b200.so: the product has no CPU fallback")
    return n
