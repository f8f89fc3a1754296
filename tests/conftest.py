import os
import sys

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(__file__))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_sessionstart(session):
    """The shared libraries are build artefacts (git-ignored): build whatever is missing so that a fresh checkout can run
    the suite directly.  oracle/c takes a second; the CUDA library takes a few minutes (nvcc cross-compiles without a GPU)
    and is normally already there (`__graft_entry__.build()`)."""
    import shutil
    import subprocess

    if not os.path.exists(os.path.join(ROOT, "oracle", "c", "libpmpc_oracle.so")) and shutil.which("gcc"):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle", "c")], check=False, capture_output=True)
    if not os.path.exists(os.path.join(ROOT, "pmpc_b200", "libpmpc_b200.so")) and shutil.which("nvcc"):
        subprocess.run(["make", "-C", os.path.join(ROOT, "pmpc_b200", "csrc"), "-j", str(min(8, os.cpu_count() or 1))],
                       check=False, capture_output=True)


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
