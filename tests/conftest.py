import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for pth in (ROOT, os.path.join(ROOT, "tests")):
    if pth not in sys.path:
        sys.path.insert(0, pth)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def _have_gpu() -> bool:
    try:
        import ctypes as C
        import sabc_b200
        n = C.c_int(0)
        return sabc_b200._lib.lib().sabc_device_count(C.byref(n)) == 0 and n.value > 0
    except Exception:
        return False


@pytest.fixture(scope="session")
def gpu():
    if not _have_gpu():
        pytest.fail("GPU test selected but no CUDA device / libsabc_b200.so available (no CPU fallback exists)")
    return True
