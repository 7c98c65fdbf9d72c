import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.lib()
    return O


@pytest.fixture(scope="session")
def orbx():
    """The product library through its C ABI; fails loudly if it is missing (no fallback)."""
    from orbslam_in_practice_b200 import _lib
    _lib.load()
    if _lib.load().orbx_device_count() < 1:
        pytest.fail("liborbx.so loaded but no sm_100 device is visible: GPU tests cannot run")
    return _lib
