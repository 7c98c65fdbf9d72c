import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session", autouse=True)
def _native_artifacts():
    """Built files are git-ignored: (re)build liborbx.so / the C++ front end in-tree when they are missing or older than
    their sources (nvcc and g++ exist both in the build container and on the GPU box).  A failed build fails the tests:
    there is no fallback path to hide behind."""
    from orbslam_in_practice_b200 import build as b
    b.build()
    b.build_cpp()


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
