"""CPU: the C-ABI shared library loads and exports exactly what include/orbx.h declares.
No compute calls here (no GPU in this container); compute parity lives in the -m gpu tests."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    txt = open(os.path.join(ROOT, "include", "orbx.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(orb[xm]_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    from orbslam_in_practice_b200 import _lib
    L = _lib.load()
    declared = _header_symbols()
    assert declared, "no declarations parsed"
    missing = [s for s in declared if not hasattr(L, s)]
    assert not missing, "liborbx.so lacks %s" % missing
    assert sorted(_lib.ABI_SYMBOLS) == declared, "ABI_SYMBOLS out of sync with include/orbx.h"


def test_no_cpu_fallback_without_device():
    """On a box without an sm_100 device every entry point must refuse loudly (ORBX_E_NODEVICE)."""
    from orbslam_in_practice_b200 import _lib
    L = _lib.load()
    if L.orbx_device_count() > 0:
        pytest.skip("a B200 is visible; the refusal path is exercised on CPU-only boxes")
    with pytest.raises(_lib.OrbxError, match="no sm_100"):
        _lib.Extractor()
    with pytest.raises(_lib.OrbxError, match="no sm_100"):
        _lib.Matcher(16, 16)


def test_strerror_and_version():
    from orbslam_in_practice_b200 import _lib
    L = _lib.load()
    assert L.orbx_version() >= 100
    assert L.orbx_strerror(0) == b"ok" and b"capacity" in L.orbx_strerror(-4)


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under the product package may reference it."""
    pkg = os.path.join(ROOT, "orbslam_in_practice_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, fn), errors="ignore").read()
                assert "oracle" not in txt.lower().replace("ncu, round", ""), "%s mentions the oracle" % fn
