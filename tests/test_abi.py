"""CPU: the C-ABI library loads, exports every symbol include/afv_b200.h declares, and fails loudly without a GPU
(no compute calls here)."""
from __future__ import annotations

import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]


def _header_symbols() -> list[str]:
    txt = (ROOT / "include" / "afv_b200.h").read_text()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(afv_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    from pyafv_b200 import _build, _lib
    _build.build()
    lib = C.CDLL(str(_lib.LIB_PATH))
    declared = _header_symbols()
    assert len(declared) >= 35
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/afv_b200.h but not exported"
    assert sorted(_lib.SYMBOLS) == declared, "pyafv_b200/_lib.py binds a different set than the header declares"
    lib.afv_build_info.restype = C.c_char_p
    assert b"sm_100a" in lib.afv_build_info()


def test_sass_is_sm_100a_only():
    import shutil
    import subprocess
    from pyafv_b200 import _lib
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    out = subprocess.run(["cuobjdump", "-lelf", str(_lib.LIB_PATH)], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_philox_host_stream_is_standard_normal():
    from pyafv_b200 import _lib
    lib = _lib.load()
    x = np.array([lib.afv_philox_normal(C.c_uint64(42), C.c_int64(i), C.c_int64(3)) for i in range(20000)])
    assert abs(x.mean()) < 0.03 and abs(x.std() - 1) < 0.03
    assert lib.afv_philox_normal(C.c_uint64(42), C.c_int64(7), C.c_int64(3)) == x[7]
    assert lib.afv_philox_normal(C.c_uint64(43), C.c_int64(7), C.c_int64(3)) != x[7]


def test_no_cpu_fallback():
    """Without a usable GPU the product path must raise, not fall back to anything."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import pyafv_b200 as afv
    from pyafv_b200._lib import AfvError
    with pytest.raises(AfvError, match="no CPU fallback"):
        afv.FiniteVoronoiSimulator(np.zeros((4, 2)), afv.PhysicalParams())


def test_product_package_never_touches_the_oracle():
    for f in (ROOT / "pyafv_b200").rglob("*"):
        if f.suffix in (".py", ".cu", ".cuh", ".inc", ".h"):
            txt = f.read_text()
            assert "afv_oracle" not in txt and "oracle/" not in txt and "build_ref" not in txt, f
