"""Backend selector, mirroring ``pyafv/backend.py:3-19`` of the reference.

The reference picks the Cython extension and falls back to pure Python.  This package has exactly one
backend -- the sm_100a CUDA library behind the C-ABI of ``include/afv_b200.h`` -- and NO fallback: if the
library cannot be loaded, ``backend_impl`` is ``None``, ``_IMPORT_ERROR`` holds the reason, and constructing a
simulator raises.
"""
from __future__ import annotations

try:
    from . import _lib as _impl
    _impl.load()
    _BACKEND_NAME = "cuda"
    _IMPORT_ERROR = None
except Exception as e:  # pragma: no cover - exercised only on a box without the built library
    _impl = None
    _BACKEND_NAME = "unavailable"
    _IMPORT_ERROR = e

# ---- for explicit API ----
backend_impl = _impl

__all__ = ["backend_impl", "_BACKEND_NAME", "_IMPORT_ERROR"]
