"""ctypes front-end of the CPU oracle (oracle/afv_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs may
import this module.  The product package ``pyafv_b200`` never does.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
LIB = HERE / "_build" / "libafv_oracle.so"


class _Params(C.Structure):
    _fields_ = [(k, C.c_double) for k in ("r", "P0", "KA", "KP", "Lambda", "delta")]


def build_lib(force: bool = False) -> Path:
    src = HERE / "afv_oracle.c"
    if force or not LIB.exists() or LIB.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-s", "-C", str(HERE)] + (["-B"] if force else []), check=True)
    return LIB


_lib = None


def _load():
    global _lib
    if _lib is None:
        build_lib()
        lib = C.CDLL(str(LIB))
        lib.afv_oracle_create.restype = C.c_void_p
        lib.afv_oracle_destroy.argtypes = [C.c_void_p]
        lib.afv_oracle_last_error.restype = C.c_char_p
        lib.afv_oracle_last_error.argtypes = [C.c_void_p]
        lib.afv_oracle_build.restype = C.c_int
        lib.afv_oracle_build.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.POINTER(_Params), C.c_double,
                                         C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.afv_oracle_ring_total.restype = C.c_int64
        lib.afv_oracle_ring_total.argtypes = [C.c_void_p]
        lib.afv_oracle_get_rings.argtypes = [C.c_void_p] * 6
        lib.afv_oracle_n_connections.restype = C.c_int64
        lib.afv_oracle_n_connections.argtypes = [C.c_void_p]
        lib.afv_oracle_get_connections.argtypes = [C.c_void_p, C.c_void_p]
        _lib = lib
    return _lib


def oracle_build(pts, *, r=1.0, A0=np.pi, P0=4.8, KA=1.0, KP=1.0, Lambda=0.2, delta=None, periodic=None,
                 rings=False) -> dict:
    """Forces and geometry of the finite-Voronoi tessellation of ``pts`` (N,2), CPU oracle.

    Returns the keys of the reference's ``build()`` dict (finite_voronoi.py:935-945) that are
    implementation-independent: forces, areas, perimeters, arclens, coord_nums, connections; with
    ``rings=True`` also ring_off / ring_type / ring_nb_in / ring_nb_out / ring_xy (vertex positions relative
    to the owning cell, clockwise).
    """
    lib = _load()
    pts = np.ascontiguousarray(pts, dtype=np.float64)
    N = pts.shape[0]
    a0 = np.ascontiguousarray(np.broadcast_to(np.asarray(A0, dtype=np.float64), (N,)))
    if delta is None:
        delta = 0.45 * r
    prm = _Params(r, P0, KA, KP, Lambda, delta)
    out = dict(forces=np.zeros((N, 2)), areas=np.zeros(N), perimeters=np.zeros(N), arclens=np.zeros(N),
               coord_nums=np.zeros(N, dtype=np.int64))
    h = lib.afv_oracle_create()
    try:
        rc = lib.afv_oracle_build(h, N, pts.ctypes.data, a0.ctypes.data, C.byref(prm), float(periodic or 0.0),
                                  out["forces"].ctypes.data, out["areas"].ctypes.data, out["perimeters"].ctypes.data,
                                  out["arclens"].ctypes.data, out["coord_nums"].ctypes.data)
        if rc != 0:
            raise RuntimeError(f"oracle failed ({rc}): {lib.afv_oracle_last_error(h).decode()}")
        K = lib.afv_oracle_n_connections(h)
        conn = np.zeros((K, 2), dtype=np.int64)
        lib.afv_oracle_get_connections(h, conn.ctypes.data)
        out["connections"] = conn
        if rings:
            T = lib.afv_oracle_ring_total(h)
            out["ring_off"] = np.zeros(N + 1, dtype=np.int64)
            out["ring_type"] = np.zeros(T, dtype=np.int8)
            out["ring_nb_in"] = np.zeros(T, dtype=np.int64)
            out["ring_nb_out"] = np.zeros(T, dtype=np.int64)
            out["ring_xy"] = np.zeros((T, 2))
            lib.afv_oracle_get_rings(h, out["ring_off"].ctypes.data, out["ring_type"].ctypes.data,
                                     out["ring_nb_in"].ctypes.data, out["ring_nb_out"].ctypes.data,
                                     out["ring_xy"].ctypes.data)
    finally:
        lib.afv_oracle_destroy(h)
    return out
