"""ctypes binding of the C-ABI in include/afv_b200.h.  No CPU fallback: loading or device errors raise."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

PKG = Path(__file__).resolve().parent
LIB_PATH = Path(__import__("os").environ.get("AFV_B200_LIB", PKG / "libafv_b200.so"))   # override for A/B builds

AFV_OK, AFV_ERR_CUDA, AFV_ERR_ARG, AFV_ERR_CAPACITY, AFV_ERR_STATE = 0, -1, -2, -3, -4
BUILD_FORCES, BUILD_CONNECT, BUILD_RINGS = 1, 2, 4
F_XY, _F_UNUSED, F_THETA, F_A0, F_GID, F_FX, F_FY, F_AREA, F_PERIM, F_ARCLEN, F_COORD, F_OWNED = range(12)

# every symbol include/afv_b200.h declares (tests/test_abi.py checks the header against this list)
SYMBOLS = [
    "afv_create", "afv_destroy", "afv_last_error", "afv_build_info", "afv_set_params", "afv_set_positions",
    "afv_set_preferred_areas", "afv_set_theta", "afv_num_cells", "afv_build", "afv_step", "afv_philox_normal",
    "afv_get_positions", "afv_get_theta", "afv_get_forces", "afv_get_areas", "afv_get_perimeters", "afv_get_arclens",
    "afv_get_coord_nums", "afv_get_num_connections", "afv_get_connections", "afv_get_num_ring_entries", "afv_get_rings",
    "afv_get_diagnostics", "afv_set_profiling", "afv_get_stage_times", "afv_launch_count", "afv_host_alloc",
    "afv_host_free", "afv_measure_fp64_peak", "afv_device_ptr", "afv_stream", "afv_synchronize", "afv_set_deferred_checks",
    "afv_set_slab", "afv_set_cells_device", "afv_extract_slab", "afv_drop_halo", "afv_append_cells", "afv_num_owned", "afv_pack_owned", "afv_pack_pair", "afv_finish_pair", "afv_pack_exchange", "afv_finish_exchange", "afv_finish_exchange_step", "afv_late_step_begin", "afv_late_finish", "afv_late_step_end", "afv_stream_wait", "afv_get_exchange_stats",
    "afv_peer_buffer_create", "afv_peer_buffer_open", "afv_flag_signal", "afv_flag_wait",
    "afv_cluster_analysis", "afv_get_cluster_labels", "afv_get_cluster_sizes", "afv_components_of_edges",
    "afv_get_energy", "afv_get_cell_energies", "afv_relax",
]


class AfvParams(C.Structure):
    _fields_ = [(k, C.c_double) for k in ("r", "A0", "P0", "KA", "KP", "Lambda", "delta")]


class AfvError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"afv_b200 error {code}: {msg}")
        self.code = code


_lib = None


def load() -> C.CDLL:
    """Load libafv_b200.so (built in-tree by ``python -m pyafv_b200._build`` / ``__graft_entry__.build()``)."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise ImportError(f"{LIB_PATH} is missing: build it with `python -m pyafv_b200._build` "
                          "(this backend has no CPU fallback)")
    lib = C.CDLL(str(LIB_PATH))
    H, D, I64 = C.c_void_p, C.c_double, C.c_int64
    P = C.c_void_p
    sig = {
        "afv_create": (C.c_int, [C.c_int, C.POINTER(AfvParams), D, P, C.POINTER(H)]),
        "afv_destroy": (None, [H]),
        "afv_last_error": (C.c_char_p, [H]),
        "afv_build_info": (C.c_char_p, []),
        "afv_set_params": (C.c_int, [H, C.POINTER(AfvParams)]),
        "afv_set_positions": (C.c_int, [H, P, I64]),
        "afv_set_preferred_areas": (C.c_int, [H, P, I64]),
        "afv_set_theta": (C.c_int, [H, P, I64]),
        "afv_num_cells": (I64, [H]),
        "afv_build": (C.c_int, [H, C.c_uint]),
        "afv_step": (C.c_int, [H, D, D, D, D, C.c_uint64, I64, C.c_int, P]),
        "afv_philox_normal": (D, [C.c_uint64, I64, I64]),
        "afv_get_positions": (C.c_int, [H, P]),
        "afv_get_theta": (C.c_int, [H, P]),
        "afv_get_forces": (C.c_int, [H, P]),
        "afv_get_areas": (C.c_int, [H, P]),
        "afv_get_perimeters": (C.c_int, [H, P]),
        "afv_get_arclens": (C.c_int, [H, P]),
        "afv_get_coord_nums": (C.c_int, [H, P]),
        "afv_get_num_connections": (C.c_int, [H, C.POINTER(I64)]),
        "afv_get_connections": (C.c_int, [H, P]),
        "afv_get_num_ring_entries": (C.c_int, [H, C.POINTER(I64)]),
        "afv_get_rings": (C.c_int, [H, P, P, P]),
        "afv_get_diagnostics": (C.c_int, [H, P]),
        "afv_set_profiling": (C.c_int, [H, C.c_int]),
        "afv_get_stage_times": (C.c_int, [H, P, P, C.c_int]),
        "afv_launch_count": (I64, [H]),
        "afv_host_alloc": (P, [I64]),
        "afv_host_free": (None, [P]),
        "afv_measure_fp64_peak": (C.c_int, [H, C.POINTER(D)]),
        "afv_device_ptr": (P, [H, C.c_int]),
        "afv_stream": (P, [H]),
        "afv_synchronize": (C.c_int, [H]),
        "afv_set_deferred_checks": (C.c_int, [H, C.c_int]),
        "afv_set_slab": (C.c_int, [H, D, D]),
        "afv_set_cells_device": (C.c_int, [H, P, I64, I64]),
        "afv_extract_slab": (C.c_int, [H, D, D, D, C.c_int, P, I64, C.POINTER(I64)]),
        "afv_drop_halo": (C.c_int, [H]),
        "afv_append_cells": (C.c_int, [H, P, I64, C.c_int]),
        "afv_num_owned": (I64, [H]),
        "afv_pack_pair": (C.c_int, [H, D, D, D, D, D, D, P, P, I64]),
        "afv_finish_pair": (C.c_int, [H, P, P, I64, C.c_int, C.c_int]),
        "afv_pack_exchange": (C.c_int, [H, D, D, D, D, D, P, P, I64]),
        "afv_finish_exchange": (C.c_int, [H, P, P, I64]),
        "afv_finish_exchange_step": (C.c_int, [H, P, P, I64, D, D, D, D, C.c_uint64, I64]),
        "afv_late_step_begin": (C.c_int, [H, I64]),
        "afv_late_finish": (C.c_int, [H, P, P, I64, P]),
        "afv_late_step_end": (C.c_int, [H, D, D, D, D, C.c_uint64, I64, D, D, D, D, D, P, P, I64, P]),
        "afv_stream_wait": (C.c_int, [H, P]),
        "afv_peer_buffer_create": (C.c_int, [H, I64, C.POINTER(P), P]),
        "afv_peer_buffer_open": (C.c_int, [H, P, C.POINTER(P)]),
        "afv_flag_signal": (C.c_int, [H, P, P, P, C.c_uint]),
        "afv_flag_wait": (C.c_int, [H, P, P, P, C.c_uint]),
        "afv_get_exchange_stats": (C.c_int, [H, P]),
        "afv_pack_owned": (C.c_int, [H, P, I64, C.POINTER(I64)]),
        "afv_cluster_analysis": (C.c_int, [H, C.POINTER(I64)]),
        "afv_get_cluster_labels": (C.c_int, [H, P]),
        "afv_get_cluster_sizes": (C.c_int, [H, P]),
        "afv_components_of_edges": (C.c_int, [H, I64, P, I64, P, P, C.POINTER(I64)]),
        "afv_get_energy": (C.c_int, [H, C.POINTER(D)]),
        "afv_get_cell_energies": (C.c_int, [H, P]),
        "afv_relax": (C.c_int, [H, D, D, D, C.c_int, C.c_int, P, P, C.c_int, C.POINTER(C.c_int)]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class Handle:
    """Thin RAII wrapper of an AfvHandle (one GPU, one stream)."""

    def __init__(self, params: AfvParams, box_L: float = 0.0, device: int = 0, stream: int | None = None):
        self.lib = load()
        self._h = C.c_void_p()
        # stream: None -> the library creates its own stream; 0 (torch's default stream) -> cudaStreamLegacy (0x1),
        # because a NULL pointer means "create one" in the C-ABI; anything else is a raw cudaStream_t
        sptr = 0 if stream is None else (1 if int(stream) == 0 else int(stream))
        rc = self.lib.afv_create(int(device), C.byref(params), float(box_L), C.c_void_p(sptr), C.byref(self._h))
        if rc != AFV_OK:
            raise AfvError(rc, self.lib.afv_last_error(None).decode())

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self.lib.afv_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def check(self, rc: int):
        if rc != AFV_OK:
            raise AfvError(rc, self.lib.afv_last_error(self._h).decode())

    def call(self, name: str, *args):
        self.check(getattr(self.lib, name)(self._h, *args))

    # --- typed helpers ---
    def fetch(self, name: str, shape, dtype=np.float64) -> np.ndarray:
        out = np.empty(shape, dtype=dtype)
        self.call(name, out.ctypes.data)
        return out

    @property
    def n(self) -> int:
        return int(self.lib.afv_num_cells(self._h))


class PinnedArray:
    """numpy view over page-locked host memory from afv_host_alloc (for end-to-end runs)."""

    def __init__(self, shape, dtype=np.float64):
        self.lib = load()
        nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        self.ptr = self.lib.afv_host_alloc(max(nbytes, 1))
        if not self.ptr:
            raise AfvError(AFV_ERR_CUDA, "afv_host_alloc failed")
        buf = (C.c_char * max(nbytes, 1)).from_address(self.ptr)
        self.array = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    def free(self):
        if self.ptr:
            self.array = None
            self.lib.afv_host_free(self.ptr)
            self.ptr = None
