"""``tile_pbc`` -- compatibility shim for scripts written against the reference's periodic recipe.

The reference has no periodic simulator: users tile the images of the cells near the box faces with
``tile_pbc(pts, L, r)`` (``pyafv/_connectutil.py:12-96``), build the open system and keep rows ``[:N]``.  This
backend handles periodic boxes natively (``FiniteVoronoiSimulator(..., periodic=L)``); the function is kept so that
such scripts run unchanged (and as the bridge the parity tests use towards the reference).
"""
from __future__ import annotations

import numpy as np


def tile_pbc(pts: np.ndarray, L: float, r: float | None = None) -> tuple[np.ndarray, np.ndarray]:
    """Originals first, then the images within ``min(4.01 r, L)`` of the x faces, then -- on that augmented set --
    of the y faces.  Returns (positions (M,2), index of the original cell (M,))."""
    pts = np.asarray(pts, dtype=float)
    n = len(pts)
    width = min(4.01 * (L if r is None else r), L)
    pos, idx = [pts], [np.arange(n, dtype=np.int64)]

    def add_images(axis: int) -> None:
        cur_pos, cur_idx = np.concatenate(pos), np.concatenate(idx)
        for mask, shift in ((cur_pos[:, axis] <= width, L), (cur_pos[:, axis] >= L - width, -L)):
            img = cur_pos[mask].copy()
            img[:, axis] += shift
            pos.append(img)
            idx.append(cur_idx[mask])

    add_images(0)
    add_images(1)
    return np.concatenate(pos), np.concatenate(idx)


# ------------------------------------------------------------------------------------------------------
# Connectivity analysis of a (K,2) contact list -- the callers of build()["connections"] (pyafv/_connectutil.py:101-172).
# Same signatures and return values as the reference; the connected components are found on the GPU (union by minimum
# with path compression, csrc/afv_cluster.cuh) instead of scipy.sparse.csgraph.  FiniteVoronoiSimulator.cluster_sizes()
# is the variant that never leaves the device.
# ------------------------------------------------------------------------------------------------------
_cc_handle = None


def _components(N: int, connect: np.ndarray) -> tuple[int, np.ndarray, np.ndarray]:
    """(n_components, labels (N,), sizes (n_components,)) of the graph with N nodes and edge list ``connect``."""
    import ctypes as C

    from . import _lib
    global _cc_handle
    N = int(N)
    edges = np.ascontiguousarray(np.asarray(connect, dtype=np.int64).reshape(-1, 2))
    if N == 0:
        return 0, np.zeros(0, dtype=np.int64), np.zeros(0, dtype=np.int64)
    if _cc_handle is None:     # a parameter-free handle: only its stream and scratch buffers are used
        _cc_handle = _lib.Handle(_lib.AfvParams(1.0, 1.0, 1.0, 1.0, 1.0, 0.0, 0.45), 0.0)
    labels, sizes, nc = np.zeros(N, dtype=np.int64), np.zeros(N, dtype=np.int64), C.c_int64()
    _cc_handle.call("afv_components_of_edges", N, edges.ctypes.data, len(edges), labels.ctypes.data, sizes.ctypes.data, C.byref(nc))
    return int(nc.value), labels, sizes[: nc.value].copy()


def rebuild_connection_matrix(N: int, connect: np.ndarray):
    """(N,N) boolean ``scipy.sparse.csr_matrix`` with both (i,j) and (j,i) set for every row of the (E,2) edge list."""
    from scipy import sparse
    e = np.asarray(connect, dtype=np.int64).reshape(-1, 2)
    ij = np.concatenate([e, e[:, ::-1]])
    return sparse.csr_matrix((np.ones(len(ij), dtype=bool), (ij[:, 0], ij[:, 1])), shape=(int(N), int(N)))


def get_cluster_sizes(N: int, connect: np.ndarray) -> tuple[np.ndarray, int, np.ndarray]:
    """``(cluster_sizes (n_components,), n_components, labels (N,))`` of the contact graph."""
    nc, labels, sizes = _components(N, connect)
    return sizes, nc, labels


def select_daughter_cluster(N: int, connect: np.ndarray) -> tuple[np.ndarray | None, int, np.ndarray]:
    """One connected component chosen at random (``np.random.randint``, like the reference): its global cell indices, its
    size, and its edges re-indexed to ``0 .. size-1``; ``(None, N, connect)`` when the graph is connected."""
    nc, labels, _ = _components(N, connect)
    if nc <= 1:
        return None, N, connect
    cells = np.flatnonzero(labels == np.random.randint(nc))
    e = np.asarray(connect).reshape(-1, 2)
    inside = e[(labels[e[:, 0]] == labels[cells[0]]) & (labels[e[:, 1]] == labels[cells[0]])] if len(e) else e
    return cells, len(cells), np.searchsorted(cells, inside)


__all__ = ["tile_pbc", "rebuild_connection_matrix", "get_cluster_sizes", "select_daughter_cluster"]
