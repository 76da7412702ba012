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


__all__ = ["tile_pbc"]
