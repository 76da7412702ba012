"""``decompose_points`` / ``DomainDecomposition`` -- the reference's point-decomposition bookkeeping, kept importable.

The GPU backend does not need it (ownership is decided on the device, per step, by the slab exchange of
:mod:`pyafv_b200.sharded`), but user scripts written against the reference call it directly (``examples/parallel.py``,
``examples/MPI_wrap.py``, ``tests/test_parallel.py:13-109``).  Same contract as ``pyafv/parallel_voronoi.py:23-338``:
an ``nx x ny`` grid of owned boxes over ``domain_bounds`` (or the bounding box of the points); a point on an internal
boundary belongs to the box on its right / above (right-sided binning), the outermost boundary is clipped into the last
box; the local set of a domain is every point inside its box grown by ``halo_width`` on all sides, edges included;
domains come in row-major order (``domain_id = iy * nx + ix``).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


@dataclass(frozen=True)
class DomainDecomposition:
    """One owned box plus its halo: index bookkeeping only (same fields as the reference's dataclass)."""

    domain_id: int
    grid_ix: int
    grid_iy: int
    x_range: tuple[float, float]
    y_range: tuple[float, float]
    halo_x_range: tuple[float, float]
    halo_y_range: tuple[float, float]
    local_global_ids: np.ndarray      # global ids of the points in the halo box (ascending)
    owned_local_ids: np.ndarray       # positions, inside local_global_ids, of the points this domain owns
    local_pts: np.ndarray             # points[local_global_ids]


def _grid_shape(grid_shape) -> tuple[int, int]:
    nx, ny = (int(v) for v in grid_shape)
    if nx <= 0 or ny <= 0:
        raise ValueError("grid_shape entries must be positive")
    return nx, ny


def _bounds(domain_bounds):
    if domain_bounds is None:
        return None
    (x0, x1), (y0, y1) = domain_bounds
    x0, x1, y0, y1 = float(x0), float(x1), float(y0), float(y1)
    if not (x0 < x1 and y0 < y1):
        raise ValueError("domain_bounds must be ((xmin, xmax), (ymin, ymax)) with nonzero spans")
    return (x0, x1), (y0, y1)


def decompose_points(points: np.ndarray, grid_shape: tuple[int, int] = (2, 2), halo_width: float = 0.0, *,
                     domain_bounds=None, method: str = "dense") -> list[DomainDecomposition]:
    """Owned grid domains plus halo/local points; see the module docstring for the ownership rule.

    ``method`` (``"dense"``, ``"binned"``, ``"sorted_x"``) is accepted for compatibility; the three give identical
    results in the reference (``tests/test_parallel.py:43-90``) and one routine serves them here: the points are grouped
    by owner once, and a halo box only looks at the owner boxes it overlaps.
    """
    pts = np.asarray(points, dtype=float)
    if pts.ndim != 2 or pts.shape[1] != 2:
        raise ValueError("points must have shape (N,2)")
    if not np.isfinite(pts).all():
        raise ValueError("points must be finite")
    halo = float(halo_width)
    if halo < 0.0:
        raise ValueError("halo_width must be non-negative")
    if method not in ("dense", "binned", "sorted_x"):
        raise ValueError("method must be 'dense', 'binned', or 'sorted_x'")
    nx, ny = _grid_shape(grid_shape)
    bounds = _bounds(domain_bounds)
    if bounds is None:
        (x0, x1), (y0, y1) = ((pts[:, 0].min(), pts[:, 0].max()), (pts[:, 1].min(), pts[:, 1].max())) if len(pts) else ((0.0, 0.0), (0.0, 0.0))
        if nx * ny > 1 and (x0 == x1 or y0 == y1):
            raise ValueError("points must span a nonzero range in both x and y")
    else:
        (x0, x1), (y0, y1) = bounds
        if ((pts[:, 0] < x0) | (pts[:, 0] > x1) | (pts[:, 1] < y0) | (pts[:, 1] > y1)).any():
            raise ValueError("points must lie inside domain_bounds")
    xe, ye = np.linspace(x0, x1, nx + 1), np.linspace(y0, y1, ny + 1)
    # owner box: right-sided binning, the outermost boundary clipped into the last box
    ox = np.clip(np.searchsorted(xe, pts[:, 0], side="right") - 1, 0, nx - 1)
    oy = np.clip(np.searchsorted(ye, pts[:, 1], side="right") - 1, 0, ny - 1)
    owner = oy * nx + ox
    by_owner = np.argsort(owner, kind="stable")
    start = np.concatenate(([0], np.cumsum(np.bincount(owner, minlength=nx * ny))))
    out = []
    for dom in range(nx * ny):
        iy, ix = divmod(dom, nx)
        hx, hy = (xe[ix] - halo, xe[ix + 1] + halo), (ye[iy] - halo, ye[iy + 1] + halo)
        # owner boxes the halo box reaches
        cx0, cx1 = np.clip(np.searchsorted(xe, hx, side="right") - 1, 0, nx - 1)
        cy0, cy1 = np.clip(np.searchsorted(ye, hy, side="right") - 1, 0, ny - 1)
        cand = np.concatenate([by_owner[start[r * nx + cx0]:start[r * nx + cx1 + 1]] for r in range(cy0, cy1 + 1)])
        p = pts[cand]
        local = np.sort(cand[(p[:, 0] >= hx[0]) & (p[:, 0] <= hx[1]) & (p[:, 1] >= hy[0]) & (p[:, 1] <= hy[1])])
        owned = by_owner[start[dom]:start[dom + 1]]
        pos = np.searchsorted(local, owned)
        if len(owned) and not (pos < len(local)).all() or not np.array_equal(local[pos], owned):
            raise RuntimeError(f"Owned/local index mapping failed for domain {dom}")
        out.append(DomainDecomposition(domain_id=dom, grid_ix=ix, grid_iy=iy, x_range=(float(xe[ix]), float(xe[ix + 1])),
                                       y_range=(float(ye[iy]), float(ye[iy + 1])), halo_x_range=(float(hx[0]), float(hx[1])),
                                       halo_y_range=(float(hy[0]), float(hy[1])), local_global_ids=local, owned_local_ids=pos,
                                       local_pts=pts[local]))
    return out


__all__ = ["DomainDecomposition", "decompose_points"]
