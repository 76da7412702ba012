"""``ParallelFiniteVoronoiSimulator`` -- the reference's domain-decomposed simulator API on the GPU backend.

Mirrors ``pyafv/parallel_voronoi.py:421-720`` (constructor arguments, context-manager protocol, ``build(connect=False,
plot_mode=False)`` and its merged dictionary, ``update_*``).  The reference decomposes the points into
``grid_shape`` rectangular domains with a 4.01 l halo and farms them to worker processes because one CPU core is
too slow; here one B200 builds the whole (open-boundary) system faster than the reference's 24-worker pool
(DESIGN.md section 6), so the domains of an open-boundary system are simply not split: ``grid_shape``,
``n_workers``, ``halo_width``, ``domain_bounds`` and ``decomposition_method`` are validated and recorded but the
build runs on one GPU.  Periodic systems that need several GPUs use :mod:`pyafv_b200.sharded` (x-slabs, one process
per GPU).
"""
from __future__ import annotations

import os

import numpy as np

from .finite_voronoi import FiniteVoronoiSimulator
from .physical_params import PhysicalParams


class ParallelFiniteVoronoiSimulator:
    """Drop-in for the reference's multiprocessing simulator; see the module docstring for what changes."""

    def __init__(self, pts: np.ndarray, phys: PhysicalParams, grid_shape: tuple[int, int] = (2, 2), n_workers: int | None = None, *,
                 halo_width: float | None = None, backend: str | None = None, domain_bounds=None,
                 decomposition_method: str = "dense", periodic: float | None = None, device: int = 0):
        pts = np.asarray(pts, dtype=float)
        if pts.ndim != 2 or pts.shape[1] != 2:
            raise ValueError("pts must have shape (N,2)")
        if not isinstance(phys, PhysicalParams):
            raise TypeError("phys must be an instance of PhysicalParams")
        nx, ny = (int(v) for v in grid_shape)
        if nx <= 0 or ny <= 0:
            raise ValueError("grid_shape entries must be positive")
        self.grid_shape = (nx, ny)
        self._auto_halo_width = halo_width is None
        self.halo_width = float(4.01 * phys.r if halo_width is None else halo_width)
        if self.halo_width < 0.0:
            raise ValueError("halo_width must be non-negative")
        self.n_workers = int(n_workers if n_workers is not None else (os.cpu_count() or 1))
        if self.n_workers <= 0:
            raise ValueError("n_workers must be positive")
        if decomposition_method not in ("dense", "binned", "sorted_x"):
            raise ValueError("decomposition_method must be 'dense', 'binned', or 'sorted_x'")
        if domain_bounds is not None:
            (xmin, xmax), (ymin, ymax) = domain_bounds
            if not (float(xmin) < float(xmax) and float(ymin) < float(ymax)):
                raise ValueError("domain_bounds must be ((xmin, xmax), (ymin, ymax)) with nonzero spans")
        self.domain_bounds = domain_bounds
        self.decomposition_method = decomposition_method
        self.backend = backend
        self.device = device
        self._sim = FiniteVoronoiSimulator(pts, phys, backend=backend, periodic=periodic, device=device)

    # the reference keeps a worker pool alive inside ``with sim:``; there is nothing to keep alive here
    def __enter__(self) -> "ParallelFiniteVoronoiSimulator":
        return self

    def __exit__(self, exc_type, exc, tb) -> None:
        self.close()

    def close(self) -> None:
        pass

    @property
    def pts(self) -> np.ndarray:
        return self._sim.pts

    @property
    def N(self) -> int:
        return self._sim.N

    @property
    def phys(self) -> PhysicalParams:
        return self._sim.phys

    def build(self, connect: bool = False, plot_mode: bool = False) -> dict[str, object]:
        """Merged dictionary of ``parallel_voronoi.py:596-633``: forces, areas, perimeters, arclens, coord_nums,
        connections (sorted unique pairs), pids (here: the device ordinal), plot_mode (+ owned_global_ids / diag_plot)."""
        d = self._sim.build(connect=connect, rings=plot_mode)
        out = {k: d[k] for k in ("forces", "areas", "perimeters", "arclens", "coord_nums", "connections")}
        out["plot_mode"] = plot_mode
        out["pids"] = np.asarray([self.device], dtype=int)
        if plot_mode:
            out["owned_global_ids"] = [np.arange(self.N)]
            out["diag_plot"] = [dict(vertices=d["vertices"], edges_type=d["edges_type"], regions=d["regions"])]
        return out

    def step(self, *args, **kw) -> None:
        self._sim.step(*args, **kw)

    def update_positions(self, pts: np.ndarray, A0=None) -> None:
        self._sim.update_positions(np.array(pts, dtype=float), A0)   # the reference copies here (parallel_voronoi.py:661)

    def update_params(self, phys: PhysicalParams) -> None:
        self._sim.update_params(phys)
        if self._auto_halo_width:
            self.halo_width = 4.01 * phys.r

    def update_preferred_areas(self, A0) -> None:
        self._sim.update_preferred_areas(A0)

    @property
    def preferred_areas(self) -> np.ndarray:
        return self._sim.preferred_areas


__all__ = ["ParallelFiniteVoronoiSimulator"]
