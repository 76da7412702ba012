"""``ParallelFiniteVoronoiSimulator`` -- the reference's domain-decomposed simulator API, driving the GPUs of one node.

Mirrors ``pyafv/parallel_voronoi.py:421-720``: constructor arguments, context-manager protocol,
``build(connect=False, plot_mode=False)`` with its merged dictionary, ``update_*``.  The reference cuts the points into
``grid_shape`` rectangular domains with a 4.01 l halo and farms them to worker processes, re-decomposing on the parent
every build.  Here ``grid_shape = (nx, ny)`` asks for ``nx * ny`` x-slabs of :mod:`pyafv_b200.sharded` (1-d slabs:
with a 4 l halo on at most 8 GPUs a second direction buys nothing), dealt round-robin to the visible GPUs -- several
slabs share a GPU when there are fewer GPUs than slabs -- and driven from this one process.  Cells stay resident on
their GPU between calls; ``halo_width`` is the width of the slab halo.  An open-boundary system is embedded in a
periodic box wide enough that nothing interacts across its faces.  ``n_workers`` and ``decomposition_method`` have no
counterpart on the GPU and are only validated and recorded; ``domain_bounds`` fixes the embedding box.

Additions (as on :class:`FiniteVoronoiSimulator`): ``periodic=L`` and a fused ``step()``.
"""
from __future__ import annotations

import os

import numpy as np

from .decomposition import DomainDecomposition, _bounds, _grid_shape, decompose_points
from .finite_voronoi import FiniteVoronoiSimulator
from .physical_params import PhysicalParams


class ParallelFiniteVoronoiSimulator:
    """Drop-in for the reference's multiprocessing simulator; see the module docstring for what changes."""

    def __init__(self, pts: np.ndarray, phys: PhysicalParams, grid_shape: tuple[int, int] = (2, 2), n_workers: int | None = None, *,
                 halo_width: float | None = None, backend: str | None = None, domain_bounds=None,
                 decomposition_method: str = "dense", periodic: float | None = None, device: int = 0, devices=None):
        pts = np.asarray(pts, dtype=float)
        if pts.ndim != 2 or pts.shape[1] != 2:
            raise ValueError("pts must have shape (N,2)")
        if not isinstance(phys, PhysicalParams):
            raise TypeError("phys must be an instance of PhysicalParams")
        self.grid_shape = _grid_shape(grid_shape)
        self._auto_halo_width = halo_width is None
        self.halo_width = float(4.01 * phys.r if halo_width is None else halo_width)
        if self.halo_width < 0.0:
            raise ValueError("halo_width must be non-negative")
        self.n_workers = int(n_workers if n_workers is not None else (os.cpu_count() or 1))
        if self.n_workers <= 0:
            raise ValueError("n_workers must be positive")
        if decomposition_method not in ("dense", "binned", "sorted_x"):
            raise ValueError("decomposition_method must be 'dense', 'binned', or 'sorted_x'")
        if backend not in (None, "cuda"):
            raise ValueError(f"backend={backend!r} is not available: pyafv_b200 runs on its CUDA backend only (no CPU fallback)")
        self.domain_bounds = _bounds(domain_bounds)
        self.decomposition_method = decomposition_method
        self.backend = backend
        self.periodic = None if periodic is None else float(periodic)
        self.device = int(device)
        self._devices = None if devices is None else [int(d) for d in devices]
        self.phys = phys
        self.N = pts.shape[0]
        self._pts = pts.copy()
        self._theta = np.zeros(self.N)
        self._preferred_areas = np.full(self.N, phys.A0, dtype=float)
        self._step_index = 0
        self._single = None        # one handle: FiniteVoronoiSimulator (one slab, or plot_mode)
        self._sh = None            # several slabs: LoopbackShardedSimulator
        self._box = None           # (origin (2,), L) of the box the slabs live in
        self._stale = False        # the device holds newer positions than self._pts
        self._setup()

    # ------------------------------------------------------------------------------------------------
    def _visible_devices(self) -> list[int]:
        if self._devices is not None:
            return self._devices
        import torch
        n = torch.cuda.device_count()
        if n <= 0:
            raise ImportError("no CUDA device available: pyafv_b200 has no CPU fallback")
        return list(range(n)) if self.device == 0 else [self.device]

    def _embedding(self, pts: np.ndarray) -> tuple[np.ndarray, float]:
        """Origin and side of the periodic box the slabs live in.  Open boundaries: the bounding box of the points (or
        ``domain_bounds``) with room to move and a margin of 2.5 l all around, so that no cell sees a periodic image."""
        if self.periodic is not None:
            return np.zeros(2), self.periodic
        r = self.phys.r
        if self.domain_bounds is not None:
            (x0, x1), (y0, y1) = self.domain_bounds
        elif len(pts):
            (x0, y0), (x1, y1) = pts.min(axis=0), pts.max(axis=0)
        else:
            x0 = y0 = 0.0; x1 = y1 = 1.0
        span = max(x1 - x0, y1 - y0, 4.0 * r)
        slack = 0.0 if self.domain_bounds is not None else 0.25 * span + 2.0 * r     # cells may spread before the box is rebuilt
        margin = 2.5 * r + slack
        return np.array([x0 - margin, y0 - margin]), float(span + 2.0 * margin)

    def _n_slabs(self, L: float) -> int:
        """As many slabs as asked for, but no narrower than the halo allows (``SlabGeometry``'s two conditions)."""
        want = self.grid_shape[0] * self.grid_shape[1]
        pad = 0.5 * self.phys.r
        g = want
        while g > 1 and (L / g < self.halo_width or L / g + 2.0 * (self.halo_width + pad) > L):
            g -= 1
        return g

    def _setup(self) -> None:
        """(Re)create the device side for the current host state."""
        self.close()
        origin, L = self._embedding(self._pts)
        G = self._n_slabs(L)
        self.n_slabs = G
        if G <= 1:
            self._single = FiniteVoronoiSimulator(self._pts, self.phys, periodic=self.periodic, device=self._visible_devices()[0])
            self._single.update_preferred_areas(self._preferred_areas)
            self._single._step_index = self._step_index
            if self.N:
                self._single._h.call("afv_set_theta", np.ascontiguousarray(self._theta).ctypes.data, self.N)
            self._box = None
        else:
            from .sharded import LoopbackShardedSimulator
            devs = self._visible_devices()
            self._box = (origin, L)
            self._sh = LoopbackShardedSimulator(self._local(self._pts), self._theta, self.phys, L, G, A0=self._preferred_areas,
                                                halo_width=self.halo_width, devices=devs[: max(1, min(len(devs), G))])
            for s in self._sh.ranks:
                s._step_index = self._step_index
        self._stale = False

    def _local(self, pts: np.ndarray) -> np.ndarray:
        origin, L = self._box
        q = pts - origin
        return np.mod(q, L) if self.periodic is not None else q

    def _fits(self, pts: np.ndarray) -> bool:
        if self._box is None or self.periodic is not None:
            return True
        origin, L = self._box
        m = 2.5 * self.phys.r
        q = pts - origin
        return bool(len(pts) == 0 or (q.min() >= m and q.max() <= L - m))

    def _pull(self, want_results: bool = False):
        """Host copies of positions / angles from the slabs (after step())."""
        if self._sh is not None and self._stale:
            res = self._sh.build()
            self._pts = res["pts"] + self._box[0]
            self._theta = res["theta"]
            self._stale = False

    # the reference keeps a worker pool alive inside ``with sim:``; here the cells stay resident on the GPUs anyway
    def __enter__(self) -> "ParallelFiniteVoronoiSimulator":
        return self

    def __exit__(self, exc_type, exc, tb) -> None:
        self.close()

    def close(self) -> None:
        if getattr(self, "_single", None) is not None:
            self._single.close(); self._single = None
        if getattr(self, "_sh", None) is not None:
            self._sh.close(); self._sh = None

    @property
    def pts(self) -> np.ndarray:
        if self._single is not None:
            return self._single.pts
        self._pull()
        return self._pts

    @property
    def theta(self) -> np.ndarray:
        if self._single is not None:
            return self._single.theta
        self._pull()
        return self._theta

    # ------------------------------------------------------------------------------------------------
    def build(self, connect: bool = False, plot_mode: bool = False) -> dict[str, object]:
        """Merged dictionary of ``parallel_voronoi.py:596-633``: forces, areas, perimeters, arclens, coord_nums,
        connections (sorted unique pairs), pids (here: the CUDA ordinal of every slab), plot_mode (+ owned_global_ids /
        diag_plot: one entry, the whole system, built on one GPU -- plotting is a host-side consumer)."""
        if self._single is None and plot_mode:
            self._pull()
            one = FiniteVoronoiSimulator(self._pts, self.phys, periodic=self.periodic, device=self._visible_devices()[0])
            one.update_preferred_areas(self._preferred_areas)
            d = one.build(connect=connect, rings=True)
            one.close()
            pids = [s.device.index for s in self._sh.ranks]
        elif self._single is not None:
            d = self._single.build(connect=connect, rings=plot_mode)
            pids = [self._visible_devices()[0]]
        else:
            if self._stale and not self._fits(self._sh.build()["pts"] + self._box[0]):   # an open system outgrew its box
                self._pull(); self._setup()
            d = self._sh.build(connect=connect)
            self._pts, self._theta, self._stale = d["pts"] + self._box[0], d["theta"], False
            if not connect:
                d["connections"] = np.empty((0, 2), dtype=np.int64)
            pids = [s.device.index for s in self._sh.ranks]
        out = {k: d[k] for k in ("forces", "areas", "perimeters", "arclens", "coord_nums", "connections")}
        out["plot_mode"] = plot_mode
        out["pids"] = np.asarray(pids, dtype=int)
        if plot_mode:
            out["owned_global_ids"] = [np.arange(self.N)]
            out["diag_plot"] = [dict(vertices=d["vertices"], edges_type=d["edges_type"], regions=d["regions"])]
        return out

    def step(self, dt: float, mu: float = 1.0, v0: float = 0.0, Dr: float = 0.0, *, theta=None, seed: int | None = None,
             n_steps: int = 1) -> None:
        """``n_steps`` fused build + active-Langevin updates (``FiniteVoronoiSimulator.step``), the exchange between the
        slabs included.  The noise is keyed by (seed, cell index, step counter): the trajectory does not depend on the
        number of slabs."""
        if seed is None:
            seed = self.phys.seed if self.phys.seed is not None else 0
        if self._single is not None:
            self._single.step(dt, mu, v0, Dr, theta=theta, seed=seed, n_steps=n_steps)
            self._step_index = self._single._step_index
            return
        if theta is not None:
            th = np.asarray(theta, dtype=float)
            if th.shape != (self.N,):
                raise ValueError(f"theta must have shape ({self.N},)")
            self._pull()
            self._theta = th.copy()
            self._sh.set_cells(self._local(self._pts), self._theta, self._preferred_areas)
        if self.periodic is None and self._stale and not self._fits(self.pts):
            self._setup()
        self._sh.step(dt, mu, v0, Dr, seed=seed, n_steps=n_steps)
        self._step_index += int(n_steps)
        for s in self._sh.ranks:
            s._step_index = self._step_index
        self._stale = True

    # ------------------------------------------------------------------------------------------------
    def update_positions(self, pts: np.ndarray, A0=None) -> None:
        pts = np.array(pts, dtype=float)          # the reference copies here (parallel_voronoi.py:661)
        if pts.ndim != 2 or pts.shape[1] != 2:
            raise ValueError("pts must have shape (N,2)")
        if self._single is None:
            self._pull()                          # keeps the angles of the running system
        changed = pts.shape[0] != self.N
        self._pts, self.N = pts, pts.shape[0]
        if changed:
            self._theta = np.zeros(self.N)
            self._preferred_areas = np.full(self.N, self.phys.A0, dtype=float)
        if A0 is not None:
            self._set_areas(A0)
        if self._single is not None and self._n_slabs(self._embedding(pts)[1]) <= 1:
            self._single.update_positions(pts, self._preferred_areas if (changed or A0 is not None) else None)
        elif self._sh is not None and self._fits(pts) and self._n_slabs(self._box[1]) == self.n_slabs:
            self._sh.set_cells(self._local(pts), self._theta, self._preferred_areas)
            self._stale = False
        else:
            self._setup()

    def _set_areas(self, A0) -> None:
        arr = np.asarray(A0, dtype=float)
        if arr.ndim == 0:
            arr = np.full(self.N, float(arr))
        elif arr.shape == (1,):
            arr = np.full(self.N, float(arr[0]))
        elif arr.shape != (self.N,):
            raise ValueError(f"A0 must be scalar or have shape ({self.N},)")
        self._preferred_areas = arr.copy()

    def update_params(self, phys: PhysicalParams) -> None:
        if not isinstance(phys, PhysicalParams):
            raise TypeError("phys must be an instance of PhysicalParams")
        if self._single is None:
            self._pull()
        else:
            self._pts, self._theta = self._single.pts.copy(), self._single.theta
        self.phys = phys
        if self._auto_halo_width:
            self.halo_width = 4.01 * phys.r
        self._preferred_areas = np.full(self.N, phys.A0, dtype=float)   # like the reference (finite_voronoi.py:1205)
        self._setup()

    def update_preferred_areas(self, A0) -> None:
        self._set_areas(A0)
        if self._single is not None:
            self._single.update_preferred_areas(self._preferred_areas)
        else:
            self._pull()
            self._sh.set_cells(self._local(self._pts), self._theta, self._preferred_areas)

    @property
    def preferred_areas(self) -> np.ndarray:
        return self._preferred_areas.copy()


__all__ = ["ParallelFiniteVoronoiSimulator", "DomainDecomposition", "decompose_points"]
