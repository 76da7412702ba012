"""``FiniteVoronoiSimulator`` -- the reference's simulator facade on top of the sm_100a CUDA library.

Drop-in for ``pyafv.FiniteVoronoiSimulator`` (``pyafv/finite_voronoi.py:35-93, 866-945, 1153-1240``): same
constructor, ``build(connect=True)`` result dictionary, ``update_positions`` / ``update_params`` /
``update_preferred_areas`` / ``preferred_areas``, same ``ValueError`` / ``TypeError`` messages.  Additions the
reference leaves to user code: native periodic boundaries (``periodic=L`` instead of ``tile_pbc``) and a fused
``step()`` (``examples/active_FV.py:43-54``).

All numerics run on the GPU through the C-ABI of ``include/afv_b200.h``; there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .physical_params import PhysicalParams


def _as_params(phys: PhysicalParams) -> _lib.AfvParams:
    return _lib.AfvParams(phys.r, phys.A0, phys.P0, phys.KA, phys.KP, phys.Lambda, phys.delta)


class FiniteVoronoiSimulator:
    """Simulator for the active finite Voronoi (AFV) model, GPU resident.

    Args:
        pts: (N,2) initial cell centres.
        phys: :class:`PhysicalParams`.
        backend: ``None`` or ``"cuda"``.  The reference's ``"python"`` / ``"cython"`` values are rejected: this
            package has a single backend and no CPU fallback.
        periodic: side L of the periodic square ``[0, L)^2`` (``None`` = open boundaries).  Replaces the
            reference's ``tile_pbc(pts, L, r)`` + ``forces[:N]`` idiom (``_connectutil.py:12-96``).
        device: CUDA device ordinal.
        stream: optional raw ``cudaStream_t`` (e.g. ``torch.cuda.current_stream().cuda_stream``).
    """

    def __init__(self, pts: np.ndarray, phys: PhysicalParams, backend: str | None = None, *,
                 periodic: float | None = None, device: int = 0, stream: int | None = None):
        pts = np.asarray(pts, dtype=float)
        if pts.ndim != 2 or pts.shape[1] != 2:
            raise ValueError("pts must have shape (N,2)")
        if not isinstance(phys, PhysicalParams):
            raise TypeError("phys must be an instance of PhysicalParams")
        if backend not in (None, "cuda"):
            raise ValueError(f"backend={backend!r} is not available: pyafv_b200 runs on its CUDA backend only "
                             "(no CPU fallback)")
        from .backend import _BACKEND_NAME, _IMPORT_ERROR
        if _BACKEND_NAME != "cuda":
            raise ImportError(f"the CUDA backend could not be loaded: {_IMPORT_ERROR!r}")
        self._BACKEND = _BACKEND_NAME
        self.periodic = None if periodic is None else float(periodic)
        if self.periodic is not None and not self.periodic > 0.0:
            raise ValueError("periodic must be a positive box length or None")
        self.phys = phys
        self.N = pts.shape[0]
        self._pts = pts.copy()          # the reference copies in the constructor (finite_voronoi.py:70)
        self._pts_stale = False         # True when the device holds newer positions than self._pts
        self._preferred_areas = np.full(self.N, phys.A0, dtype=float)
        self._step_index = 0
        self._h = _lib.Handle(_as_params(phys), self.periodic or 0.0, device=device, stream=stream)
        self._upload_positions()

    # ------------------------------------------------------------------------------------------------
    def close(self) -> None:
        """Release the device memory of this simulator."""
        self._h.close()

    def _upload_positions(self) -> None:
        xy = np.ascontiguousarray(self._pts, dtype=np.float64)
        self._h.call("afv_set_positions", xy.ctypes.data, self.N)
        self._pts_stale = False

    def _upload_areas(self) -> None:
        a0 = np.ascontiguousarray(self._preferred_areas, dtype=np.float64)
        if self.N:
            self._h.call("afv_set_preferred_areas", a0.ctypes.data, self.N)

    @property
    def pts(self) -> np.ndarray:
        """(N,2) current cell centres (fetched from the device after ``step()``)."""
        if self._pts_stale:
            self._pts = self._h.fetch("afv_get_positions", (self.N, 2))
            self._pts_stale = False
        return self._pts

    @pts.setter
    def pts(self, value) -> None:
        self.update_positions(value)

    def positions_into(self, out: np.ndarray) -> np.ndarray:
        """Copy the current cell centres from the device into ``out`` ((N,2) float64, C-contiguous; e.g. a pinned
        host buffer) without allocating."""
        if out.shape != (self.N, 2) or out.dtype != np.float64 or not out.flags.c_contiguous:
            raise ValueError(f"out must be a C-contiguous float64 array of shape ({self.N},2)")
        self._h.call("afv_get_positions", out.ctypes.data)
        return out

    @property
    def theta(self) -> np.ndarray:
        """(N,) self-propulsion angles held on the device."""
        return self._h.fetch("afv_get_theta", (self.N,))

    # ------------------------------------------------------------------------------------------------
    def build(self, connect: bool = True, *, rings: bool = True) -> dict[str, object]:
        """Tessellate, measure and compute forces for the current positions.

        Returns the reference's dictionary (``finite_voronoi.py:935-945``): ``forces`` (N,2), ``areas``,
        ``perimeters``, ``arclens`` (N,), ``coord_nums`` (N,) int64, ``connections`` (K,2) int64 with i<j (rows
        sorted lexicographically), ``vertices`` (M,2), ``regions`` and ``edges_type`` (lists of lists, clockwise,
        1 = straight contact edge, 0 = arc).  Vertex ids are implementation-defined, as in the reference: every
        ring entry owns one row of ``vertices``.

        ``rings=False`` skips the ragged export (``vertices`` is empty, ``regions`` / ``edges_type`` are empty
        lists); use it inside time loops.
        """
        flags = _lib.BUILD_FORCES | (_lib.BUILD_CONNECT if connect else 0) | (_lib.BUILD_RINGS if rings else 0)
        h, N = self._h, self.N
        h.call("afv_build", flags)
        out = dict(
            forces=h.fetch("afv_get_forces", (N, 2)),
            areas=h.fetch("afv_get_areas", (N,)),
            perimeters=h.fetch("afv_get_perimeters", (N,)),
            arclens=h.fetch("afv_get_arclens", (N,)),
        )
        if rings:
            T = C.c_int64()
            h.call("afv_get_num_ring_entries", C.byref(T))
            off = np.zeros(N + 1, dtype=np.int64)
            types = np.zeros(T.value, dtype=np.int8)
            verts = np.zeros((T.value, 2), dtype=np.float64)
            h.call("afv_get_rings", off.ctypes.data, types.ctypes.data, verts.ctypes.data)
            ids = np.arange(T.value, dtype=np.int64)
            cuts = off[1:-1]
            out["vertices"] = verts
            out["edges_type"] = [a.tolist() for a in np.split(types.astype(np.int64), cuts)] if N else []
            out["regions"] = [a.tolist() for a in np.split(ids, cuts)] if N else []
        else:
            out["vertices"] = np.zeros((0, 2), dtype=np.float64)
            out["edges_type"] = []
            out["regions"] = []
        if connect:
            K = C.c_int64()
            h.call("afv_get_num_connections", C.byref(K))
            conn = np.zeros((K.value, 2), dtype=np.int64)
            h.call("afv_get_connections", conn.ctypes.data)
            if K.value:
                conn = conn[np.lexsort((conn[:, 1], conn[:, 0]))]
            out["connections"] = conn
        else:
            out["connections"] = np.empty((0, 2), dtype=np.int64)
        out["coord_nums"] = h.fetch("afv_get_coord_nums", (N,), dtype=np.int64)
        return out

    # ------------------------------------------------------------------------------------------------
    def step(self, dt: float, mu: float = 1.0, v0: float = 0.0, Dr: float = 0.0, *, theta: np.ndarray | None = None,
             noise: np.ndarray | None = None, seed: int | None = None, n_steps: int = 1) -> None:
        r"""``n_steps`` fused build + overdamped active-Langevin updates on the device.

        One step is the user-level loop of ``examples/active_FV.py:43-54``:
        :math:`r_i \leftarrow r_i + (\mu F_i + v_0 \hat n(\theta_i))\,dt`, then
        :math:`\theta_i \leftarrow \theta_i + \sqrt{2 D_r dt}\,\xi_i`; positions are wrapped into ``[0, L)`` when
        periodic.  ``noise`` (n_steps, N) supplies a recorded :math:`\xi` stream; otherwise :math:`\xi` comes from
        the counter-based Philox stream keyed by (seed, cell index, step counter), independent of cell order and
        GPU count.  ``theta`` (N,) replaces the angles before stepping.
        """
        if n_steps < 0:
            raise ValueError("n_steps must be non-negative")
        if theta is not None:
            th = np.ascontiguousarray(theta, dtype=np.float64)
            if th.shape != (self.N,):
                raise ValueError(f"theta must have shape ({self.N},)")
            if self.N:
                self._h.call("afv_set_theta", th.ctypes.data, self.N)
        ptr = None
        if noise is not None:
            nz = np.ascontiguousarray(noise, dtype=np.float64)
            if nz.shape == (self.N,) and n_steps == 1:
                nz = nz.reshape(1, self.N)
            if nz.shape != (n_steps, self.N):
                raise ValueError(f"noise must have shape ({n_steps},{self.N})")
            ptr = nz.ctypes.data
        if seed is None:
            seed = self.phys.seed if self.phys.seed is not None else 0
        self._h.call("afv_step", float(dt), float(mu), float(v0), float(Dr), C.c_uint64(int(seed)),
                     C.c_int64(self._step_index), int(n_steps), ptr)
        self._step_index += int(n_steps)
        if self.N and n_steps:
            self._pts_stale = True

    # ------------------------------------------------------------------------------------------------
    def energy(self) -> float:
        r"""Total energy :math:`E = \sum_i K_A (A_i - A_{0,i})^2 + K_P (P_i - P_0)^2 + \Lambda L_{arc,i}` of the last
        ``build()`` / ``step()`` geometry, reduced on the device in a fixed order.  This is the energy whose gradient the
        reference's force code implements (``finite_voronoi.py:529,746``); the reference never evaluates it."""
        e = C.c_double()
        self._h.call("afv_get_energy", C.byref(e))
        return float(e.value)

    def cell_energies(self) -> np.ndarray:
        """(N,) per-cell energies of the last build (they sum to :meth:`energy`)."""
        return self._h.fetch("afv_get_cell_energies", (self.N,))

    def relax(self, max_steps: int = 2000, *, dt: float = 0.01, dt_max: float | None = None, f_tol: float = 1e-6,
              check_every: int = 20) -> dict[str, object]:
        """Energy minimisation by FIRE (Bitzek et al., PRL 97, 170201) on the device: replaces the fixed-``dt`` Euler loop
        of ``examples/relaxation.py:41-45``.  Build, forces, the global sums, the adaptive time step and the move of one
        iteration never leave the GPU; the host looks at ``max|F|`` every ``check_every`` iterations.

        Returns ``steps``, ``energy``, ``fmax``, ``dt`` (last time step), ``converged`` and ``energies`` (the energy at every
        check).  The geometry left on the device is that of the final positions (``forces`` etc. are current)."""
        if max_steps < 0:
            raise ValueError("max_steps must be non-negative")
        dt_max = 10.0 * dt if dt_max is None else float(dt_max)
        n_e = max_steps // max(check_every, 1) + 3
        energies = np.zeros(n_e)
        out = np.zeros(5)
        nchk = C.c_int()
        self._h.call("afv_relax", float(dt), dt_max, float(f_tol), int(max_steps), int(check_every), out.ctypes.data,
                     energies.ctypes.data, int(n_e), C.byref(nchk))
        if self.N:
            self._pts_stale = True
        return dict(steps=int(out[0]), energy=float(out[1]), fmax=float(out[2]), dt=float(out[3]), converged=bool(out[4]),
                    energies=energies[: nchk.value].copy())

    def cluster_sizes(self) -> tuple[np.ndarray, int, np.ndarray]:
        """Connected components of the contact graph of the last build, computed on the device from the rings (no
        ``connections`` array is built or copied).  Same return value as the reference's
        ``get_cluster_sizes(N, connections)`` (``_connectutil.py:154-172``): ``(cluster_sizes, n_components, labels)``,
        labels numbered like ``scipy.sparse.csgraph.connected_components`` numbers them."""
        nc = C.c_int64()
        self._h.call("afv_cluster_analysis", C.byref(nc))
        labels = self._h.fetch("afv_get_cluster_labels", (self.N,), dtype=np.int64)
        sizes = self._h.fetch("afv_get_cluster_sizes", (nc.value,), dtype=np.int64)
        return sizes, int(nc.value), labels

    # ------------------------------------------------------------------------------------------------
    def update_positions(self, pts: np.ndarray, A0: float | np.ndarray | None = None) -> None:
        """New cell centres (uploaded immediately: the caller may mutate its array afterwards)."""
        pts = np.asarray(pts, dtype=float)
        if pts.ndim != 2 or pts.shape[1] != 2:
            raise ValueError("pts must have shape (N,2)")
        N = pts.shape[0]
        self._pts = pts
        changed = N != self.N
        self.N = N
        self._upload_positions()   # resets A0 on the device when N changed
        if changed:
            if A0 is None:
                self._preferred_areas = np.full(N, self.phys.A0, dtype=float)
            else:
                self.update_preferred_areas(A0)
        elif A0 is not None:
            self.update_preferred_areas(A0)

    def update_params(self, phys: PhysicalParams) -> None:
        """New physical parameters; also resets every preferred area to ``phys.A0`` like the reference."""
        if not isinstance(phys, PhysicalParams):
            raise TypeError("phys must be an instance of PhysicalParams")
        self.phys = phys
        p = _as_params(phys)
        self._h.call("afv_set_params", C.byref(p))
        self.update_preferred_areas(phys.A0)

    def update_preferred_areas(self, A0: float | np.ndarray) -> None:
        """Scalar, length-1 or (N,) preferred areas."""
        arr = np.asarray(A0, dtype=float)
        if arr.ndim == 0:
            arr = np.full(self.N, float(arr), dtype=float)
        elif arr.shape == (1,):
            arr = np.full(self.N, float(arr[0]), dtype=float)
        elif arr.shape != (self.N,):
            raise ValueError(f"A0 must be scalar or have shape ({self.N},)")
        self._preferred_areas = arr
        self._upload_areas()

    @property
    def preferred_areas(self) -> np.ndarray:
        """Copy of the per-cell preferred areas."""
        return self._preferred_areas.copy()

    # ------------------------------------------------------------------------------------------------
    def plot_2d(self, ax=None, show: bool = False, **kw):
        """Draw the finite Voronoi cells (straight contact edges and medium arcs) of the current positions.

        A light-weight counterpart of the reference's ``plot_2d`` (``finite_voronoi.py:948-1107``): it consumes the
        same ``build()`` dictionary (``vertices`` / ``regions`` / ``edges_type``), so the reference's own
        ``visualize_2d(pts, diag, r)`` works on this backend's output as well.  Needs matplotlib.
        """
        import matplotlib.pyplot as plt   # imported lazily, like the reference does
        ax = ax if ax is not None else plt.gca()
        diag, pts, r = self.build(connect=False), self.pts, self.phys.r
        V = diag["vertices"]
        for i, (reg, types) in enumerate(zip(diag["regions"], diag["edges_type"])):
            if len(reg) < 2:
                ax.add_patch(plt.Circle(pts[i], r, fill=False, lw=kw.get("line_width", 1.0), color=kw.get("line_colors", "k")))
                continue
            for k, t in enumerate(types):
                a, b = V[reg[k]], V[reg[(k + 1) % len(reg)]]
                if t == 1:
                    ax.plot([a[0], b[0]], [a[1], b[1]], "-", color=kw.get("line_colors", "k"), lw=kw.get("line_width", 1.0))
                else:   # clockwise arc from a to b about the cell centre
                    t0, t1 = np.arctan2(*(a - pts[i])[::-1]), np.arctan2(*(b - pts[i])[::-1])
                    if t1 > t0:
                        t1 -= 2 * np.pi
                    th = np.linspace(t0, t1, max(int(abs(t1 - t0) / 0.05), 2))
                    ax.plot(pts[i, 0] + r * np.cos(th), pts[i, 1] + r * np.sin(th), "-", color=kw.get("line_colors", "k"), lw=kw.get("line_width", 1.0))
        if kw.get("show_points", True):
            ax.plot(pts[:, 0], pts[:, 1], ".", ms=kw.get("point_size", 2))
        ax.set_aspect("equal")
        if show:
            plt.show()
        return ax

    # ------------------------------------------------------------------------------------------------
    def diagnostics(self) -> dict[str, int]:
        """Counters of the last build: slow-path cells, unmatched ring look-ups, max ring length, bins."""
        out = np.zeros(4, dtype=np.int64)
        self._h.call("afv_get_diagnostics", out.ctypes.data)
        return dict(slow_path_cells=int(out[0]), lookup_misses=int(out[1]), max_ring_len=int(out[2]), n_bins=int(out[3]))


__all__ = ["FiniteVoronoiSimulator"]
