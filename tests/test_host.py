"""CPU: host-side mirror of the reference interface (parameters, validation, error messages)."""
from __future__ import annotations

import numpy as np
import pytest

import pyafv_b200 as afv


def test_physical_params_defaults_and_validation():
    p = afv.PhysicalParams()
    assert (p.r, p.A0, p.P0, p.KA, p.KP, p.Lambda, p.seed) == (1.0, np.pi, 4.8, 1.0, 1.0, 0.2, None)
    assert p.delta == pytest.approx(0.45)
    assert afv.PhysicalParams(r=2).delta == pytest.approx(0.9)
    assert afv.PhysicalParams(delta=0.0).delta == 0.0
    q = p.replace(A0=5.0, delta=0.3)
    assert q.A0 == 5.0 and q.delta == 0.3 and p.A0 == np.pi
    with pytest.raises(Exception):
        p.r = 3.0                      # frozen
    with pytest.raises(TypeError):
        afv.PhysicalParams(r=True)
    with pytest.raises(TypeError):
        afv.PhysicalParams(KA="1")
    with pytest.raises(ValueError):
        afv.PhysicalParams(P0=float("nan"))
    with pytest.raises(ValueError):
        afv.PhysicalParams(seed=-1)
    with pytest.raises(TypeError):
        afv.PhysicalParams(seed=1.5)
    with pytest.raises(TypeError):
        afv.PhysicalParams(delta="x")
    assert isinstance(afv.PhysicalParams(r=np.float32(1.5)).r, float)


def test_doublet_closed_forms():
    """tests/test_core.py:38-57 of the reference."""
    phys = afv.PhysicalParams()
    l, d = phys.get_steady_state()
    assert abs(l - 0.87) < 1e-3 and abs(d - 0.75) < 1e-3
    params = phys.with_optimal_radius()
    assert abs(params.r - l) < 1e-6 and params.delta == pytest.approx(0.45 * params.r)
    assert abs(afv.target_delta(params, 4.0) - 0.16868) < 1e-4
    p2 = phys.with_optimal_radius(digits=1, delta=0.5)
    assert p2.r == round(l, 1) and p2.delta == 0.5
    with pytest.raises(ValueError):
        afv.target_delta(params, 1e9)


def test_constructor_argument_errors_precede_device_use():
    with pytest.raises(ValueError, match=r"pts must have shape \(N,2\)"):
        afv.FiniteVoronoiSimulator(np.zeros((3, 3)), afv.PhysicalParams())
    with pytest.raises(TypeError, match="phys must be an instance of PhysicalParams"):
        afv.FiniteVoronoiSimulator(np.zeros((3, 2)), dict(r=1.0))
    with pytest.raises(ValueError, match="no CPU fallback"):
        afv.FiniteVoronoiSimulator(np.zeros((3, 2)), afv.PhysicalParams(), backend="python")


def test_backend_module_names():
    from pyafv_b200 import backend
    assert hasattr(backend, "backend_impl") and hasattr(backend, "_BACKEND_NAME") and hasattr(backend, "_IMPORT_ERROR")
    assert backend._BACKEND_NAME == "cuda" and backend._IMPORT_ERROR is None


def test_slab_geometry():
    from pyafv_b200.sharded import SlabGeometry
    g0, g3 = SlabGeometry(100.0, 0, 4, 1.0), SlabGeometry(100.0, 3, 4, 1.0)
    assert (g0.lo, g0.hi, g3.lo, g3.hi) == (0.0, 25.0, 75.0, 100.0)
    # nothing is shifted across the box face: offsets to the slab are minimum images of the box's own coordinates
    assert g0.shift_to_left == 0.0 and g3.shift_to_right == 0.0
    assert np.allclose(g0.offset([99.5, 0.5, 26.0]), [-0.5, 0.5, 26.0])
    assert np.allclose(g3.offset([0.5, 74.0, 99.0]), [25.5, -1.0, 24.0])
    assert g0.halo == pytest.approx(4.01)
    with pytest.raises(ValueError):
        SlabGeometry(10.0, 0, 4, 1.0)
    with pytest.raises(ValueError):
        SlabGeometry(17.0, 0, 2, 1.0)       # two slabs: the two halo strips of the one neighbour would overlap


def test_tile_pbc_matches_reference_semantics():
    """Originals first, x images, then y images of the augmented set (pyafv/_connectutil.py:12-96)."""
    rng = np.random.default_rng(0)
    L, r = 12.0, 1.0
    pts = rng.random((200, 2)) * L
    pos, idx = afv.tile_pbc(pts, L, r)
    assert np.array_equal(pos[:200], pts) and np.array_equal(idx[:200], np.arange(200))
    d = pos - pts[idx]
    assert np.all(np.isin(np.round(d / L), (-1, 0, 1))) and np.allclose(d, np.round(d / L) * L)
    # every image within 4.01 r outside the box is present exactly once
    want = 0
    for sx in (-1, 0, 1):
        for sy in (-1, 0, 1):
            if sx or sy:
                q = pts + np.array([sx, sy]) * L
                want += int(((q[:, 0] >= -4.01 * r) & (q[:, 0] <= L + 4.01 * r) & (q[:, 1] >= -4.01 * r) & (q[:, 1] <= L + 4.01 * r)).sum())
    assert len(pos) - 200 == want


def test_parallel_simulator_argument_validation():
    P = afv.ParallelFiniteVoronoiSimulator
    with pytest.raises(ValueError, match=r"pts must have shape \(N,2\)"):
        P(np.zeros((3, 3)), afv.PhysicalParams())
    with pytest.raises(TypeError):
        P(np.zeros((3, 2)), None)
    with pytest.raises(ValueError, match="grid_shape"):
        P(np.zeros((3, 2)), afv.PhysicalParams(), grid_shape=(0, 2))
    with pytest.raises(ValueError, match="decomposition_method"):
        P(np.zeros((3, 2)), afv.PhysicalParams(), decomposition_method="fancy")


# ---- decompose_points / DomainDecomposition: the reference's own unit tests (tests/test_parallel.py:13-109) ----
def test_decompose_points_tracks_owned_and_local_ids():
    pts = np.array([[0.0, 0.0], [0.5, 0.0], [1.0, 0.0], [0.0, 0.5], [0.5, 0.5], [1.0, 1.0]])
    domains = afv.decompose_points(pts, grid_shape=(2, 2), halo_width=0.1, domain_bounds=((0.0, 1.0), (0.0, 1.0)))
    owned = np.concatenate([d.local_global_ids[d.owned_local_ids] for d in domains])
    assert len(domains) == 4 and np.array_equal(np.sort(owned), np.arange(len(pts)))
    for d in domains:
        assert np.array_equal(d.local_pts, pts[d.local_global_ids])
    # right-sided binning with the outermost boundary clipped: (0.5, 0) belongs to the box on its right, (1, 1) to the last box
    assert [sorted(d.local_global_ids[d.owned_local_ids].tolist()) for d in domains] == [[0], [1, 2], [3], [4, 5]]
    assert domains[1].halo_x_range == (0.4, 1.1) and domains[2].y_range == (0.5, 1.0) and (domains[3].grid_ix, domains[3].grid_iy) == (1, 1)


def test_decompose_points_methods_match_and_halos_are_inclusive():
    rng = np.random.default_rng(12)
    pts = rng.random((100, 2)) * np.array([8.0, 5.0])
    pts[:6] = np.array([[0.0, 0.0], [4.0, 0.0], [8.0, 0.0], [0.0, 5.0], [4.0, 5.0], [8.0, 5.0]])
    kw = dict(grid_shape=(4, 3), halo_width=0.4, domain_bounds=((0.0, 8.0), (0.0, 5.0)))
    dense, sx, binned = (afv.decompose_points(pts, method=m, **kw) for m in ("dense", "sorted_x", "binned"))
    for a, b, c in zip(dense, sx, binned):
        for d in (b, c):
            assert np.array_equal(d.local_global_ids, a.local_global_ids) and np.array_equal(d.owned_local_ids, a.owned_local_ids)
    for d in dense:      # brute force: every point inside the halo box, edges included
        (x0, x1), (y0, y1) = d.halo_x_range, d.halo_y_range
        want = np.flatnonzero((pts[:, 0] >= x0) & (pts[:, 0] <= x1) & (pts[:, 1] >= y0) & (pts[:, 1] <= y1))
        assert np.array_equal(d.local_global_ids, want)
    with pytest.raises(ValueError):
        afv.decompose_points(pts, method="fancy")
    with pytest.raises(ValueError):
        afv.decompose_points(pts + 10.0, **kw)               # outside domain_bounds


def test_decompose_points_one_domain_allows_zero_span_axes():
    for pts in (np.array([[1.0, 2.0]]), np.array([[0.0, 3.0], [1.0, 3.0]]), np.array([[4.0, 0.0], [4.0, 1.0]])):
        (d,) = afv.decompose_points(pts, grid_shape=(1, 1), halo_width=0.5)
        assert np.array_equal(d.local_global_ids, np.arange(len(pts))) and np.array_equal(d.owned_local_ids, np.arange(len(pts)))
        assert np.array_equal(d.local_pts, pts) and d.halo_x_range == (pts[:, 0].min() - 0.5, pts[:, 0].max() + 0.5)


def test_decompose_points_equals_the_live_reference_if_available():
    """When oracle/_ref (the compiled, unmodified reference) is present: every field of every domain is identical to the
    reference's decompose_points (pyafv/parallel_voronoi.py:125-338), for all three of its methods."""
    import build_ref
    if not build_ref.ref_available():
        pytest.skip("oracle/_ref not built on this machine")
    ref = build_ref.import_reference()
    rng = np.random.default_rng(12)
    for shape, halo, N in (((4, 3), 0.4, 100), ((2, 2), 0.1, 6), ((1, 1), 0.5, 3), ((5, 1), 1.3, 500), ((3, 4), 0.0, 200)):
        pts = rng.random((N, 2)) * np.array([8.0, 5.0])
        k = min(6, N)
        pts[:k] = np.array([[0, 0], [4, 0], [8, 0], [0, 5], [4, 5], [8, 5]], dtype=float)[:k]     # points on the domain boundaries
        for bounds in (None, ((0.0, 8.0), (0.0, 5.0))):
            for m in ("dense", "binned", "sorted_x"):
                a = afv.decompose_points(pts, shape, halo, domain_bounds=bounds, method=m)
                b = ref.decompose_points(pts, shape, halo, domain_bounds=bounds, method=m)
                assert len(a) == len(b)
                for da, db in zip(a, b):
                    for f in ("domain_id", "grid_ix", "grid_iy", "x_range", "y_range", "halo_x_range", "halo_y_range"):
                        assert getattr(da, f) == getattr(db, f), f
                    assert np.array_equal(da.local_global_ids, db.local_global_ids)
                    assert np.array_equal(da.owned_local_ids, db.owned_local_ids) and np.array_equal(da.local_pts, db.local_pts)
