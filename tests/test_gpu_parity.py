"""GPU parity: the CUDA path (through the C-ABI / FiniteVoronoiSimulator) against the committed golden
vectors of the unmodified reference, the reference's MATLAB vectors, and the CPU oracle on seeded inputs.

Tolerances (BASELINE.json north_star): topology bit-exact; areas, perimeters, arc lengths and forces within
rtol 1e-9 in FP64 -- asserted as |d| <= 1e-9 * max(1, |ref|_inf), the form SURVEY 7 derives from the reference's
own noise floor.
"""
from __future__ import annotations

import numpy as np
import pytest

from conftest import STATIC_CASES, load_golden, oracle_kwargs, phys_from, rings_equal_up_to_rotation

pytestmark = pytest.mark.gpu
RTOL = 1e-9


def _close(a, b, what):
    scale = max(1.0, float(np.abs(b).max()) if np.size(b) else 1.0)
    err = float(np.abs(np.asarray(a) - np.asarray(b)).max()) if np.size(b) else 0.0
    assert err <= RTOL * scale, f"{what}: max abs err {err:.3e} > {RTOL} * {scale:.3g}"


def _sim_for(g):
    import pyafv_b200 as afv
    L = float(g["L"]) if "L" in g else 0.0
    sim = afv.FiniteVoronoiSimulator(g["pts"], phys_from(g), periodic=L if L > 0 else None)
    if "A0" in g:
        sim.update_preferred_areas(g["A0"])
    return sim


def _flat_rings(diag):
    off = np.zeros(len(diag["regions"]) + 1, dtype=np.int64)
    off[1:] = np.cumsum([len(r) for r in diag["regions"]])
    typ = np.array([t for e in diag["edges_type"] for t in e], dtype=np.int8)
    return off, typ


@pytest.mark.parametrize("case", STATIC_CASES)
def test_build_matches_reference_golden(case):
    g = load_golden(case)
    sim = _sim_for(g)
    d = sim.build()
    _close(d["areas"], g["areas"], "areas")
    _close(d["perimeters"], g["perimeters"], "perimeters")
    _close(d["arclens"], g["arclens"], "arclens")
    _close(d["forces"], g["forces"], "forces")
    assert d["coord_nums"].dtype == np.int64 and np.array_equal(d["coord_nums"], g["coord_nums"])
    assert np.array_equal(d["connections"], g["connections"])
    off, typ = _flat_rings(d)
    assert rings_equal_up_to_rotation(off, typ, g["ring_off"], g["edges_type"])
    # ring vertex coordinates: same point sets per cell
    V = d["vertices"]
    for i in range(len(g["areas"])):
        a = V[off[i]:off[i + 1]]
        b = g["ring_xy"][g["ring_off"][i]:g["ring_off"][i + 1]]
        if len(a):
            dist = np.abs(a[:, None, :] - b[None, :, :]).max(axis=2).min(axis=1)
            assert dist.max() <= 1e-9 * max(1.0, np.abs(b).max())
    assert sim.diagnostics()["lookup_misses"] == 0
    sim.close()


def test_matlab_golden_forces(golden_dir):
    """The reference's own pins: tests/test_core.py:5-35 and tests/test_vary_A0.py:4-39 (abs 1e-8)."""
    import pyafv_b200 as afv
    pts = np.loadtxt(golden_dir / "init_pts.csv", delimiter=",")
    sim = afv.FiniteVoronoiSimulator(np.array([[0.0, 0.0]]), afv.PhysicalParams())
    sim.update_params(sim.phys.replace(delta=0.0))
    sim.update_positions(pts)
    F = sim.build()["forces"]
    assert np.abs(F - np.loadtxt(golden_dir / "init_forces.csv", delimiter=",")).max() < 1e-8
    A0 = np.loadtxt(golden_dir / "varying_A0.csv", delimiter=",")
    sim.update_positions(pts, A0)
    F = sim.build()["forces"]
    assert np.abs(F - np.loadtxt(golden_dir / "varying_A0_forces.csv", delimiter=",")).max() < 1e-8
    sim.close()


@pytest.mark.parametrize("N,dens,periodic,seed", [(20000, 1.0, False, 1), (20000, 1.0, True, 2), (20000, 0.1, True, 3),
                                                  (20000, 1.0 / np.pi, False, 4), (50000, 1.8, True, 5),
                                                  (100000, 1.0, True, 6)])
def test_build_matches_oracle_seeded(N, dens, periodic, seed):
    import pyafv_b200 as afv
    from afv_oracle import oracle_build
    rng = np.random.default_rng(seed)
    L = float(np.sqrt(N / dens))
    pts = rng.random((N, 2)) * L
    A0 = np.pi * (1.0 + 0.2 * rng.standard_normal(N))
    sim = afv.FiniteVoronoiSimulator(pts, afv.PhysicalParams(), periodic=L if periodic else None)
    sim.update_preferred_areas(A0)
    d = sim.build(rings=False)
    o = oracle_build(pts, A0=A0, periodic=L if periodic else None)
    for k in ("areas", "perimeters", "arclens", "forces"):
        _close(d[k], o[k], k)
    assert np.array_equal(d["coord_nums"], o["coord_nums"])
    assert np.array_equal(d["connections"], o["connections"])
    assert sim.diagnostics()["lookup_misses"] == 0
    sim.close()


@pytest.mark.parametrize("case", ["traj_open", "traj_pbc"])
def test_trajectory_100_steps(case):
    """100 Euler steps with the recorded noise stream vs the reference loop (examples/active_FV.py:43-54)."""
    import pyafv_b200 as afv
    g = load_golden(case)
    L = float(g["L"])
    mu, v0, Dr, dt = (float(v) for v in g["step"])
    sim = afv.FiniteVoronoiSimulator(g["pts0"], phys_from(g), periodic=L if L > 0 else None)
    n_steps = g["noise"].shape[0]
    done = 0
    first = True
    for stop in (1, 10, 50, n_steps):
        sim.step(dt, mu, v0, Dr, theta=g["theta0"] if first else None, noise=g["noise"][done:stop], n_steps=stop - done)
        first = False
        done = stop
        ref = g[f"pts_{stop}"]
        d = sim.pts - ref
        if L > 0:
            d -= L * np.round(d / L)
        assert np.abs(d).max() <= 1e-7, f"step {stop}: {np.abs(d).max():.3e}"
    assert np.abs(sim.theta - g["theta_final"]).max() <= 1e-10
    sim.close()


def test_determinism_and_order_independence():
    """Same input twice -> bit-identical; permuted input -> same values per cell (no atomics in the force path)."""
    import pyafv_b200 as afv
    rng = np.random.default_rng(9)
    N, L = 30000, float(np.sqrt(30000))
    pts = rng.random((N, 2)) * L
    sim = afv.FiniteVoronoiSimulator(pts, afv.PhysicalParams(), periodic=L)
    a = sim.build(connect=False, rings=False)
    b = sim.build(connect=False, rings=False)
    for k in ("forces", "areas", "perimeters", "arclens"):
        assert np.array_equal(a[k], b[k])
    perm = rng.permutation(N)
    sim.update_positions(pts[perm])
    c = sim.build(connect=False, rings=False)
    assert np.abs(c["forces"] - a["forces"][perm]).max() <= 1e-11 * max(1.0, np.abs(a["forces"]).max())
    assert np.array_equal(c["coord_nums"], a["coord_nums"][perm])
    sim.close()


def test_graph_replayed_time_loop_equals_single_step_calls():
    """A periodic ``step(n_steps >= 6)`` is replayed as a CUDA graph with the step index of the noise stream kept on the
    device; one-step calls take the plain launches with the index passed by the host.  Same bits either way, also across
    several long calls (the graph is kept in the handle) and after a parameter change (it is captured again)."""
    import pyafv_b200 as afv
    N = 20000
    L = float(np.sqrt(N))
    rng = np.random.default_rng(77)
    pts, theta = rng.random((N, 2)) * L, 2 * np.pi * rng.random(N) - np.pi
    phys = afv.PhysicalParams()
    a = afv.FiniteVoronoiSimulator(pts, phys, periodic=L)
    b = afv.FiniteVoronoiSimulator(pts, phys, periodic=L)
    a.step(0.01, 1.0, 2.4, 0.3, theta=theta, seed=5, n_steps=25)      # 2 plain steps + 11 replays + 1 plain step
    a.step(0.01, 1.0, 2.4, 0.3, seed=5, n_steps=16)                   # the kept graph (odd start parity: captured again)
    a.step(0.02, 1.0, 1.0, 0.1, seed=6, n_steps=10)                   # other parameters
    b.step(0.01, 1.0, 2.4, 0.3, theta=theta, seed=5, n_steps=1)
    for _ in range(40):
        b.step(0.01, 1.0, 2.4, 0.3, seed=5, n_steps=1)
    for _ in range(10):
        b.step(0.02, 1.0, 1.0, 0.1, seed=6, n_steps=1)
    assert np.array_equal(a.pts, b.pts) and np.array_equal(a.theta, b.theta)
    assert a.diagnostics()["lookup_misses"] == 0
    a.close(); b.close()


def test_philox_stream_matches_host_formula():
    """The device noise stream equals the host evaluation of the same counter-based generator."""
    import ctypes as C
    import pyafv_b200 as afv
    from pyafv_b200 import _lib
    N = 4096
    pts = np.random.default_rng(3).random((N, 2)) * 200.0   # isolated cells: F = 0, positions move by v0 only
    sim = afv.FiniteVoronoiSimulator(pts, afv.PhysicalParams())
    sim.step(0.5, 1.0, 0.0, 1.0, seed=1234, n_steps=1)     # theta <- 0 + sqrt(2*1*0.5) xi = xi
    th = sim.theta
    lib = _lib.load()
    want = np.array([lib.afv_philox_normal(C.c_uint64(1234), C.c_int64(i), C.c_int64(0)) for i in range(N)])
    assert np.abs(th - want).max() <= 1e-12
    assert abs(th.mean()) < 0.1 and abs(th.std() - 1.0) < 0.1
    sim.close()


def test_large_periodic_invariants():
    """N = 10^6 periodic (BASELINE config 4): size-independent properties -- zero net force, area tiling bound,
    determinism of the step, coordination = degree in the contact graph."""
    import pyafv_b200 as afv
    N = 1_000_000
    L = float(np.sqrt(N))
    pts = np.random.default_rng(42).random((N, 2)) * L
    sim = afv.FiniteVoronoiSimulator(pts, afv.PhysicalParams(), periodic=L)
    d = sim.build(connect=True, rings=False)
    F = d["forces"]
    assert np.abs(F.sum(axis=0)).max() <= 1e-8 * np.abs(F).sum()
    assert d["areas"].sum() <= L * L * (1 + 1e-12) and d["areas"].min() > 0
    deg = np.bincount(d["connections"].ravel(), minlength=N)
    assert np.array_equal(deg, d["coord_nums"])
    assert sim.diagnostics()["lookup_misses"] == 0
    sim.close()


@pytest.mark.parametrize("n_ring,radius", [(24, 1.0), (40, 0.9), (70, 1.2)])
def test_many_neighbour_cell_uses_overflow_pool(n_ring, radius):
    """A cell ringed by n_ring close neighbours has more ring entries than the fixed stride holds: the slow path
    and the overflow pool must give the oracle's numbers (no silent truncation)."""
    import pyafv_b200 as afv
    from afv_oracle import oracle_build
    rng = np.random.default_rng(n_ring)
    ang = 2 * np.pi * (np.arange(n_ring) + 0.1 * rng.random(n_ring)) / n_ring
    ringpts = radius * (1 + 0.01 * rng.random(n_ring))[:, None] * np.column_stack((np.cos(ang), np.sin(ang)))
    filler = (rng.random((400, 2)) - 0.5) * 30.0
    filler = filler[np.hypot(filler[:, 0], filler[:, 1]) > 2.5]
    pts = np.vstack([[0.0, 0.0], ringpts, filler]) + 15.0
    sim = afv.FiniteVoronoiSimulator(pts, afv.PhysicalParams())
    d = sim.build()
    o = oracle_build(pts, rings=True)
    n_expect = int(o["ring_off"][1] - o["ring_off"][0])
    assert n_expect > 16   # more than the fixed ring stride of the fast path
    for k in ("areas", "perimeters", "arclens", "forces"):
        _close(d[k], o[k], k)
    assert np.array_equal(d["coord_nums"], o["coord_nums"]) and d["coord_nums"][0] == n_expect
    assert np.array_equal(d["connections"], o["connections"])
    assert len(d["regions"][0]) == n_expect
    diag = sim.diagnostics()
    assert diag["slow_path_cells"] >= 1 and diag["lookup_misses"] == 0 and diag["max_ring_len"] == n_expect
    sim.close()


@pytest.mark.parametrize("n_ring", [5, 6, 7, 8])
def test_cell_with_five_to_eight_arcs_stays_on_the_fast_path(n_ring):
    """n_ring neighbours whose bisectors each cut the disk separately: n_ring contact edges with n_ring arcs in between, a
    ring of 2 n_ring <= 16 entries.  The fast path stores RS/2 = 8 arcs per cell (it stored four, and such cells were
    the whole population of the slow path at rho l^2 = 1); the neighbours look their arc ends up in the second half of
    the cell's arc table."""
    import pyafv_b200 as afv
    from afv_oracle import oracle_build
    rng = np.random.default_rng(200 + n_ring)
    ang = 2 * np.pi * (np.arange(n_ring) + 0.05 * rng.random(n_ring)) / n_ring
    ringpts = (1.9 + 0.05 * rng.random(n_ring))[:, None] * np.column_stack((np.cos(ang), np.sin(ang)))
    filler = (rng.random((400, 2)) - 0.5) * 30.0
    filler = filler[np.hypot(filler[:, 0], filler[:, 1]) > 4.2]
    pts = np.vstack([[0.0, 0.0], ringpts, filler]) + 15.0
    sim = afv.FiniteVoronoiSimulator(pts, afv.PhysicalParams())
    d = sim.build()
    o = oracle_build(pts, rings=True)
    assert int(o["ring_off"][1] - o["ring_off"][0]) == 2 * n_ring and o["coord_nums"][0] == n_ring
    for k in ("areas", "perimeters", "arclens", "forces"):
        _close(d[k], o[k], k)
    assert np.array_equal(d["coord_nums"], o["coord_nums"]) and np.array_equal(d["connections"], o["connections"])
    diag = sim.diagnostics()
    assert diag["slow_path_cells"] == 0 and diag["lookup_misses"] == 0
    sim.step(1e-3, 1.0, 0.0, 0.0, n_steps=3)
    _close(sim.build(connect=False, rings=False)["forces"], oracle_build(sim.pts)["forces"], "forces after 3 steps")
    assert sim.diagnostics()["lookup_misses"] == 0
    sim.close()


@pytest.mark.parametrize("n_ring", [9, 12, 15])
def test_long_ring_with_short_polygon_is_emitted_into_the_pool(n_ring):
    """n_ring <= 16 neighbours whose bisectors each cut the disk separately: the polygon fits the fast path but the ring
    (entry + exit point per neighbour, arcs in between) has 2 n_ring > 16 entries.  k_ring queues the cell and the slow
    kernel emits the ring again from the same polygon, into the pool."""
    import pyafv_b200 as afv
    from afv_oracle import oracle_build
    rng = np.random.default_rng(100 + n_ring)
    ang = 2 * np.pi * (np.arange(n_ring) + 0.05 * rng.random(n_ring)) / n_ring
    ringpts = (1.97 + 0.01 * rng.random(n_ring))[:, None] * np.column_stack((np.cos(ang), np.sin(ang)))
    filler = (rng.random((400, 2)) - 0.5) * 30.0
    filler = filler[np.hypot(filler[:, 0], filler[:, 1]) > 4.2]
    pts = np.vstack([[0.0, 0.0], ringpts, filler]) + 15.0
    sim = afv.FiniteVoronoiSimulator(pts, afv.PhysicalParams())
    d = sim.build()
    o = oracle_build(pts, rings=True)
    n_expect = int(o["ring_off"][1] - o["ring_off"][0])
    assert 16 < n_expect <= 32 and n_expect == 2 * n_ring
    for k in ("areas", "perimeters", "arclens", "forces"):
        _close(d[k], o[k], k)
    assert np.array_equal(d["coord_nums"], o["coord_nums"]) and d["coord_nums"][0] == n_ring
    assert np.array_equal(d["connections"], o["connections"])
    assert len(d["regions"][0]) == n_expect
    diag = sim.diagnostics()
    assert diag["slow_path_cells"] >= 1 and diag["lookup_misses"] == 0 and diag["max_ring_len"] == n_expect
    # and the same over a few fused steps (the pooled ring is looked up by its neighbours' force gathers)
    sim.step(1e-3, 1.0, 0.0, 0.0, n_steps=3)
    o2 = oracle_build(sim.pts)
    _close(sim.build(connect=False, rings=False)["forces"], o2["forces"], "forces after 3 steps")
    sim.close()


def test_parallel_simulator_equals_serial():
    """tests/test_parallel.py:112-176 of the reference: the parallel API returns the serial results."""
    import pyafv_b200 as afv
    g = load_golden("rand_dense")
    phys = phys_from(g)
    with afv.ParallelFiniteVoronoiSimulator(g["pts"], phys, grid_shape=(3, 2), n_workers=4, decomposition_method="binned") as sim:
        d = sim.build(connect=True, plot_mode=True)
    for k in ("forces", "areas", "perimeters", "arclens"):
        _close(d[k], g[k], k)
    assert np.array_equal(d["coord_nums"], g["coord_nums"]) and np.array_equal(d["connections"], g["connections"])
    assert d["plot_mode"] is True and len(d["diag_plot"]) == 1 and len(d["owned_global_ids"][0]) == len(g["pts"])
    assert set(d["diag_plot"][0]) == {"vertices", "edges_type", "regions"}
    d2 = afv.ParallelFiniteVoronoiSimulator(g["pts"], phys).build()
    assert d2["connections"].shape == (0, 2) and d2["plot_mode"] is False


def test_c_abi_error_paths():
    """Negative return codes + afv_last_error instead of exceptions across the ABI; state checks."""
    import ctypes as C
    import pyafv_b200 as afv
    from pyafv_b200 import _lib
    from pyafv_b200._lib import AfvError
    lib = _lib.load()
    h = C.c_void_p()
    p = _lib.AfvParams(1.0, np.pi, 4.8, 1.0, 1.0, 0.2, 0.45)
    assert lib.afv_create(0, C.byref(p), 3.0, None, C.byref(h)) == _lib.AFV_ERR_ARG          # periodic box < 4 r
    assert b"4 r" in lib.afv_last_error(None)
    assert lib.afv_create(99, C.byref(p), 0.0, None, C.byref(h)) == _lib.AFV_ERR_ARG         # no such device
    bad = _lib.AfvParams(-1.0, np.pi, 4.8, 1.0, 1.0, 0.2, 0.45)
    assert lib.afv_create(0, C.byref(bad), 0.0, None, C.byref(h)) == _lib.AFV_ERR_ARG
    sim = afv.FiniteVoronoiSimulator(np.random.default_rng(0).random((50, 2)) * 5, afv.PhysicalParams())
    out = np.zeros((50, 2))
    assert lib.afv_get_forces(sim._h._h, out.ctypes.data) == _lib.AFV_ERR_STATE              # getter before build
    assert b"afv_build" in lib.afv_last_error(sim._h._h)
    k = C.c_int64()
    sim.build(connect=False, rings=False)
    assert lib.afv_get_num_connections(sim._h._h, C.byref(k)) == _lib.AFV_ERR_STATE           # connect was not requested
    with pytest.raises(ValueError, match=r"A0 must be scalar or have shape \(50,\)"):
        sim.update_preferred_areas(np.ones(49))
    with pytest.raises(ValueError, match="noise must have shape"):
        sim.step(0.01, noise=np.zeros((3, 7)), n_steps=3)
    with pytest.raises(AfvError):
        afv.FiniteVoronoiSimulator(np.zeros((3, 2)), afv.PhysicalParams(r=2.0), periodic=5.0)
    sim.close()


def test_update_semantics_follow_the_reference():
    """finite_voronoi.py:1153-1240: N change resets A0; update_params resets A0; update_positions snapshots its input."""
    import pyafv_b200 as afv
    from afv_oracle import oracle_build
    rng = np.random.default_rng(3)
    pts = rng.random((300, 2)) * 14
    A0 = np.pi * (1 + 0.3 * rng.random(300))
    sim = afv.FiniteVoronoiSimulator(pts, afv.PhysicalParams())
    sim.update_preferred_areas(A0)
    d1 = sim.build()
    _close(d1["forces"], oracle_build(pts, A0=A0)["forces"], "forces with A0")
    work = pts.copy()
    sim.update_positions(work)
    work += 100.0                                  # the caller mutates its array after the call (examples/relaxation.py:44-45)
    _close(sim.build()["forces"], d1["forces"], "snapshot at update_positions")
    assert np.array_equal(sim.preferred_areas, A0) and sim.preferred_areas is not sim._preferred_areas
    sim.update_params(sim.phys.replace(KA=2.0))    # resets every A0 to phys.A0
    assert np.allclose(sim.preferred_areas, np.pi)
    _close(sim.build()["forces"], oracle_build(pts, KA=2.0)["forces"], "after update_params")
    sim.update_positions(pts[:120])                # N changes: A0 back to default, results sized to the new N
    d3 = sim.build()
    assert d3["forces"].shape == (120, 2) and len(d3["regions"]) == 120
    _close(d3["forces"], oracle_build(pts[:120], KA=2.0)["forces"], "after N change")
    sim.update_positions(pts, A0)                  # N changes again, A0 given
    _close(sim.build()["forces"], oracle_build(pts, A0=A0, KA=2.0)["forces"], "A0 passed with new N")
    sim.close()


def test_step_matches_host_euler_with_philox_noise():
    """step() == build + host Euler update with the same (host-evaluated) Philox noise, 5 steps, theta included."""
    import ctypes as C
    import pyafv_b200 as afv
    from pyafv_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(8)
    N, L = 500, 30.0
    pts = rng.random((N, 2)) * L
    theta = 2 * np.pi * rng.random(N) - np.pi
    phys = afv.PhysicalParams(seed=11)
    a = afv.FiniteVoronoiSimulator(pts, phys, periodic=L)
    a.step(0.002, 1.0, 1.5, 0.4, theta=theta, n_steps=5)          # seed from phys.seed
    b = afv.FiniteVoronoiSimulator(pts, phys, periodic=L)
    p, th = pts.copy(), theta.copy()
    for s in range(5):
        b.update_positions(p)
        F = b.build(connect=False, rings=False)["forces"]
        xi = np.array([lib.afv_philox_normal(C.c_uint64(11), C.c_int64(i), C.c_int64(s)) for i in range(N)])
        p = np.mod(p + (1.0 * F + 1.5 * np.column_stack((np.cos(th), np.sin(th)))) * 0.002, L)
        th = th + np.sqrt(2 * 0.4 * 0.002) * xi
    d = a.pts - p
    d -= L * np.round(d / L)
    assert np.abs(d).max() < 1e-10 and np.abs(a.theta - th).max() < 1e-12
    a.close(); b.close()


@pytest.mark.parametrize("N,L,periodic", [(0, 10.0, False), (1, 10.0, True), (2, 10.0, True), (30, 4.5, True), (60, 5.9, True),
                                          (200, 8.1, True), (3, 50.0, False)])
def test_tiny_and_small_box_cases(N, L, periodic):
    """Degenerate sizes and periodic boxes with only 2-4 bins per side (minimum-image candidate search)."""
    import pyafv_b200 as afv
    from afv_oracle import oracle_build
    rng = np.random.default_rng(100 + N)
    pts = rng.random((N, 2)) * L
    sim = afv.FiniteVoronoiSimulator(pts, afv.PhysicalParams(), periodic=L if periodic else None)
    d = sim.build()
    assert d["forces"].shape == (N, 2) and d["coord_nums"].shape == (N,) and len(d["regions"]) == N
    if N:
        o = oracle_build(pts, periodic=L if periodic else None)
        for k in ("areas", "perimeters", "arclens", "forces"):
            _close(d[k], o[k], k)
        assert np.array_equal(d["coord_nums"], o["coord_nums"]) and np.array_equal(d["connections"], o["connections"])
    sim.step(0.01, 1.0, 1.0, 0.1, n_steps=2)
    assert sim.pts.shape == (N, 2)
    sim.close()


@pytest.mark.parametrize("dens", [2.5, 4.0])
def test_high_density_uses_longer_candidate_lists(dens):
    """Compressed tissue: 30-50 cells within 2 l of every cell; the candidate-list capacity adapts, nothing is dropped."""
    import pyafv_b200 as afv
    from afv_oracle import oracle_build
    N = 20000
    L = float(np.sqrt(N / dens))
    pts = np.random.default_rng(17).random((N, 2)) * L
    sim = afv.FiniteVoronoiSimulator(pts, afv.PhysicalParams(), periodic=L)
    d = sim.build(rings=False)
    o = oracle_build(pts, periodic=L)
    for k in ("areas", "perimeters", "arclens", "forces"):
        _close(d[k], o[k], k)
    assert np.array_equal(d["coord_nums"], o["coord_nums"]) and np.array_equal(d["connections"], o["connections"])
    assert sim.diagnostics()["slow_path_cells"] < N // 100 and sim.diagnostics()["lookup_misses"] == 0
    sim.close()


def test_long_runs_of_the_baseline_configs_stay_consistent():
    """BASELINE configs [1] and [2] in miniature: an open active blob (free medium interface) and a periodic jammed
    hex sheet with per-cell A0, thousands of fused steps: finite state, no unmatched ring look-ups, net force ~ 0, and
    the relaxation lowers the force scale; the final state still agrees with the CPU oracle."""
    import pyafv_b200 as afv
    from afv_oracle import oracle_build
    rng = np.random.default_rng(42)
    phys = afv.PhysicalParams()
    # config [1]: examples/active_FV.py with N = 1000 (blob of density ~1.8), 200 relaxation + 2000 active steps
    pts = (rng.random((1000, 2)) * 0.3 + 0.35) * 25.0 * np.sqrt(10.0)
    sim = afv.FiniteVoronoiSimulator(pts, phys)
    sim.step(0.01, 1.0, 0.0, 0.0, n_steps=200)
    sim.step(0.01, 1.0, 2.4, 0.3, theta=2 * np.pi * rng.random(1000) - np.pi, seed=42, n_steps=2000)
    p = sim.pts
    assert np.isfinite(p).all() and np.isfinite(sim.theta).all()
    d = sim.build()
    o = oracle_build(p)
    _close(d["forces"], o["forces"], "forces after 2200 steps")
    assert np.array_equal(d["coord_nums"], o["coord_nums"]) and sim.diagnostics()["lookup_misses"] == 0
    assert np.abs(d["forces"].sum(axis=0)).max() < 1e-9 * max(1.0, np.abs(d["forces"]).sum())
    sim.close()
    # config [2]: periodic hex sheet 97 x 112 (N = 10864), a = 1.5197, jitter 0.05, A0 = 2 (1 + 0.1 N(0,1)) >= 0.4
    ncol, nrow, a = 97, 112, 1.5197
    L = ncol * a
    ii, jj = np.meshgrid(np.arange(ncol), np.arange(nrow), indexing="ij")
    pts = np.stack([((ii + 0.5 * (jj % 2)) * a).ravel(), (jj * (L / nrow)).ravel()], axis=1)
    pts = np.mod(pts + 0.05 * np.random.default_rng(3).standard_normal(pts.shape), L)
    A0 = np.maximum(2.0 * (1 + 0.1 * np.random.default_rng(4).standard_normal(len(pts))), 0.4)
    sim = afv.FiniteVoronoiSimulator(pts, phys, periodic=L)
    sim.update_preferred_areas(A0)
    f0 = np.abs(sim.build(connect=False, rings=False)["forces"]).max()
    sim.step(0.01, 1.0, 0.0, 0.0, n_steps=1500)
    d = sim.build(connect=False, rings=False)
    assert np.abs(d["forces"]).max() < 0.2 * f0 and sim.diagnostics()["lookup_misses"] == 0
    o = oracle_build(sim.pts, A0=A0, periodic=L)
    _close(d["forces"], o["forces"], "forces after relaxation")
    _close(d["areas"], o["areas"], "areas after relaxation")
    assert np.array_equal(d["coord_nums"], o["coord_nums"])
    sim.close()


def test_dense_open_blob_grows_candidate_lists():
    """Open boundaries give no density in advance: a blob with ~45 cells within 2 l of each cell overflows the
    initial candidate lists of far more cells than the slow path can queue; the build must grow them and succeed."""
    import pyafv_b200 as afv
    from afv_oracle import oracle_build
    rng = np.random.default_rng(23)
    N = 60000
    pts = rng.random((N, 2)) * np.sqrt(N / 3.5) + 1000.0
    sim = afv.FiniteVoronoiSimulator(pts, afv.PhysicalParams())
    d = sim.build(rings=False)
    o = oracle_build(pts)
    for k in ("areas", "perimeters", "arclens", "forces"):
        _close(d[k], o[k], k)
    assert np.array_equal(d["coord_nums"], o["coord_nums"])
    sim.step(1e-4, 1.0, 0.0, 0.0, n_steps=3)
    assert np.isfinite(sim.pts).all()
    sim.close()


def test_full_size_parity_with_oracle_1e6():
    """BASELINE config [3] at full size: N = 10^6 periodic dense, heterogeneous A0 -- topology exact, FP64 quantities to
    1e-9, against the CPU oracle (which finishes 10^6 cells in a few seconds)."""
    import pyafv_b200 as afv
    from afv_oracle import oracle_build
    N = 1_000_000
    L = float(np.sqrt(N))
    rng = np.random.default_rng(42)
    pts = rng.random((N, 2)) * L
    A0 = np.pi * (1 + 0.1 * rng.standard_normal(N))
    sim = afv.FiniteVoronoiSimulator(pts, afv.PhysicalParams(), periodic=L)
    sim.update_preferred_areas(A0)
    d = sim.build(connect=True, rings=False)
    o = oracle_build(pts, A0=A0, periodic=L)
    for k in ("areas", "perimeters", "arclens", "forces"):
        _close(d[k], o[k], k)
    assert np.array_equal(d["coord_nums"], o["coord_nums"])
    assert np.array_equal(d["connections"], o["connections"])
    assert sim.diagnostics()["lookup_misses"] == 0
    sim.close()


@pytest.mark.parametrize("seed", [0, 1, 2, 3, 4, 5])
def test_random_parameter_sets_match_oracle(seed):
    """Random physical parameters (radius, stiffnesses, tensions of either sign, delta from 0 to r), random densities,
    open and periodic, per-cell A0: every term of the force (area, perimeter, arc tension, delta cut-off) is exercised."""
    import pyafv_b200 as afv
    from afv_oracle import oracle_build
    rng = np.random.default_rng(1000 + seed)
    r = float(rng.uniform(0.5, 2.0))
    kw = dict(r=r, A0=float(rng.uniform(0.3, 1.5) * np.pi * r * r), P0=float(rng.uniform(2.0, 7.0) * r), KA=float(rng.uniform(0.0, 3.0)),
              KP=float(rng.uniform(0.0, 3.0)), Lambda=float(rng.uniform(-0.5, 0.8)), delta=float(rng.uniform(0.0, 1.0) * r))
    N = 8000
    dens = float(rng.choice([0.15, 0.4, 1.0, 1.6])) / (r * r)
    L = float(np.sqrt(N / dens))
    periodic = bool(seed % 2)
    pts = rng.random((N, 2)) * L + (0.0 if periodic else -0.5 * L)
    A0 = kw["A0"] * (1 + 0.3 * rng.random(N))
    sim = afv.FiniteVoronoiSimulator(pts, afv.PhysicalParams(**kw), periodic=L if periodic else None)
    sim.update_preferred_areas(A0)
    d = sim.build()
    okw = dict(kw); okw["A0"] = A0
    o = oracle_build(pts, periodic=L if periodic else None, **okw)
    for k in ("areas", "perimeters", "arclens", "forces"):
        _close(d[k], o[k], k)
    assert np.array_equal(d["coord_nums"], o["coord_nums"]) and np.array_equal(d["connections"], o["connections"])
    assert sim.diagnostics()["lookup_misses"] == 0
    sim.close()


def _degree_from_connections(N, conn):
    deg = np.zeros(N, dtype=np.int64)
    if len(conn):
        np.add.at(deg, conn[:, 0], 1)
        np.add.at(deg, conn[:, 1], 1)
    return deg


def test_exact_square_and_near_square_return(golden_dir):
    """The reference's degenerate inputs (tests/test_geom.py:83-97): an exact square -- four cocircular centres, an
    excluded tie: the reference itself picks a random triple there (finite_voronoi.py:549-552) -- and the near-square
    ``pts_at_error.npy`` of its joggle regression (a square to the last bit or two: the same tie).  Both must return a
    finite result without any retry or crash -- the reference only asks that build() returns there."""
    import pyafv_b200 as afv
    phys = afv.PhysicalParams()
    square = np.array([[1., 1.], [-1., 1.], [-1., -1.], [1., -1.]]) * 0.5
    near = np.load(golden_dir / "pts_at_error.npy")
    assert near.shape == (4, 2)
    for pts in (square, near, near + 100.0, square * 1.7):
        sim = afv.FiniteVoronoiSimulator(pts, phys)
        d = sim.build()
        for k in ("forces", "areas", "perimeters", "arclens"):
            assert np.isfinite(d[k]).all(), k
        assert d["coord_nums"].min() >= 2 and d["coord_nums"].max() <= 3   # the two lattice neighbours (+ the diagonal, at most)
        assert all(len(r) == len(t) for r, t in zip(d["regions"], d["edges_type"]))
        # areas / perimeters do not depend on which diagonal the tie resolves to (the extra edge has length ~0); the
        # four cells may disagree about that zero-length contact, so coord_nums is only pinned up to it
        assert np.ptp(d["areas"]) < 1e-9 and np.ptp(d["perimeters"]) < 1e-9
        deg = _degree_from_connections(len(pts), d["connections"])
        assert np.abs(d["coord_nums"] - deg).max() <= 1
        sim.close()
    # a square that is clearly broken (1e-6) is no tie: unique tessellation, must match the oracle
    from afv_oracle import oracle_build
    broken = square + np.array([[1e-6, 0.0], [0.0, 0.0], [0.0, 0.0], [0.0, -2e-6]])
    sim = afv.FiniteVoronoiSimulator(broken, phys)
    d = sim.build()
    o = oracle_build(broken)
    assert np.array_equal(d["coord_nums"], o["coord_nums"]) and np.array_equal(d["connections"], o["connections"])
    assert np.array_equal(d["coord_nums"], _degree_from_connections(4, d["connections"]))
    for k in ("forces", "areas", "perimeters", "arclens"):
        assert np.abs(d[k] - o[k]).max() <= 1e-9 * max(1.0, np.abs(o[k]).max()), k
    assert sim.diagnostics()["lookup_misses"] == 0
    sim.close()


def test_debug200_coordination_equals_connection_degree(golden_dir):
    """tests/test_geom.py:50-59 of the reference: coord_nums equals the degree in the graph of ``connections``."""
    import pyafv_b200 as afv
    pts = np.load(golden_dir / "debug_pts.npy")
    d = afv.FiniteVoronoiSimulator(pts, afv.PhysicalParams()).build()
    assert np.array_equal(d["coord_nums"], _degree_from_connections(len(pts), d["connections"]))


def test_slow_path_queue_restarts_every_step_of_one_call():
    """2000 steps in ONE afv_step call on a workload that queues a few cells for the slow path in most steps (dense
    random: long rings and 15-gons): the queue must restart every step -- a queue that kept growing would re-run stale
    entries, overflow after a few hundred steps and end in a spurious AFV_ERR_CAPACITY -- and the final state must
    still be the oracle's."""
    import pyafv_b200 as afv
    from afv_oracle import oracle_build
    N = 200_000
    L = float(np.sqrt(N))
    rng = np.random.default_rng(77)
    pts = rng.random((N, 2)) * L
    phys = afv.PhysicalParams()
    sim = afv.FiniteVoronoiSimulator(pts, phys, periodic=L)
    A0 = np.pi * (1 + 0.1 * rng.standard_normal(N))
    sim.update_preferred_areas(A0)
    queued = 0
    for _ in range(4):                                          # the first steps of a random start are the busiest
        sim.step(0.01, 1.0, 2.4, 0.3, seed=3, n_steps=1)
        queued += sim.diagnostics()["slow_path_cells"]
    sim.step(0.01, 1.0, 2.4, 0.3, seed=3, n_steps=2000)         # one call
    diag = sim.diagnostics()
    assert diag["slow_path_cells"] <= 64 and diag["lookup_misses"] == 0, diag     # the count of the LAST step, not a running total
    d = sim.build(connect=False, rings=False)
    o = oracle_build(sim.pts, A0=A0, periodic=L)
    for k in ("forces", "areas", "perimeters", "arclens"):
        assert np.abs(d[k] - o[k]).max() <= 1e-9 * max(1.0, np.abs(o[k]).max()), k
    assert np.array_equal(d["coord_nums"], o["coord_nums"])
    sim.close()


def test_capacity_overflow_does_not_move_the_cells():
    """A step whose geometry exceeds every capacity reports AFV_ERR_CAPACITY and leaves the positions where they were
    (the error is about that state), instead of advancing the overflowed cells with zero force."""
    import pyafv_b200 as afv
    from pyafv_b200._lib import AfvError
    phys = afv.PhysicalParams()
    # 130 cells on a small circle around a centre cell: more than the 96 polygon edges the slow path holds
    n = 130
    ang = 2 * np.pi * np.arange(n) / n
    pts = np.vstack([[0.0, 0.0], 1.2 * np.column_stack((np.cos(ang), np.sin(ang)))]) + 50.0
    sim = afv.FiniteVoronoiSimulator(pts, phys, periodic=100.0)
    with pytest.raises(AfvError) as ei:
        sim.step(0.01, 1.0, 0.0, 0.0, n_steps=3)
    assert ei.value.code == -3
    assert np.array_equal(sim.pts, pts)
    sim.close()
