"""GPU: the reference's parallel API (pyafv/parallel_voronoi.py:421-720) driving slabs on the GPUs of the node --
``grid_shape = (nx, ny)`` -> nx*ny x-slabs, dealt to the visible GPUs (all on one GPU when only one is visible).  The
criterion is the reference's own (tests/test_parallel.py:112-176): parallel == serial."""
from __future__ import annotations

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _rows_sorted(c):
    return c[np.lexsort((c[:, 1], c[:, 0]))] if len(c) else c.reshape(0, 2)


def test_parallel_periodic_slabs_equal_the_single_handle_bit_for_bit():
    import pyafv_b200 as afv
    N = 20000
    L = float(np.sqrt(N))
    rng = np.random.default_rng(21)
    pts, A0 = rng.random((N, 2)) * L, np.pi * (1 + 0.1 * rng.standard_normal(N))
    phys = afv.PhysicalParams()
    ser = afv.FiniteVoronoiSimulator(pts, phys, periodic=L)
    ser.update_preferred_areas(A0)
    a = ser.build(connect=True, rings=False)
    with afv.ParallelFiniteVoronoiSimulator(pts, phys, grid_shape=(4, 1), n_workers=4, periodic=L) as par:
        par.update_preferred_areas(A0)
        assert par.n_slabs == 4 and np.array_equal(par.preferred_areas, A0)
        b = par.build(connect=True)
        assert set(b) == {"forces", "areas", "perimeters", "arclens", "coord_nums", "connections", "plot_mode", "pids"}
        for k in ("forces", "areas", "perimeters", "arclens", "coord_nums"):
            assert np.array_equal(a[k], b[k]), k
        assert np.array_equal(_rows_sorted(b["connections"]), a["connections"]) and len(b["pids"]) == 4
        assert par.build()["connections"].shape == (0, 2)              # connect=False is the default here, unlike the serial class
        # a fused time loop: same trajectory as the single handle, whatever the number of slabs
        theta = 2 * np.pi * rng.random(N) - np.pi
        ser.step(0.01, 1.0, 2.4, 0.3, theta=theta, seed=4, n_steps=12)
        par.step(0.01, 1.0, 2.4, 0.3, theta=theta, seed=4, n_steps=12)
        assert np.array_equal(par.pts, ser.pts) and np.array_equal(par.theta, ser.theta)
        # the reference-style host loop on top of it
        p2 = np.mod(par.pts + 0.01 * par.build()["forces"], L)
        par.update_positions(p2)
        ser.update_positions(p2)
        assert np.array_equal(par.build()["forces"], ser.build(connect=False, rings=False)["forces"])
        d = par.build(connect=False, plot_mode=True)
        assert d["plot_mode"] is True and len(d["diag_plot"]) == 1 and len(d["owned_global_ids"][0]) == N
    ser.close()


def test_parallel_open_boundary_grid_equals_serial():
    """tests/test_parallel.py:112-176 with a real grid: an open-boundary blob cut into 2 x 2 = 4 slabs."""
    import pyafv_b200 as afv
    rng = np.random.default_rng(123)
    pts = rng.random((6000, 2)) * 70.0 - 20.0
    phys = afv.PhysicalParams()
    ser = afv.FiniteVoronoiSimulator(pts, phys)
    a = ser.build(connect=True, rings=False)
    par = afv.ParallelFiniteVoronoiSimulator(pts, phys, grid_shape=(2, 2), n_workers=1, decomposition_method="binned")
    assert par.n_slabs == 4
    b = par.build(connect=True)
    for k in ("forces", "areas", "perimeters", "arclens"):
        assert np.abs(a[k] - b[k]).max() <= 1e-10 * max(1.0, np.abs(a[k]).max()), k      # (the slabs see shifted coordinates)
    assert np.array_equal(a["coord_nums"], b["coord_nums"]) and np.array_equal(_rows_sorted(b["connections"]), a["connections"])
    assert np.allclose(par.pts, pts, rtol=0, atol=1e-12)
    # N changes: preferred areas reset, like the reference (finite_voronoi.py:1177-1182)
    par.update_preferred_areas(2.0)
    par.update_positions(pts[:3000])
    assert par.N == 3000 and np.all(par.preferred_areas == phys.A0)
    ser.update_positions(pts[:3000])
    assert np.abs(par.build()["forces"] - ser.build(connect=False, rings=False)["forces"]).max() <= 1e-9
    # the cloud outgrows its box: the box is rebuilt, results stay right
    far = pts[:3000] * 3.0
    par.update_positions(far); ser.update_positions(far)
    assert np.abs(par.build()["areas"] - ser.build(connect=False, rings=False)["areas"]).max() <= 1e-10
    par.update_params(afv.PhysicalParams(r=0.8, A0=2.0))
    ser.update_params(afv.PhysicalParams(r=0.8, A0=2.0))
    assert par.halo_width == pytest.approx(4.01 * 0.8) and np.all(par.preferred_areas == 2.0)
    assert np.abs(par.build()["forces"] - ser.build(connect=False, rings=False)["forces"]).max() <= 1e-9
    par.close(); ser.close()


def test_parallel_slab_count_follows_the_box():
    import pyafv_b200 as afv
    phys = afv.PhysicalParams()
    pts = np.random.default_rng(1).random((50, 2)) * 10.0
    one = afv.ParallelFiniteVoronoiSimulator(pts, phys, grid_shape=(1, 1), n_workers=1)
    assert one.n_slabs == 1
    small = afv.ParallelFiniteVoronoiSimulator(pts, phys, grid_shape=(8, 1))      # a 10 l blob cannot hold 8 slabs with a 4 l halo
    assert 1 <= small.n_slabs < 8
    ser = afv.FiniteVoronoiSimulator(pts, phys)
    for sim in (one, small):
        assert np.abs(sim.build()["forces"] - ser.build(connect=False)["forces"]).max() <= 1e-10
        sim.close()
    with pytest.raises(ValueError):
        afv.ParallelFiniteVoronoiSimulator(pts, phys, backend="python")
    ser.close()


def test_parallel_uses_every_visible_gpu():
    """With several GPUs visible every slab runs on its own device (gpurun --gpus 2 ...); same results."""
    import torch
    import pyafv_b200 as afv
    n_dev = torch.cuda.device_count()
    if n_dev < 2:
        pytest.skip("needs at least two visible GPUs")
    N = 40000
    L = float(np.sqrt(N))
    rng = np.random.default_rng(3)
    pts, theta = rng.random((N, 2)) * L, 2 * np.pi * rng.random(N) - np.pi
    phys = afv.PhysicalParams()
    ser = afv.FiniteVoronoiSimulator(pts, phys, periodic=L)
    par = afv.ParallelFiniteVoronoiSimulator(pts, phys, grid_shape=(n_dev, 1), periodic=L)
    d = par.build(connect=True)
    assert sorted(set(d["pids"].tolist())) == list(range(n_dev))
    a = ser.build(connect=True, rings=False)
    assert np.array_equal(d["forces"], a["forces"]) and np.array_equal(_rows_sorted(d["connections"]), a["connections"])
    ser.step(0.01, 1.0, 2.4, 0.3, theta=theta, seed=2, n_steps=10)
    par.step(0.01, 1.0, 2.4, 0.3, theta=theta, seed=2, n_steps=10)
    assert np.array_equal(par.pts, ser.pts)
    par.close(); ser.close()
