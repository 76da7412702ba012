"""GPU: slab sharding.  All ranks of a world of 2..4 slabs run inside one process on one GPU (LoopbackSharded-
Simulator: same phases, same C-ABI primitives and kernels as the one-process-per-GPU run) and must reproduce the
single-handle periodic result: the reference's own parallel == serial criterion (tests/test_parallel.py:112-176)."""
from __future__ import annotations

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _setup(N, dens, seed):
    rng = np.random.default_rng(seed)
    L = float(np.sqrt(N / dens))
    pts = rng.random((N, 2)) * L
    theta = 2 * np.pi * rng.random(N) - np.pi
    A0 = np.pi * (1 + 0.1 * rng.standard_normal(N))
    return pts, theta, A0, L


@pytest.mark.parametrize("world", [2, 3, 4])
@pytest.mark.parametrize("dens", [1.0, 0.2])
def test_sharded_build_equals_single(world, dens):
    import pyafv_b200 as afv
    from pyafv_b200.sharded import LoopbackShardedSimulator
    pts, theta, A0, L = _setup(20000, dens, 11)
    phys = afv.PhysicalParams()
    ref = afv.FiniteVoronoiSimulator(pts, phys, periodic=L)
    ref.update_preferred_areas(A0)
    a = ref.build(connect=False, rings=False)
    sh = LoopbackShardedSimulator(pts, theta, phys, L, world, A0=A0)
    b = sh.build()
    assert np.array_equal(b["gids"], np.arange(len(pts)))
    for k in ("forces", "areas", "perimeters", "arclens"):
        scale = max(1.0, np.abs(a[k]).max())
        assert np.abs(a[k] - b[k]).max() <= 1e-9 * scale, k
    assert np.array_equal(a["coord_nums"], b["coord_nums"])
    ref.close(); sh.close()


def test_sharded_trajectory_equals_single():
    """30 active steps with the Philox stream keyed by global cell id: the sharded run follows the single-GPU run
    (cells migrate between slabs on the way)."""
    import pyafv_b200 as afv
    from pyafv_b200.sharded import LoopbackShardedSimulator
    pts, theta, A0, L = _setup(6000, 0.4, 12)
    phys = afv.PhysicalParams()
    ref = afv.FiniteVoronoiSimulator(pts, phys, periodic=L)
    ref.step(0.01, 1.0, 0.0, 0.0, n_steps=200)              # relax first (see tests/golden/make_golden.py)
    pts = ref.pts.copy()
    ref.step(0.01, 1.0, 2.4, 0.3, theta=theta, seed=5, n_steps=30)
    sh = LoopbackShardedSimulator(pts, theta, phys, L, 3)
    for s in sh.ranks:
        s._step_index = 200                                    # same position in the noise stream as the reference run
    sh.step(0.01, 1.0, 2.4, 0.3, seed=5, n_steps=30)
    b = sh.build()
    d = b["pts"] - ref.pts
    d -= L * np.round(d / L)
    assert np.abs(d).max() <= 1e-8
    assert np.abs(b["theta"] - ref.theta).max() <= 1e-10
    assert sum(s.n_owned for s in sh.ranks) == len(pts)
    ref.close(); sh.close()
