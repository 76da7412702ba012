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
    # the result must not depend on the number of GPUs -- bit for bit.  Every rank holds the box's own
    # coordinates, displacements are minimum images of the same operands, vertices and sums are taken in an order the
    # geometry defines (not the sort), so the slabs reproduce the single handle exactly.
    assert np.array_equal(b["pts"], pts)
    for k in ("forces", "areas", "perimeters", "arclens", "coord_nums"):
        assert np.array_equal(a[k], b[k]), (k, np.abs(a[k] - b[k]).max())
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
    assert np.array_equal(b["pts"], ref.pts)                   # bit-identical trajectories, whatever the number of slabs
    assert np.array_equal(b["theta"], ref.theta)
    assert sum(s.n_owned for s in sh.ranks) == len(pts)
    ref.close(); sh.close()


@pytest.mark.parametrize("world", [2, 3, 4])
def test_late_cell_exchange_equals_plain(world):
    """Exchange overlapped with the interior geometry (late cells, afv_late_*): the same trajectory as exchange-then-step,
    from an UNRELAXED random start (single cells jump many radii per step, so received transfers land beyond the face
    columns now and then and the full-pass fallback runs too); mixed with plain steps and builds in between."""
    import pyafv_b200 as afv
    from pyafv_b200.sharded import LoopbackShardedSimulator
    pts, theta, A0, L = _setup(12000, 1.0, 31)
    phys = afv.PhysicalParams()
    out = []
    for overlap in (False, True):
        sh = LoopbackShardedSimulator(pts, theta, phys, L, world, A0=A0, overlap=overlap)
        sh.step(0.01, 1.0, 2.4, 0.3, seed=9, n_steps=25)
        mid = sh.build()                                        # leaves the late loop (pending messages are consumed)
        assert sum(s.n_owned for s in sh.ranks) == len(pts)
        sh.step(0.01, 1.0, 2.4, 0.3, seed=9, n_steps=15)       # and re-enters it
        assert all(s.overlapped_steps == (40 if overlap else 0) for s in sh.ranks)
        if overlap:
            stats = [s.exchange_stats() for s in sh.ranks]
            assert all(st["late_steps"] == 40 and st["late_cells"] > 0 for st in stats)
        out.append((mid, sh.build()))
        sh.close()
    # the two runs meet the neighbours of a cell in different orders (row-major against column-major bins, late cells
    # behind the sorted ones); nothing may depend on that: the trajectories are bit-identical even though this violent
    # dynamics would amplify a last-bit difference to O(1) within the 40 steps
    for a, b in zip(out[0], out[1]):
        for k in ("gids", "pts", "theta", "forces", "areas", "perimeters", "arclens", "coord_nums"):
            assert np.array_equal(a[k], b[k]), (k, np.abs(a[k] - b[k]).max())


def test_late_cells_beyond_the_face_columns_trigger_the_full_face_pass():
    """Every cell drifts 5.5 radii in +x per step: the transfers that started next to the face land beyond the receiver's
    face columns (4.8 radii here), the sender flags them and the receiver's face pass covers every cell.  A rigid
    translation leaves the geometry unchanged."""
    import pyafv_b200 as afv
    from pyafv_b200.sharded import LoopbackShardedSimulator
    pts, theta, A0, L = _setup(12000, 1.0, 5)
    phys = afv.PhysicalParams()
    ref = afv.FiniteVoronoiSimulator(pts, phys, periodic=L)
    ref.update_preferred_areas(A0)
    a = ref.build(connect=False, rings=False)
    sh = LoopbackShardedSimulator(pts, np.zeros(len(pts)), phys, L, 3, A0=A0, overlap=True)
    sh.step(1.0, 0.0, 5.5, 0.0, seed=1, n_steps=4)              # mu = 0: pure drift
    stats = [s.exchange_stats() for s in sh.ranks]
    assert all(st["full_face_passes"] >= 3 for st in stats), stats
    b = sh.build()
    d = b["pts"] - (pts + [22.0, 0.0])
    d -= L * np.round(d / L)
    assert np.abs(d).max() <= 1e-9
    for k in ("forces", "areas", "perimeters", "arclens"):
        scale = max(1.0, np.abs(a[k]).max())
        assert np.abs(a[k] - b[k]).max() <= 1e-9 * scale, k
    assert np.array_equal(a["coord_nums"], b["coord_nums"])
    ref.close(); sh.close()


def test_late_cell_exchange_follows_single_gpu():
    import pyafv_b200 as afv
    from pyafv_b200.sharded import LoopbackShardedSimulator
    pts, theta, A0, L = _setup(6000, 0.4, 12)
    phys = afv.PhysicalParams()
    ref = afv.FiniteVoronoiSimulator(pts, phys, periodic=L)
    ref.step(0.01, 1.0, 0.0, 0.0, n_steps=200)
    pts = ref.pts.copy()
    ref.step(0.01, 1.0, 2.4, 0.3, theta=theta, seed=5, n_steps=30)
    sh = LoopbackShardedSimulator(pts, theta, phys, L, 4, overlap=True)
    for s in sh.ranks:
        s._step_index = 200
    sh.step(0.01, 1.0, 2.4, 0.3, seed=5, n_steps=30)
    assert all(s.overlapped_steps == 30 for s in sh.ranks)
    b = sh.build()
    assert np.array_equal(b["pts"], ref.pts) and np.array_equal(b["theta"], ref.theta)
    ref.close(); sh.close()


def test_peer_memory_exchange_across_two_processes():
    """One process per GPU: the pack kernel writes into the neighbours' inboxes over NVLink (CUDA IPC + flag words,
    ``NvlinkPeerComm``) -- same bits as NCCL send / receive and as one handle, plain and overlapped.  Needs two GPUs."""
    import subprocess
    import sys
    from pathlib import Path
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    worker = Path(__file__).with_name("_peer_exchange_worker.py")
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                          "--master-port", "29533", str(worker)], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0 and "PEER-EXCHANGE-OK" in res.stdout, res.stdout[-2000:] + res.stderr[-4000:]
