#!/usr/bin/env python
"""Generate the committed golden fixtures ``tests/golden/ref_*.npz`` from the UNMODIFIED reference.

Run in the build container only (needs /root/reference, or a previous ``oracle/build_ref.py`` build):

    python tests/golden/make_golden.py

Each fixture holds the inputs and the reference's ``FiniteVoronoiSimulator.build()`` outputs
(finite_voronoi.py:935-945) for one case: forces, areas, perimeters, arclens, coord_nums, connections
(sorted, unique), and the ragged rings flattened as ``ring_off`` / ``edges_type`` / ``ring_xy`` (the
coordinates ``vertices[regions[i]]`` -- vertex ids themselves are implementation-defined).
Periodic cases go through the reference's own ``tile_pbc`` (_connectutil.py:12-96) and keep rows [:N].
Trajectory cases replay ``examples/active_FV.py:43-54`` with a recorded noise array.

The CSV / NPY inputs next to this script (init_pts.csv, init_forces.csv, varying_A0.csv,
varying_A0_forces.csv, init_seed407_pts.csv, init_seed46_pts.csv, debug_pts.npy, pts_at_error.npy) are
verbatim copies of the reference's test data (tests/data/), i.e. the vectors its own tests pin.
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parents[1] / "oracle"))
import build_ref  # noqa: E402

build_ref.build()
afv = build_ref.import_reference()


def pack(diag, n=None):
    n = len(diag["areas"]) if n is None else n
    conn = np.asarray(diag["connections"], dtype=np.int64).reshape(-1, 2)
    conn = np.unique(np.sort(conn, axis=1), axis=0) if conn.size else conn
    conn = conn[np.all(conn < n, axis=1)] if conn.size else conn
    off = np.zeros(n + 1, dtype=np.int64)
    et, xy = [], []
    V = np.asarray(diag["vertices"])
    for i in range(n):
        reg = diag["regions"][i]
        off[i + 1] = off[i] + len(reg)
        et.extend(diag["edges_type"][i])
        if len(reg):
            xy.append(V[np.asarray(reg, dtype=int)])
    return dict(forces=diag["forces"][:n], areas=diag["areas"][:n], perimeters=diag["perimeters"][:n],
                arclens=diag["arclens"][:n], coord_nums=np.asarray(diag["coord_nums"][:n], dtype=np.int64),
                connections=conn, ring_off=off, edges_type=np.asarray(et, dtype=np.int8),
                ring_xy=np.concatenate(xy) if xy else np.zeros((0, 2)))


def phys_dict(ph):
    return dict(r=ph.r, A0=ph.A0, P0=ph.P0, KA=ph.KA, KP=ph.KP, Lambda=ph.Lambda, delta=ph.delta)


def save(name, pts, ph, diag, A0=None, L=0.0, n=None, **extra):
    d = pack(diag, n)
    d.update(pts=np.asarray(pts, dtype=float), phys=np.array([ph.r, ph.A0, ph.P0, ph.KA, ph.KP, ph.Lambda, ph.delta]),
             L=np.float64(L), **extra)
    if A0 is not None:
        d["A0"] = np.asarray(A0, dtype=float)
    np.savez_compressed(HERE / f"ref_{name}.npz", **d)
    print(f"ref_{name}.npz  N={len(d['areas'])}  max|F|={np.abs(d['forces']).max():.3g}")


def open_case(name, pts, ph, A0=None):
    sim = afv.FiniteVoronoiSimulator(pts, ph)
    if A0 is not None:
        sim.update_preferred_areas(A0)
    save(name, pts, ph, sim.build(), A0=A0)


def pbc_case(name, pts, L, ph, A0=None):
    N = len(pts)
    tiled, idx = afv.tile_pbc(pts, L, ph.r)
    sim = afv.FiniteVoronoiSimulator(tiled, ph)
    if A0 is not None:
        sim.update_preferred_areas(np.asarray(A0)[idx])
    diag = sim.build()
    d = pack(diag, N)
    # connections through images: map both ends back to original ids
    conn = np.asarray(diag["connections"], dtype=np.int64).reshape(-1, 2)
    conn = conn[(conn[:, 0] < N) | (conn[:, 1] < N)]
    conn = np.unique(np.sort(idx[conn], axis=1), axis=0)
    d["connections"] = conn
    d.update(pts=np.asarray(pts, dtype=float), phys=np.array([ph.r, ph.A0, ph.P0, ph.KA, ph.KP, ph.Lambda, ph.delta]),
             L=np.float64(L))
    if A0 is not None:
        d["A0"] = np.asarray(A0, dtype=float)
    np.savez_compressed(HERE / f"ref_{name}.npz", **d)
    print(f"ref_{name}.npz  N={N} (tiled {len(tiled)})  max|F|={np.abs(d['forces']).max():.3g}")


def hex_jitter(ncol, nrow, a, sigma, seed):
    """Hex lattice ncol x nrow in a square box (y rescaled), Gaussian jitter."""
    rng = np.random.default_rng(seed)
    L = ncol * a
    ii, jj = np.meshgrid(np.arange(ncol), np.arange(nrow), indexing="ij")
    x = (ii + 0.5 * (jj % 2)) * a
    y = jj * (L / nrow)
    pts = np.stack([x.ravel(), y.ravel()], axis=1) + sigma * rng.standard_normal((ncol * nrow, 2))
    return np.mod(pts, L), L


def ref_forces(pts, ph, L, A0=None):
    """Forces of the reference on pts (periodic through its own tile_pbc when L > 0)."""
    N = len(pts)
    if L > 0:
        tiled, idx = afv.tile_pbc(pts, L, ph.r)
        sim = afv.FiniteVoronoiSimulator(tiled, ph)
        if A0 is not None:
            sim.update_preferred_areas(np.asarray(A0)[idx])
        return sim.build(connect=False)["forces"][:N]
    sim = afv.FiniteVoronoiSimulator(pts, ph)
    if A0 is not None:
        sim.update_preferred_areas(A0)
    return sim.build(connect=False)["forces"]


def trajectory(name, pts_init, ph, L, n_relax, n_steps, mu, v0, Dr, dt, seed, A0=None):
    """examples/active_FV.py:37-54: relax to mechanical equilibrium first (v0 = 0), then record n_steps active
    steps with a recorded noise stream.  (Without the relaxation the explicit Euler map is violently unstable --
    |F| dt ~ 2 l -- and even two builds that agree to 1e-14 diverge to O(1) within 50 steps.)"""
    rng = np.random.default_rng(seed)
    N = len(pts_init)
    pts = pts_init.copy()
    for _ in range(n_relax):
        pts = pts + mu * ref_forces(pts, ph, L, A0) * dt
        if L > 0:
            pts = np.mod(pts, L)
    pts0 = pts.copy()
    theta0 = 2 * np.pi * rng.random(N) - np.pi
    noise = rng.standard_normal((n_steps, N))
    theta = theta0.copy()
    snaps = {}
    for s in range(n_steps):
        F = ref_forces(pts, ph, L, A0)
        pts = pts + (mu * F + v0 * np.column_stack((np.cos(theta), np.sin(theta)))) * dt
        theta = theta + np.sqrt(2 * Dr * dt) * noise[s]
        if L > 0:
            pts = np.mod(pts, L)
        if s + 1 in (1, 10, 50, n_steps):
            snaps[f"pts_{s + 1}"] = pts.copy()
    d = dict(pts0=pts0, theta0=theta0, noise=noise, theta_final=theta, L=np.float64(L),
             phys=np.array([ph.r, ph.A0, ph.P0, ph.KA, ph.KP, ph.Lambda, ph.delta]),
             step=np.array([mu, v0, Dr, dt]), **snaps)
    if A0 is not None:
        d["A0"] = np.asarray(A0, dtype=float)
    np.savez_compressed(HERE / f"ref_{name}.npz", **d)
    print(f"ref_{name}.npz  N={N} relax={n_relax} steps={n_steps}  max|F0|={np.abs(ref_forces(pts0, ph, L, A0)).max():.3g}")


def static_cases(P):
    init = np.loadtxt(HERE / "init_pts.csv", delimiter=",")
    varA0 = np.loadtxt(HERE / "varying_A0.csv", delimiter=",")
    open_case("init_d0", init, P(delta=0.0))
    open_case("init_d045", init, P())
    open_case("varyA0_d0", init, P(delta=0.0), A0=varA0)
    open_case("varyA0_d045", init, P(), A0=varA0)
    open_case("seed407", np.loadtxt(HERE / "init_seed407_pts.csv", delimiter=","), P())
    open_case("seed46", np.loadtxt(HERE / "init_seed46_pts.csv", delimiter=","), P())
    open_case("debug200", np.load(HERE / "debug_pts.npy"), P())
    # small clusters, tests/test_geom.py:13-35,83-85
    open_case("n1", np.array([[0.0, 0.0]]), P())
    open_case("n2_touch", np.array([[0.5, 0.5], [0.54, 0.52]]) * 25.0, P())
    open_case("n2_apart", np.array([[0.0, 0.0], [1.0, 10.0]]), P())
    open_case("n3_collinear", np.array([[0.0, 0.0], [1.0, 0.0], [2.0, 0.0]]), P())
    open_case("right_angle", np.array([[-1.0, 2.0], [-1.0, -1.0], [1.0, -1.0]]) * 0.5, P())
    # other parameter sets
    open_case("init_params2", init, P(r=1.3, A0=2.0, P0=3.5, KA=0.7, KP=1.9, Lambda=-0.3, delta=0.2))
    # random clouds at three densities
    for tag, N, dens, seed in (("rand_dense", 2000, 1.0, 11), ("rand_mid", 2000, 1.0 / np.pi, 12),
                               ("rand_sparse", 2000, 0.1, 13), ("rand_squeezed", 1500, 1.8, 14)):
        rng = np.random.default_rng(seed)
        open_case(tag, rng.random((N, 2)) * np.sqrt(N / dens), P())
    # periodic through the reference's tile_pbc
    rng = np.random.default_rng(21)
    N = 1600
    pbc_case("pbc_dense", rng.random((N, 2)) * np.sqrt(N), np.sqrt(N), P())
    rng = np.random.default_rng(22)
    L = np.sqrt(1200 / 0.1)
    pbc_case("pbc_sparse", rng.random((1200, 2)) * L, L, P())
    pts, L = hex_jitter(30, 34, 1.5197, 0.05, 3)
    rng = np.random.default_rng(4)
    A0 = np.maximum(2.0 * (1 + 0.1 * rng.standard_normal(len(pts))), 0.4)
    pbc_case("pbc_hex_varyA0", pts, L, P(), A0=A0)


def main(only=None):
    P = afv.PhysicalParams
    if only is None:
        static_cases(P)
    # trajectories (examples/active_FV.py parameters: mu=1, v0=2.4, Dr=0.3, dt=0.01), relaxed first (:37-41)
    if only is None or "traj" in only:
        rng = np.random.default_rng(31)
        trajectory("traj_open", (rng.random((150, 2)) * 0.3 + 0.35) * 25.0 * np.sqrt(1.5), P(), 0.0, 300, 100, 1.0, 2.4, 0.3, 0.01, 32)
        rng = np.random.default_rng(33)
        trajectory("traj_pbc", rng.random((400, 2)) * 30.0, P(), 30.0, 300, 100, 1.0, 2.4, 0.3, 0.01, 34)


if __name__ == "__main__":
    main(only=sys.argv[1:] or None)
