#!/usr/bin/env python
"""Small run through every kernel path (for compute-sanitizer): open / periodic / slow path / exports / sharded."""
import sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import pyafv_b200 as afv
from pyafv_b200.sharded import LoopbackShardedSimulator

rng = np.random.default_rng(0)
phys = afv.PhysicalParams()
# open, with a many-neighbour cell (slow path + overflow pool)
ang = 2 * np.pi * np.arange(40) / 40
pts = np.vstack([[0, 0], 0.9 * np.column_stack((np.cos(ang), np.sin(ang))), (rng.random((500, 2)) - 0.5) * 40]) + 30
sim = afv.FiniteVoronoiSimulator(pts, phys)
d = sim.build()
sim.step(0.001, 1.0, 1.0, 0.1, n_steps=3)
print("open", sim.diagnostics(), len(d["connections"]))
sim.close()
# periodic dense + sparse
for dens in (1.0, 0.1, 3.0):
    N = 20000
    L = float(np.sqrt(N / dens))
    sim = afv.FiniteVoronoiSimulator(rng.random((N, 2)) * L, phys, periodic=L)
    sim.update_preferred_areas(np.pi * (1 + 0.1 * rng.random(N)))
    d = sim.build()
    sim.step(0.005, 1.0, 2.4, 0.3, theta=rng.random(N), n_steps=5)
    sim.update_positions(sim.pts)
    sim.build(connect=False, rings=False)
    print("periodic", dens, sim.diagnostics())
    sim.close()
# sharded loopback
N, L = 20000, float(np.sqrt(20000 / 0.5))
sh = LoopbackShardedSimulator(rng.random((N, 2)) * L, rng.random(N), phys, L, 3)
sh.step(0.01, 1.0, 2.4, 0.3, seed=1, n_steps=5)
b = sh.build()
print("sharded", len(b["gids"]), [s.n_owned for s in sh.ranks])
sh.close()
