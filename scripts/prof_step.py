#!/usr/bin/env python
"""Minimal driver for ncu captures: N periodic dense cells, a few fused steps, nothing else."""
import sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import pyafv_b200 as afv

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
dens = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
L = float(np.sqrt(N / dens))
rng = np.random.default_rng(42)
sim = afv.FiniteVoronoiSimulator(rng.random((N, 2)) * L, afv.PhysicalParams(), periodic=L)
sim.step(0.01, 1.0, 2.4, 0.3, theta=2 * np.pi * rng.random(N) - np.pi, seed=7, n_steps=steps)
print("done", N, steps, sim.diagnostics())
