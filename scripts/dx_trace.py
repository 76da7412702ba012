#!/usr/bin/env python
"""Largest |dx| per step of the bench workload (uniform random start, rho l^2 = 1), from a 2-slab loopback run on one GPU:
the time series that decides when the overlapped exchange may engage.  python scripts/dx_trace.py [N] [steps]"""
import sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import pyafv_b200 as afv
from pyafv_b200.sharded import LoopbackShardedSimulator

N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 200
rng = np.random.default_rng(0)
L = float(np.sqrt(N))
sh = LoopbackShardedSimulator(rng.random((N, 2)) * L, 2 * np.pi * rng.random(N), afv.PhysicalParams(), L, 2, overlap=False)
series = []
for k in range(steps):
    sh.step(0.01, 1.0, 2.4, 0.3, seed=7, n_steps=1)
    sh._refresh_halo()                       # the exchange reads the displacement slots
    series.append(max(s._dx_hist[-1] for s in sh.ranks))
print("N", N, "max |dx| per step (value read = max of the last two steps):")
print(" ".join("%.2f" % v for v in series))
