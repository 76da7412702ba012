#!/usr/bin/env python
"""Loopback timing of exchange-then-step against the late-cell form (2 slabs on one GPU): python scripts/late_probe.py [N] [steps]"""
import sys, time
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import pyafv_b200 as afv
from pyafv_b200.sharded import LoopbackShardedSimulator
N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 2_000_000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 50
rng = np.random.default_rng(0)
L = float(np.sqrt(N))
pts, th = rng.random((N, 2)) * L, 2 * np.pi * rng.random(N)
for overlap in (False, True, False, True):
    sh = LoopbackShardedSimulator(pts, th, afv.PhysicalParams(), L, 2, overlap=overlap)
    sh.step(0.01, 1.0, 2.4, 0.3, seed=7, n_steps=10)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    sh.step(0.01, 1.0, 2.4, 0.3, seed=7, n_steps=steps)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(overlap, "%.3f ms per step of both slabs" % (1e3 * dt / steps), [s.exchange_stats() for s in sh.ranks] if overlap else "")
    sh.close()
