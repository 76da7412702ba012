#!/usr/bin/env python
"""Stage times of the step loop (ms per step) for the library selected by AFV_B200_LIB: N dens steps [jammed]
"jammed": per-cell A0 (10 % spread), 500 relaxation steps first (v0 = 0), then timed passive relaxation steps
(BASELINE config[2] shape: jammed tissue with heterogeneous A0, energy minimisation)."""
import sys, os
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import pyafv_b200 as afv
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
dens = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 30
L = float(np.sqrt(N / dens))
rng = np.random.default_rng(42)
jammed = len(sys.argv) > 4 and sys.argv[4] == "jammed"
sim = afv.FiniteVoronoiSimulator(rng.random((N, 2)) * L, afv.PhysicalParams(), periodic=L)
v0, Dr = (0.0, 0.0) if jammed else (2.4, 0.3)
if jammed:
    sim.update_preferred_areas(np.pi * (1.0 + 0.1 * rng.standard_normal(N)))
    sim.step(0.01, 1.0, 0.0, 0.0, n_steps=500)
sim.step(0.01, 1.0, v0, Dr, theta=2 * np.pi * rng.random(N) - np.pi, seed=7, n_steps=5)
h = sim._h
h.call("afv_set_profiling", 1)
ms, l = np.zeros(6), np.zeros(6, dtype=np.int64)
h.call("afv_get_stage_times", ms.ctypes.data, l.ctypes.data, 1)
sim.step(0.01, 1.0, v0, Dr, seed=7, n_steps=steps)
h.call("afv_get_stage_times", ms.ctypes.data, l.ctypes.data, 1)
print(os.environ.get("AFV_B200_LIB", "default"), "jammed" if jammed else "active", "N", N, "dens", dens, "sort %.3f hull %.3f ring %.3f force %.3f slow %.3f total %.3f ms/step" % (ms[0]/steps, ms[1]/steps, ms[2]/steps, ms[3]/steps, ms[5]/steps, ms.sum()/steps), sim.diagnostics())
