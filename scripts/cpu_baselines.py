#!/usr/bin/env python
"""CPU baselines of the unmodified reference on this box: serial Cython build + Euler step for
N = 1e2..1e6 (periodic via tile_pbc), and ParallelFiniteVoronoiSimulator over all cores for N = 1e5, 1e6."""
import json, os, sys, time
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "oracle"))
import bench

out = {"cores": os.cpu_count(), "serial": {}, "parallel": {}}
for n in (100, 1000, 10_000, 100_000, 1_000_000):
    t, _ = bench.reference_step_time(n, 1, steps=1 if n >= 100_000 else 3, warmup=1)
    out["serial"][str(n)] = {"s_per_step": t, "cell_steps_per_s": n / t}
    print("serial", n, t, n / t, flush=True)
for n in (100_000, 1_000_000):
    t, grid = bench.reference_step_time(n, os.cpu_count(), steps=3, warmup=1)
    out["parallel"][str(n)] = {"s_per_step": t, "cell_steps_per_s": n / t, "grid": list(grid), "workers": os.cpu_count()}
    print("parallel", n, t, n / t, grid, flush=True)
json.dump(out, open(ROOT / "gpurun_out" / "cpu_baselines.json", "w"), indent=1)
