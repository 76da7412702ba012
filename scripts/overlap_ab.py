#!/usr/bin/env python
"""Exchange-then-step against the overlapped form on a CALM workload (jammed tissue: relaxed first, then weakly active), the
regime in which the overlapped form engages.  Run under torchrun, one rank per GPU:
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 scripts/overlap_ab.py [cells_per_gpu]"""
import os, sys, json
from pathlib import Path
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import pyafv_b200 as afv
from pyafv_b200.sharded import SlabShardedSimulator

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local),
                        pg_options=dist.ProcessGroupNCCL.Options(is_high_priority_stream=True))
cells = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000
relax = int(sys.argv[2]) if len(sys.argv) > 2 else 400
N = cells * world
L = float(np.sqrt(N)); W = L / world
rng = np.random.default_rng(100 + rank)
pts = np.c_[rank * W + rng.random(cells) * W, rng.random(cells) * L]
stream = torch.cuda.Stream()
with torch.cuda.stream(stream):
    sim = SlabShardedSimulator(pts, 2 * np.pi * rng.random(cells), afv.PhysicalParams(), L, rank=rank, world=world, device=local,
                               stream=stream.cuda_stream, overlap=False)
    sim.step(0.01, 1.0, 0.0, 0.0, seed=1, n_steps=relax)                 # relax: passive, exchange-then-step
    out = {}
    for name, overlap in (("exchange_then_step", False), ("overlapped", True), ("exchange_then_step_again", False)):
        sim.overlap = overlap
        if overlap and sim._xs is None:
            sim._xs = torch.cuda.Stream(priority=-1)
        sim.step(0.01, 1.0, 0.3, 0.3, seed=2, n_steps=20)               # warm-up of this mode (fills the displacement history)
        ov0 = sim.overlapped_steps
        dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        sim.step(0.01, 1.0, 0.3, 0.3, seed=2, n_steps=200)
        e1.record(stream)
        dist.barrier(); torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX)
        out[name] = {"ms_per_step": float(t.item()) / 200, "cell_steps_per_s": N * 200 / (float(t.item()) * 1e-3),
                     "overlapped_steps": sim.overlapped_steps - ov0, "max_dx_recent": max(sim._dx_hist)}
if rank == 0:
    print(json.dumps({"n_gpus": world, "cells_per_gpu": cells, "workload": f"relaxed {relax} steps, then dt=0.01 v0=0.3 Dr=0.3", **out}))
dist.barrier(); dist.destroy_process_group()
