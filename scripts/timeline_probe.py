#!/usr/bin/env python
"""Kernel timeline of the steady-state sharded step (CUPTI through torch.profiler; one process per GPU, torchrun):
which kernels run, how long, and how long the stream idles in front of each.
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 scripts/timeline_probe.py [cells_per_gpu] [steps]"""
import os, sys
from collections import defaultdict
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch, torch.distributed as dist
from torch.profiler import profile, ProfilerActivity
import pyafv_b200 as afv
from pyafv_b200.sharded import SlabShardedSimulator
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
cells = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
K = int(sys.argv[2]) if len(sys.argv) > 2 else 10
L = float(np.sqrt(world * cells)); W = L / world
rng = np.random.default_rng(42 + rank)
pts = np.column_stack(((rank + rng.random(cells)) * W * (1 - 1e-12), rng.random(cells) * L))
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
sim = SlabShardedSimulator(pts, 2 * np.pi * rng.random(cells), afv.PhysicalParams(), L, rank=rank, world=world, device=local, stream=stream.cuda_stream)
sim.step(0.01, 1.0, 2.4, 0.3, seed=7, n_steps=30)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    sim.step(0.01, 1.0, 2.4, 0.3, seed=7, n_steps=K)
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
if rank == 0 and evs:
    t_first, t_last = evs[0].time_range.start, max(e.time_range.end for e in evs)
    dur, gap, cnt = defaultdict(float), defaultdict(float), defaultdict(int)
    prev_end = None
    for e in evs:
        name = e.name.split("(")[0].split("<")[0][-40:]
        dur[name] += e.time_range.end - e.time_range.start
        cnt[name] += 1
        if prev_end is not None:
            gap[name] += max(0.0, e.time_range.start - prev_end)
        prev_end = e.time_range.end if prev_end is None else max(prev_end, e.time_range.end)
    print(f"{world} ranks, {cells} cells/GPU, {K} steps: span {(t_last - t_first) / K:.1f} us/step; per step, in launch order:")
    order = []
    for e in evs:
        name = e.name.split("(")[0].split("<")[0][-40:]
        if name not in order:
            order.append(name)
    tot_d = tot_g = 0.0
    for name in order:
        print(f"  {name:42s} x{cnt[name] / K:4.1f}  busy {dur[name] / K:7.1f} us   idle before {gap[name] / K:6.1f} us")
        tot_d += dur[name] / K; tot_g += gap[name] / K
    print(f"  total busy {tot_d:.1f} us, idle {tot_g:.1f} us")
if world > 1:
    dist.destroy_process_group()
