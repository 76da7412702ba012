#!/usr/bin/env python
"""Where the time of one exchange goes (one process per GPU, torchrun): CUDA events around pack / transport / finish.
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 scripts/exchange_probe.py [cells_per_gpu]"""
import os, sys, time
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch, torch.distributed as dist
import pyafv_b200 as afv
from pyafv_b200.sharded import SlabShardedSimulator
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
cells = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
L = float(np.sqrt(world * cells)); W = L / world
rng = np.random.default_rng(42 + rank)
pts = np.column_stack(((rank + rng.random(cells)) * W * (1 - 1e-12), rng.random(cells) * L))
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
sim = SlabShardedSimulator(pts, 2 * np.pi * rng.random(cells), afv.PhysicalParams(), L, rank=rank, world=world, device=local, stream=stream.cuda_stream)
sim.step(0.01, 1.0, 2.4, 0.3, seed=7, n_steps=5)
sim._refresh_halo()
ev = lambda: torch.cuda.Event(enable_timing=True)
acc = np.zeros(5); K = 30
for it in range(K):
    e = [ev() for _ in range(6)]
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    e[0].record(stream)
    msgs = sim.exchange_out(); e[1].record(stream)
    t1 = time.perf_counter()
    inbox = sim.comm.exchange_fixed(*msgs); e[2].record(stream)
    t2 = time.perf_counter()
    sim.exchange_in(*inbox); e[3].record(stream)
    t3 = time.perf_counter()
    sim.compute_step(0.01, 1.0, 2.4, 0.3, 7); e[4].record(stream)
    torch.cuda.synchronize()
    acc += [e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2]), e[2].elapsed_time(e[3]), e[3].elapsed_time(e[4]), (t2 - t1) * 1e3]
    sim._sent = None
if rank == 0:
    a = acc / K
    print(f"GPU ms: pack {a[0]:.3f}  transport {a[1]:.3f}  finish(counts+append) {a[2]:.3f}  step {a[3]:.3f}   | host ms in exchange_fixed {a[4]:.3f}")
dist.destroy_process_group()
