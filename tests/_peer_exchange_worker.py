"""Worker of tests/test_gpu_sharded.py::test_peer_memory_exchange_across_two_processes (one process per GPU, torchrun):
the same 40 steps with the exchange over peer memory and over NCCL, both against one handle holding the whole system."""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import pyafv_b200 as afv  # noqa: E402
from pyafv_b200.sharded import NvlinkPeerComm, SlabShardedSimulator, TorchDistComm  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
N = 60000
L = float(np.sqrt(N))
rng = np.random.default_rng(5)
pts, theta = rng.random((N, 2)) * L, 2 * np.pi * rng.random(N) - np.pi
phys = afv.PhysicalParams()
mine = np.minimum((pts[:, 0] // (L / world)).astype(int), world - 1) == rank
out = {}
for name in ("peer", "nccl"):
    sim = SlabShardedSimulator(pts[mine], theta[mine], phys, L, rank=rank, world=world, device=local, gids=np.flatnonzero(mine),
                               comm=None if name == "peer" else TorchDistComm(rank, world))
    assert isinstance(sim.comm, NvlinkPeerComm if name == "peer" else TorchDistComm), type(sim.comm)
    sim.step(0.01, 1.0, 2.4, 0.3, seed=3, n_steps=20)
    sim._refresh_halo()
    sim.overlap, sim._xs = True, torch.cuda.Stream(sim.device, priority=-1)      # the overlapped form over the same transport
    sim.step(0.01, 1.0, 2.4, 0.3, seed=3, n_steps=20)
    res = sim.build()
    order = np.argsort(res["gids"])                     # (the order of the cells inside a bin differs from run to run; no result does)
    out[name] = {k: v[order] for k, v in res.items()}
    sim.close()
assert np.array_equal(out["peer"]["gids"], out["nccl"]["gids"])
for k in ("pts", "forces", "areas"):
    assert np.array_equal(out["peer"][k], out["nccl"][k]), k
ref = afv.FiniteVoronoiSimulator(pts, phys, periodic=L, device=local)
ref.step(0.01, 1.0, 2.4, 0.3, theta=theta, seed=3, n_steps=40)
g = out["peer"]["gids"]
assert np.array_equal(out["peer"]["pts"], ref.pts[g]) and np.array_equal(out["peer"]["forces"], ref.build(connect=False, rings=False)["forces"][g])
ref.close()
tot = torch.tensor([len(g)], device="cuda")
dist.all_reduce(tot)
assert int(tot.item()) == N
if rank == 0:
    print("PEER-EXCHANGE-OK")
dist.destroy_process_group()
