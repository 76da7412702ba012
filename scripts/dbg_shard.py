import sys, numpy as np
sys.path.insert(0, '/root/repo')
import torch
import pyafv_b200 as afv
from pyafv_b200.sharded import LoopbackShardedSimulator, loopback_exchange
rng = np.random.default_rng(12)
N=6000; L=float(np.sqrt(N/0.4)); world=3
pts = rng.random((N,2))*L; theta = 2*np.pi*rng.random(N)-np.pi
phys = afv.PhysicalParams()
sh = LoopbackShardedSimulator(pts, theta, phys, L, world)
for it in range(60):
    for sim, inbox in zip(sh.ranks, loopback_exchange([s.halo_out() for s in sh.ranks])):
        sim.halo_in(*inbox)
    for sim in sh.ranks:
        sim.compute_step(0.01, 1.0, 2.4, 0.3, 5)
    before = [s.n_owned for s in sh.ranks]
    out = [s.migrants_out() for s in sh.ranks]
    torch.cuda.synchronize()
    hdr = [(int(a[0,0].item()), int(b[0,0].item())) for a, b in out]
    rows = [(a[1:1+int(a[0,0].item())].cpu().numpy().round(4).tolist(), b[1:1+int(b[0,0].item())].cpu().numpy().round(4).tolist()) for a,b in out]
    inb = loopback_exchange(out)
    for sim, inbox in zip(sh.ranks, inb):
        sim.migrants_in(*inbox)
    after = [s.n_owned for s in sh.ranks]
    if sum(after) != N:
        print('step', it, 'before', before, 'hdr(to_left,to_right)', hdr, 'after', after)
        print('rows', rows)
        print('inbox hdr', [(int(a[0,0].item()), int(b[0,0].item())) for a,b in inb])
        break
else:
    print('no loss')
