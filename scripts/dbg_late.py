import sys, numpy as np
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import torch
import pyafv_b200 as afv
from pyafv_b200.sharded import LoopbackShardedSimulator, results_dict
from test_gpu_sharded import _setup
pts, theta, A0, L = _setup(12000, 1.0, 31)
phys = afv.PhysicalParams()
res = []
for overlap in (False, True):
    sh = LoopbackShardedSimulator(pts, theta, phys, L, 2, A0=A0, overlap=overlap)
    sh.step(0.01, 1.0, 2.4, 0.3, seed=9, n_steps=1)
    torch.cuda.synchronize()
    parts = [results_dict(s.store.results().clone()) for s in sh.ranks]
    out = {}
    gids = np.concatenate([p["gids"] for p in parts]); order = np.argsort(gids)
    for k in parts[0]:
        out[k] = np.concatenate([p[k] for p in parts])[order]
    res.append(out)
a, b = res
common = np.intersect1d(a["gids"], b["gids"])
ia = np.searchsorted(a["gids"], common); ib = np.searchsorted(b["gids"], common)
a = {k: v[ia] for k, v in a.items()}; b = {k: v[ib] for k, v in b.items()}
pts = pts[common]
print("common", len(common))
for k in ("forces", "areas", "perimeters", "arclens", "coord_nums"):
    d = np.abs(a[k] - b[k]); d = d.reshape(len(d), -1).max(axis=1)
    print(k, "max", d.max(), "n differing", int((d > 0).sum()))
d = np.abs(a["forces"] - b["forces"]).max(axis=1)
bad = np.flatnonzero(d > 0)
W = L / 2
print("x of differing cells (mod W):", np.sort(np.round(pts[bad, 0] % W, 2)))
dA = np.abs(a["areas"] - b["areas"]); print("cells with differing area:", np.flatnonzero(dA > 0), "x mod W", pts[np.flatnonzero(dA > 0), 0] % W)
