"""CPU, world_size 2 and 3 over gloo: the halo / migration exchange logic of pyafv_b200.sharded with a numpy
cell store standing in for the GPU handle.  Checks, against brute force on the gathered global state, that after
the halo exchange every rank holds exactly the cells within the halo width of its slab (in the box's own, wrapped
coordinates), and that migration keeps every cell owned by exactly one rank, the right one."""
from __future__ import annotations

import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pyafv_b200.sharded import ROW, SlabGeometry, TorchDistComm, halo_capacity


class NumpyCellStore:
    def __init__(self, rows):
        self.rows = np.asarray(rows, dtype=np.float64).reshape(-1, ROW)
        self.owned = np.ones(len(self.rows), dtype=bool)

    @property
    def n_owned(self):
        return int(self.owned.sum())

    def extract(self, x_lo, x_hi, x_shift, keep, key):
        sel = self.owned & (self.rows[:, 0] >= x_lo) & (self.rows[:, 0] < x_hi)
        out = self.rows[sel].copy()
        out[:, 0] += x_shift
        if not keep:
            self.rows, self.owned = self.rows[~sel], self.owned[~sel]
        return torch.from_numpy(out)

    def append(self, rows, is_halo):
        r = rows.numpy().reshape(-1, ROW)
        self.rows = np.vstack([self.rows, r])
        self.owned = np.concatenate([self.owned, np.full(len(r), not is_halo)])

    def drop_halo(self):
        self.rows, self.owned = self.rows[self.owned], self.owned[self.owned]

    # the message format of afv_pack_exchange / afv_finish_exchange: (cap+1, ROW), header [n_transfer, n_halo, 0, 0, 0],
    # transfers in rows 1 .. n_transfer, halo rows from the back (halo row i at row cap - i)
    def pack_exchange(self, geo, cap):
        self.drop_halo()                              # copies of the previous exchange are retired by the pack pass
        x = geo.offset(self.rows[:, 0])               # minimum-image offset to the lower face (k_exchange_count / _pack)
        self._geo = geo
        out = []
        for mig, halo, shift in (((x < 0), (x >= 0) & (x < geo.halo), geo.shift_to_left),
                                 ((x >= geo.W), (x >= geo.W - geo.halo) & (x < geo.W), geo.shift_to_right)):
            a, b = self.rows[self.owned & mig].copy(), self.rows[self.owned & halo].copy()
            assert len(a) + len(b) <= cap
            msg = np.zeros((cap + 1, ROW))
            msg[0, 0], msg[0, 1] = len(a), len(b)
            msg[1:1 + len(a)] = a                       # transfers from the front ...
            msg[cap - np.arange(len(b))] = b            # ... halo rows from the back (k_exchange_pack)
            msg[1:, 0] += shift
            out.append(torch.from_numpy(msg))
        return out

    def finish_exchange(self, from_left, from_right, cap):
        x = self._geo.offset(self.rows[:, 0])
        self.owned &= (x >= 0) & (x < self._geo.W)   # transferred cells stay as halo copies
        for m in (from_left, from_right):
            na, nb = int(m[0, 0]), int(m[0, 1])
            self.append(m[1:1 + na].clone(), False)
            self.append(m[cap - torch.arange(nb)].clone(), True)


def _worker(rank, world, port, L, N, halo, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(123)          # same stream on every rank: the global state
        pts = rng.random((N, 2)) * L
        geo = SlabGeometry(L, rank, world, 1.0, halo)
        mine = (pts[:, 0] >= geo.lo) & (pts[:, 0] < geo.hi)
        rows = np.zeros((mine.sum(), ROW))
        rows[:, 0:2] = pts[mine]
        rows[:, 4] = np.flatnonzero(mine)
        store, comm = NumpyCellStore(rows), TorchDistComm(rank, world)
        cap = halo_capacity(N, geo)
        for it in range(5):
            # ---- one exchange: transfers (cells that left the slab during the last drift) + halo ----
            store.finish_exchange(*comm.exchange_fixed(*store.pack_exchange(geo, cap)), cap)
            xs = pts[:, 0]                                       # global truth, identical on every rank
            # ownership: exactly the cells whose true position is inside the slab
            ids = store.rows[store.owned, 4].astype(int)
            want_mine = np.flatnonzero((xs >= geo.lo) & (xs < geo.hi))
            assert np.array_equal(np.sort(ids), want_mine), (rank, it)
            assert np.allclose(store.rows[store.owned, 0], xs[ids], atol=1e-9)
            counts = torch.tensor([len(ids)])
            dist.all_reduce(counts)
            assert int(counts) == N
            # halo: every cell of another rank within the halo width of the slab (minimum-image offsets), held once, in the
            # box's own coordinates
            off = geo.offset(xs)
            want = set(np.flatnonzero(((off >= -halo) & (off < 0)) | ((off >= geo.W) & (off < geo.W + halo))).tolist())
            got = {}
            for g, x in zip(store.rows[~store.owned, 4], store.rows[~store.owned, 0]):
                got.setdefault(int(g), []).append(float(x))
            assert want <= set(got), (rank, it, sorted(want - set(got))[:5])
            for i in got:                                       # nothing but the other ranks' cells, once each, unshifted
                assert len(got[i]) == 1 and abs(got[i][0] - xs[i]) < 1e-9 and not (geo.lo <= xs[i] < geo.hi), (rank, it, i, got[i])
            if it % 2:                                          # explicit drop and drop-by-the-next-pack both occur
                store.drop_halo()
            # ---- "dynamics": a deterministic drift that pushes cells across faces (< halo width) ----
            drift = 0.9 * np.sin(np.arange(N) + it)
            ids = store.rows[:, 4].astype(int)
            store.rows[:, 0] = np.mod(store.rows[:, 0] + drift[ids], L)   # wrapped into the box, as the GPU step leaves it
            pts[:, 0] = np.mod(pts[:, 0] + drift, L)
        q.put((rank, "ok"))
    except Exception as e:   # surface the failure to the parent
        q.put((rank, repr(e)))
        raise
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_halo_and_migration_over_gloo(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + world + (os.getpid() % 500)
    procs = [ctx.Process(target=_worker, args=(r, world, port, 40.0, 600, 4.01, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
    res = dict(q.get(timeout=5) for _ in range(world))
    assert all(v == "ok" for v in res.values()), res
