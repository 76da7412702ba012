"""CPU, world_size 2 and 3 over gloo: the halo / migration exchange logic of pyafv_b200.sharded with a numpy
cell store standing in for the GPU handle.  Checks, against brute force on the gathered global state, that after
the halo exchange every rank holds exactly the cells within the halo width of its slab (in its own unwrapped
coordinates), and that migration keeps every cell owned by exactly one rank, the right one."""
from __future__ import annotations

import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pyafv_b200.sharded import ROW, SlabGeometry, TorchDistComm, halo_capacity, halo_ranges, migrant_ranges


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

    # the message format of afv_pack_pair / afv_finish_pair: (cap+1, ROW), row 0 = [count, 0, 0, 0, 0]
    def pack_pair(self, r0, r1, cap, key):
        self._packed = (r0, r1)
        out = []
        for lo, hi, shift in (r0, r1):
            sel = self.owned & (self.rows[:, 0] >= lo) & (self.rows[:, 0] < hi)
            msg = np.zeros((cap + 1, ROW))
            k = int(sel.sum())
            assert k <= cap
            msg[0, 0] = k
            msg[1:k + 1] = self.rows[sel]
            msg[1:k + 1, 0] += shift
            out.append(torch.from_numpy(msg))
        return out

    def finish_pair(self, in0, in1, cap, is_halo, remove_packed):
        if remove_packed:
            (lo0, hi0, _), (lo1, hi1, _) = self._packed
            x = self.rows[:, 0]
            sel = self.owned & (((x >= lo0) & (x < hi0)) | ((x >= lo1) & (x < hi1)))
            self.rows, self.owned = self.rows[~sel], self.owned[~sel]
        for m in (in0, in1):
            k = int(m[0, 0])
            self.append(m[1:k + 1].clone(), is_halo)


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
        for it in range(4):
            # ---- halo ----
            cap = halo_capacity(N, geo)
            store.finish_pair(*comm.exchange_fixed(*store.pack_pair(*halo_ranges(geo), cap, 'halo')), cap, True, False)
            got = {int(g): x for g, x in zip(store.rows[~store.owned, 4], store.rows[~store.owned, 0])}
            # brute force on the global state: minimum-image x distance to the slab
            xs = pts[:, 0]
            want = {}
            for i in range(N):
                if geo.lo <= xs[i] < geo.hi:
                    continue
                for sh in (-L, 0.0, L):
                    x = xs[i] + sh
                    if (geo.lo - halo <= x < geo.lo) or (geo.hi <= x < geo.hi + halo):
                        # a cell can be a halo cell on both sides only when world == 2; keep both images then
                        want.setdefault(i, []).append(x)
            got_multi = {}
            for g, x in zip(store.rows[~store.owned, 4], store.rows[~store.owned, 0]):
                got_multi.setdefault(int(g), []).append(float(x))
            assert set(got_multi) == set(want), (rank, it, len(got_multi), len(want))
            for i in want:
                assert np.allclose(sorted(got_multi[i]), sorted(want[i]), atol=1e-12)
            store.drop_halo()
            # ---- "dynamics": a deterministic drift that pushes cells across faces ----
            drift = 0.9 * np.sin(np.arange(N) + it)
            pts[:, 0] = np.mod(pts[:, 0] + drift, L)              # global truth (every rank tracks it identically)
            ids = store.rows[:, 4].astype(int)
            store.rows[:, 0] += drift[ids]                          # local, un-wrapped like the GPU step leaves it
            # ---- migration ----
            store.finish_pair(*comm.exchange_fixed(*store.pack_pair(*migrant_ranges(geo), cap, 'mig')), cap, False, True)
            ids = store.rows[:, 4].astype(int)
            assert store.owned.all()
            assert np.allclose(store.rows[:, 0], pts[ids, 0], atol=1e-9)
            assert ((store.rows[:, 0] >= geo.lo - 1e-9) & (store.rows[:, 0] < geo.hi + 1e-9)).all()
            counts = torch.tensor([len(ids)])
            dist.all_reduce(counts)
            assert int(counts) == N
            want_mine = np.flatnonzero((pts[:, 0] >= geo.lo) & (pts[:, 0] < geo.hi))
            assert np.array_equal(np.sort(ids), want_mine)
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
