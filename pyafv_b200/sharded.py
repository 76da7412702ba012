"""Slab sharding of one periodic box across the GPUs of a node: one process (rank) per GPU.

Replaces the reference's multiprocessing domain decomposition (``pyafv/parallel_voronoi.py:125-338`` decompose,
``:351-418`` per-domain build, ``:590-633`` merge): instead of re-decomposing all points on a parent process and
pickling them to workers every step, every rank keeps its x-slab resident on its GPU and exchanges, per step,

1. a halo: the owned cells within ``halo_width`` (default 4.01 l, the reference's default at
   ``parallel_voronoi.py:485`` -- forces need a 4 l dependence range, ``_connectutil.py:19``) of each slab face go
   to the neighbouring rank, in the box's own coordinates (displacements are minimum images on every rank, so the
   operands -- and therefore the results, bit for bit -- are those of the single-GPU run);
2. ownership transfers: cells that crossed a face during the previous update change owner (the old owner keeps a
   halo copy for this step -- the cell is still within the halo of the face it crossed).

Both travel in ONE fixed-capacity message per neighbour and step (header row = the two counts, written on the device),
point-to-point with the two ring neighbours (NCCL send/recv under ``torch.distributed``); there
is no collective on the data path.  The exchange logic is written against two small interfaces so that it can be
exercised without GPUs: a *cell store* (``GpuCellStore`` over the C-ABI handle; the tests use a numpy one) and a
*communicator* (``TorchDistComm``; ``LoopbackComm`` runs all ranks inside one process).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

ROW = 5        # x, y, theta, A0, gid
RES_ROW = 11   # + fx, fy, area, perimeter, arclen, coord_num


# ------------------------------------------------------------------------------------------------------
# cell stores
# ------------------------------------------------------------------------------------------------------
class GpuCellStore:
    """The cells of one rank, resident on one GPU behind an ``AfvHandle``."""

    def __init__(self, handle, device: torch.device):
        self.h = handle
        self.device = device
        self._buf = {}

    def _rows(self, key: str, n: int, width: int = ROW) -> torch.Tensor:
        t = self._buf.get(key)
        if t is None or t.shape[0] < n or t.shape[1] != width:
            t = torch.empty((max(int(n * 1.25) + 64, 256), width), dtype=torch.float64, device=self.device)
            self._buf[key] = t
        return t

    @property
    def n_owned(self) -> int:
        return int(self.h.lib.afv_num_owned(self.h._h))

    @property
    def n(self) -> int:
        return self.h.n

    def set_cells(self, rows: torch.Tensor, n_owned: int) -> None:
        rows = rows.contiguous()
        self.h.call("afv_set_cells_device", rows.data_ptr(), int(n_owned), int(rows.shape[0] - n_owned))

    def extract(self, x_lo: float, x_hi: float, x_shift: float, keep: bool, key: str) -> torch.Tensor:
        cap = max(self.n_owned, 1)
        buf = self._rows(key, cap)
        cnt = C.c_int64()
        self.h.call("afv_extract_slab", float(x_lo), float(x_hi), float(x_shift), 1 if keep else 0, buf.data_ptr(),
                    int(buf.shape[0]), C.byref(cnt))
        return buf[: cnt.value]

    def append(self, rows: torch.Tensor, is_halo: bool) -> None:
        if rows.shape[0]:
            rows = rows.contiguous()
            self.h.call("afv_append_cells", rows.data_ptr(), int(rows.shape[0]), 1 if is_halo else 0)

    def drop_halo(self) -> None:
        self.h.call("afv_drop_halo")

    def pack_exchange(self, geo, cap: int, outbox=None):
        """(to_left, to_right) messages carrying ownership transfers + halo rows; no host synchronisation.  Halo copies
        left from the previous exchange are retired by the same pass.  ``outbox``: where to write them (the neighbours'
        inboxes when the exchange runs over peer memory) instead of this rank's own message buffers."""
        ml, mr = outbox if outbox is not None else (self._msg("xl", cap), self._msg("xr", cap))
        self.h.call("afv_pack_exchange", float(geo.lo), float(geo.hi), float(geo.halo), float(geo.shift_to_left),
                    float(geo.shift_to_right), ml.data_ptr(), mr.data_ptr(), int(cap))
        return ml, mr

    def finish_exchange(self, from_left: torch.Tensor, from_right: torch.Tensor, cap: int) -> None:
        self.h.call("afv_finish_exchange", from_left.data_ptr(), from_right.data_ptr(), int(cap))

    def finish_exchange_step(self, from_left: torch.Tensor, from_right: torch.Tensor, cap: int, dt, mu, v0, Dr, seed, step_index) -> None:
        """finish_exchange + one step; the host's synchronisation for the counts overlaps the cell-list launches."""
        self.h.call("afv_finish_exchange_step", from_left.data_ptr(), from_right.data_ptr(), int(cap), float(dt), float(mu), float(v0),
                    float(Dr), C.c_uint64(int(seed)), C.c_int64(int(step_index)))

    # -- exchange overlapped with the interior geometry: the received rows become "late" cells (afv_late_*) --
    def late_begin(self, cap: int) -> None:
        self.h.call("afv_late_step_begin", int(cap))

    def late_finish(self, from_left: torch.Tensor, from_right: torch.Tensor, cap: int, exchange_stream: int) -> None:
        self.h.call("afv_late_finish", from_left.data_ptr(), from_right.data_ptr(), int(cap), C.c_void_p(int(exchange_stream)))

    def late_end(self, geo, cap: int, exchange_stream: int, dt, mu, v0, Dr, seed, step_index, outbox=None):
        ml, mr = outbox if outbox is not None else (self._msg("xl", cap), self._msg("xr", cap))
        self.h.call("afv_late_step_end", float(dt), float(mu), float(v0), float(Dr), C.c_uint64(int(seed)), C.c_int64(int(step_index)),
                    float(geo.lo), float(geo.hi), float(geo.halo), float(geo.shift_to_left), float(geo.shift_to_right),
                    ml.data_ptr(), mr.data_ptr(), int(cap), C.c_void_p(int(exchange_stream)))
        return ml, mr

    def stream_wait(self, other_stream: int) -> None:
        self.h.call("afv_stream_wait", C.c_void_p(int(other_stream)))

    def _msg(self, key: str, cap: int) -> torch.Tensor:
        t = self._buf.get(key)
        if t is None or t.shape[0] != cap + 1:
            t = torch.zeros((cap + 1, ROW), dtype=torch.float64, device=self.device)
            self._buf[key] = t
        return t

    def results(self) -> torch.Tensor:
        buf = self._rows("res", max(self.n_owned, 1), RES_ROW)
        cnt = C.c_int64()
        self.h.call("afv_pack_owned", buf.data_ptr(), int(buf.shape[0]), C.byref(cnt))
        return buf[: cnt.value]


# ------------------------------------------------------------------------------------------------------
# communicators: exchange one buffer with each ring neighbour
# ------------------------------------------------------------------------------------------------------
class TorchDistComm:
    """Ring neighbours over ``torch.distributed`` (NCCL on GPUs, gloo on CPU)."""

    def __init__(self, rank: int, world: int):
        self.rank, self.world = rank, world
        self.left, self.right = (rank - 1) % world, (rank + 1) % world

    def exchange_fixed(self, to_left: torch.Tensor, to_right: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
        """Fixed-capacity messages (header row carries the count): one batch of two sends and two receives."""
        import torch.distributed as dist
        key = (tuple(to_left.shape), to_left.device)
        if getattr(self, "_rx_key", None) != key:
            self._rx = (torch.zeros_like(to_left), torch.zeros_like(to_right))
            self._rx_key = key
        from_left, from_right = self._rx
        # order matters when both neighbours are the same rank (world == 2): NCCL matches per peer in issue order
        ops = [dist.P2POp(dist.isend, to_left, self.left), dist.P2POp(dist.isend, to_right, self.right),
               dist.P2POp(dist.irecv, from_right, self.right), dist.P2POp(dist.irecv, from_left, self.left)]
        for w in dist.batch_isend_irecv(ops):
            w.wait()
        return from_left, from_right


class NvlinkPeerComm:
    """Ring neighbours over peer memory (NVLink / NVSwitch), one process per GPU on one node.

    Every rank owns two inboxes per side (used alternately) and one flag word per side, and maps those of its two
    neighbours into its address space through CUDA IPC.  ``outbox()`` hands the pack kernel the neighbours' inboxes, so
    the messages are written straight into the other GPU's memory; ``exchange_fixed`` then only publishes the sequence
    number of the exchange in the neighbours' flag words (``afv_flag_signal``, behind the pack on the same stream) and
    makes the stream wait for the neighbours' numbers in its own (``afv_flag_wait``).  No NCCL call, no staging copy and
    no host work beyond two tiny launches per exchange.  ``torch.distributed`` is used once, to pass the IPC handles
    around."""

    def __init__(self, rank: int, world: int, device: torch.device, handle):
        self.rank, self.world = rank, world
        self.left, self.right = (rank - 1) % world, (rank + 1) % world
        self.device, self.h = device, handle
        self._seq = 0
        self._cap = None

    def _setup(self, cap: int) -> None:
        import torch.distributed as dist
        self._msg_bytes = (cap + 1) * ROW * 8
        nbytes = 4 * self._msg_bytes + 256               # [parity][0: from left, 1: from right] messages, then one flag word per side
        mine, ipc = C.c_void_p(), (C.c_ubyte * 64)()
        self.h.call("afv_peer_buffer_create", nbytes, C.byref(mine), ipc)
        everyone = [None] * self.world
        dist.all_gather_object(everyone, bytes(ipc))
        self._base = {self.rank: mine.value}
        for r in {self.left, self.right} - {self.rank}:
            q = C.c_void_p()
            self.h.call("afv_peer_buffer_open", (C.c_ubyte * 64).from_buffer_copy(everyone[r]), C.byref(q))
            self._base[r] = q.value
        self._cap = cap
        dist.barrier()

    def _msg(self, r: int, parity: int, side: int) -> "DevicePointer":
        return DevicePointer(self._base[r] + (2 * parity + side) * self._msg_bytes)

    def _flag(self, r: int, side: int) -> int:
        return self._base[r] + 4 * self._msg_bytes + 4 * side

    def outbox(self, cap: int):
        """(to_left, to_right): the inboxes of the two neighbours that the NEXT exchange fills."""
        if self._cap != cap:
            if self._cap is not None:
                raise RuntimeError("the message capacity of a peer-memory exchange is fixed at first use")
            self._setup(cap)
        par = (self._seq + 1) & 1
        return self._msg(self.left, par, 1), self._msg(self.right, par, 0)    # I am my left neighbour's right neighbour

    def exchange_fixed(self, to_left, to_right):
        """The messages are in place already (``outbox``): publish, then wait for the neighbours' -- all on the current stream."""
        self._seq += 1
        seq, par = self._seq & 0x7FFFFFFF, self._seq & 1
        st = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        self.h.call("afv_flag_signal", st, self._flag(self.left, 1), self._flag(self.right, 0), seq)
        self.h.call("afv_flag_wait", st, self._flag(self.rank, 0), self._flag(self.rank, 1), seq)
        return self._msg(self.rank, par, 0), self._msg(self.rank, par, 1)


class DevicePointer:
    """A raw device address where the exchange code expects a message tensor (it only ever asks for ``data_ptr()``)."""

    def __init__(self, ptr: int):
        self._ptr = int(ptr)

    def data_ptr(self) -> int:
        return self._ptr


def default_comm(rank: int, world: int, device: torch.device, handle):
    """The transport of a one-process-per-GPU run: peer memory when every rank can set it up (``AFV_EXCHANGE=nccl`` forces
    the send / receive pairs of ``TorchDistComm``).  The choice is collective -- one rank without peer access and every
    rank uses NCCL -- and visible as ``type(sim.comm).__name__`` (bench.py reports it)."""
    import os
    import torch.distributed as dist
    want = os.environ.get("AFV_EXCHANGE", "auto").lower()
    if want == "nccl" or not (dist.is_available() and dist.is_initialized()) or dist.get_backend() != "nccl":
        return TorchDistComm(rank, world)
    ok = 1
    try:
        n_dev = torch.cuda.device_count()
        if world > n_dev:
            ok = 0
        else:
            me = device.index
            for r in {(rank - 1) % world, (rank + 1) % world}:
                if r != rank and not torch.cuda.can_device_access_peer(me, r % n_dev):
                    ok = 0
    except Exception:
        ok = 0
    # a trial run of the set-up itself (a 4 KB buffer per rank, mapped by its neighbours): CUDA IPC can be unavailable even where
    # peer access is not (containers without a shared IPC namespace)
    everyone = [None] * world
    mine, ipc = C.c_void_p(), (C.c_ubyte * 64)()
    if ok:
        try:
            handle.call("afv_peer_buffer_create", 4096, C.byref(mine), ipc)
        except Exception:
            ok = 0
    dist.all_gather_object(everyone, bytes(ipc) if ok else None)
    if ok and all(e is not None for e in everyone):
        try:
            for r in {(rank - 1) % world, (rank + 1) % world} - {rank}:
                q = C.c_void_p()
                handle.call("afv_peer_buffer_open", (C.c_ubyte * 64).from_buffer_copy(everyone[r]), C.byref(q))
        except Exception:
            ok = 0
    else:
        ok = 0
    t = torch.tensor([ok], device=device, dtype=torch.int32)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if int(t.item()) == 0:
        if want == "peer":
            raise RuntimeError("AFV_EXCHANGE=peer, but not every rank can map the inboxes of its ring neighbours")
        return TorchDistComm(rank, world)
    return NvlinkPeerComm(rank, world, device, handle)


def loopback_exchange(outboxes: list[tuple[torch.Tensor, torch.Tensor]]) -> list[tuple[torch.Tensor, torch.Tensor]]:
    """All ranks in one process: ``outboxes[r] = (to_left, to_right)`` -> ``[(from_left, from_right)]``."""
    world = len(outboxes)
    return [(outboxes[(r - 1) % world][1], outboxes[(r + 1) % world][0]) for r in range(world)]


# ------------------------------------------------------------------------------------------------------
# the exchange logic (device agnostic)
# ------------------------------------------------------------------------------------------------------
class SlabGeometry:
    """x-slab ``rank`` of ``world`` equal slabs of the periodic square ``[0, L)^2``.

    Every rank keeps the box's own coordinates (wrapped into ``[0, L)``, exactly like the single-GPU run); offsets to the
    slab faces and displacements between cells are minimum images.  Nothing is shifted across the box face, so every rank
    forms a displacement from the same two operands as the single-GPU run: results do not depend on the number of GPUs.
    """

    def __init__(self, L: float, rank: int, world: int, r: float, halo_width: float | None = None):
        self.L, self.rank, self.world, self.r = float(L), int(rank), int(world), float(r)
        self.W = self.L / world
        self.lo, self.hi = rank * self.W, (rank + 1) * self.W
        self.halo = float(4.01 * r if halo_width is None else halo_width)
        if world > 1 and self.W < self.halo:
            raise ValueError(f"slab width {self.W:g} is smaller than the halo width {self.halo:g}: use fewer GPUs")
        pad = 0.5 * r
        if world > 1 and self.W + 2.0 * (self.halo + pad) > self.L:
            # (with two slabs both neighbours are the same rank: its two halo strips must not overlap)
            raise ValueError(f"slab + halo ({self.W + 2 * (self.halo + pad):g}) does not fit the box ({self.L:g}): use fewer GPUs")
        self.shift_to_left = self.shift_to_right = 0.0     # messages carry the box's own coordinates
        self.grid_x0, self.grid_x1 = self.lo - self.halo - pad, self.hi + self.halo + pad

    def offset(self, x):
        """Minimum-image offset of x to the lower face of the slab (what the exchange kernels classify cells by)."""
        d = np.asarray(x, dtype=np.float64) - (self.lo + 0.5 * self.W)     # to the slab centre: left and right stay apart
        return (d - self.L * np.floor(d / self.L + 0.5) if self.world > 1 else d) + 0.5 * self.W


def halo_capacity(n_owned: int, geo: SlabGeometry) -> int:
    """Rows per message: 2.5 x the uniform-density expectation of the halo (+ transfers); checked on arrival."""
    return max(1024, int(2.5 * n_owned * geo.halo / geo.W) + 256)


# ------------------------------------------------------------------------------------------------------
# one rank of the sharded simulator
# ------------------------------------------------------------------------------------------------------
class SlabShardedSimulator:
    """One rank's share of a periodic AFV system sharded into x-slabs (one rank per GPU).

    Args:
        pts: (n,2) positions of the cells this rank owns initially (all inside its slab).
        theta: (n,) self-propulsion angles, or None.
        phys: :class:`~pyafv_b200.PhysicalParams`.
        L: side of the periodic box.
        rank, world: position in the ring of slabs.
        gids: (n,) global cell ids (default: ``rank * 2**40 + arange(n)``; any ids unique across ranks work --
            the Philox noise stream is keyed by them, so results do not depend on the number of GPUs).
        A0: per-cell preferred areas (default ``phys.A0``).
        comm: communicator (default :class:`TorchDistComm`); ``None`` with ``world == 1``.
        overlap: ``False`` (default): exchange, then step.  ``True``: the exchange for a step is packed right after the
            previous update and travels on a second stream while the step already sorts its cells and builds the
            geometry of its interior cells; what arrives joins as *late cells* (a bin table of their own) and the cells
            near the faces are built next to the interior ones.  Exact for any displacement (``afv_late_step_begin`` /
            ``afv_late_finish`` / ``afv_late_step_end``).
        exchange_stream: ``torch.cuda.Stream`` for the overlapped exchange (default: a new high-priority one).  For
            the transport to overlap as well, create the NCCL process group with
            ``pg_options=ProcessGroupNCCL.Options(is_high_priority_stream=True)``.
    """

    def __init__(self, pts, theta, phys, L: float, *, rank: int = 0, world: int = 1, device: int = 0, stream: int | None = None,
                 gids=None, A0=None, halo_width: float | None = None, comm=None, overlap: bool = False,
                 exchange_stream=None):
        from . import _lib
        from .finite_voronoi import _as_params
        self.phys, self.L, self.rank, self.world = phys, float(L), rank, world
        self.geo = SlabGeometry(L, rank, world, phys.r, halo_width)
        self.device = torch.device("cuda", device)
        if stream is None:   # run on torch's current stream: the message buffers and NCCL ops are ordered against it
            stream = torch.cuda.current_stream(self.device).cuda_stream
        self._h = _lib.Handle(_as_params(phys), self.L, device=device, stream=stream)
        if world > 1:
            self._h.call("afv_set_slab", self.geo.grid_x0, self.geo.grid_x1)
            self._h.call("afv_set_deferred_checks", 1)    # one host synchronisation per step (at the exchange), not two
        self.store = GpuCellStore(self._h, self.device)
        self.comm = comm if comm is not None else (default_comm(rank, world, self.device, self._h) if world > 1 else None)
        self._step_index = 0
        self.overlap = bool(overlap) and world > 1
        self.overlapped_steps = 0         # steps taken in the overlapped form so far
        self._xs = None
        self._inbox = None
        self._sent = None                 # plain form: messages for the current positions, packed and on their way already
        self.set_cells(pts, theta, A0, gids)
        if self.overlap:
            # high priority: its blocks (pack, append, the face pass) are dispatched ahead of the pending blocks of the interior pass
            self._xs = exchange_stream if exchange_stream is not None else torch.cuda.Stream(self.device, priority=-1)

    def set_cells(self, pts, theta=None, A0=None, gids=None) -> None:
        """Replace the cells of this rank (all inside its slab): what the constructor does with its arguments."""
        pts = np.asarray(pts, dtype=np.float64).reshape(-1, 2)
        n = pts.shape[0]
        rows = np.empty((n, ROW))
        rows[:, 0:2] = pts
        rows[:, 2] = 0.0 if theta is None else np.asarray(theta, dtype=np.float64)
        rows[:, 3] = self.phys.A0 if A0 is None else np.asarray(A0, dtype=np.float64)
        rows[:, 4] = (self.rank * float(2**40) + np.arange(n)) if gids is None else np.asarray(gids, dtype=np.float64)
        if n and self.world > 1 and ((pts[:, 0] < self.geo.lo).any() or (pts[:, 0] >= self.geo.hi).any()):
            raise ValueError("initial cells must lie inside the rank's slab")
        if self._inbox is not None and self._xs is not None:   # a pending overlapped exchange belongs to the old cells
            self.store.stream_wait(self._xs.cuda_stream)
        self.store.set_cells(torch.from_numpy(rows).to(self.device), n)
        self._cap = None
        self._halo_fresh = False          # the halo copies on the device belong to the current positions
        self._inbox = None                # overlapped form: messages packed for the current positions, on their way, not yet appended
        self._sent = None

    @property
    def n_owned(self) -> int:
        if self._sent is not None:        # transfers of an exchange in flight are booked when it is finished
            self._refresh_halo()
        return self.store.n_owned

    def close(self) -> None:
        if isinstance(self.comm, NvlinkPeerComm) and self.comm._cap is not None:
            # the neighbours write into this rank's inboxes and this rank into theirs: nothing may be released before every
            # rank has seen its last message arrive (collective, like the exchange itself)
            import torch.distributed as dist
            torch.cuda.synchronize(self.device)
            if dist.is_initialized():
                dist.barrier()
        self._h.close()

    def connections(self) -> np.ndarray:
        """(K,2) int64 contact pairs (global ids, i < j) found by this rank in its last ``compute_build(connect=True)``: every
        contact of the system is reported by exactly one rank (the owner of the cell with the smaller id)."""
        K = C.c_int64()
        self._h.call("afv_get_num_connections", C.byref(K))
        conn = np.zeros((K.value, 2), dtype=np.int64)
        self._h.call("afv_get_connections", conn.ctypes.data)
        return conn

    # -- phases, exposed separately so that a loopback driver can interleave the ranks of one process --------
    def _caps(self):
        if self._cap is None:   # fixed at first use: every rank must use the same message size
            n = self.n_owned
            if self.world > 1 and self.comm:
                import torch.distributed as dist
                t = torch.tensor([n], device=self.device, dtype=torch.int64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                n = int(t.item())
            self._cap = (halo_capacity(n, self.geo),)
        return self._cap

    def exchange_out(self):
        """Messages for the (left, right) neighbour: cells that left the slab (ownership transfer) + halo cells."""
        cap = self._caps()[0]
        return self.store.pack_exchange(self.geo, cap, self.comm.outbox(cap) if hasattr(self.comm, "outbox") else None)

    def exchange_in(self, from_left, from_right) -> None:
        self.store.finish_exchange(from_left, from_right, self._caps()[0])
        self._halo_fresh = True

    def _refresh_halo(self) -> None:
        """Halo copies for the current positions, appended the plain way (before a build or an exchange-then-step)."""
        if self.world == 1 or self._halo_fresh:
            return
        if self._inbox is not None:       # left over from the overlapped form: already packed and on its way
            self.store.stream_wait(self._xs.cuda_stream)
            inbox, self._inbox = self._inbox, None
            self.exchange_in(*inbox)
        elif self._sent is not None:      # issued right behind the last update (see step())
            inbox, self._sent = self._sent, None
            self.exchange_in(*inbox)
        else:
            self.exchange_in(*self.comm.exchange_fixed(*self.exchange_out()))

    def exchange_stats(self) -> dict[str, int]:
        out = np.zeros(4, dtype=np.int64)
        self._h.call("afv_get_exchange_stats", out.ctypes.data)
        return dict(late_steps=int(out[0]), full_face_passes=int(out[1]), late_cells=int(out[2]), face_columns=int(out[3]))

    # -- overlapped form, phase by phase (a loopback driver interleaves the ranks of one process) --
    def late_begin(self) -> None:
        self.store.late_begin(self._caps()[0])

    def late_finish(self) -> None:
        inbox, self._inbox = self._inbox, None
        self.store.late_finish(inbox[0], inbox[1], self._caps()[0], self._xs.cuda_stream)

    def late_end(self, dt, mu, v0, Dr, seed):
        cap = self._caps()[0]
        msgs = self.store.late_end(self.geo, cap, self._xs.cuda_stream, dt, mu, v0, Dr, seed, self._step_index,
                                   self.comm.outbox(cap) if hasattr(self.comm, "outbox") else None)
        self._step_index += 1
        self.overlapped_steps += 1
        self._halo_fresh = False
        return msgs

    def compute_step(self, dt, mu, v0, Dr, seed, want_results: bool = False):
        self._h.call("afv_step", float(dt), float(mu), float(v0), float(Dr), C.c_uint64(int(seed)), C.c_int64(self._step_index), 1, None)
        self._step_index += 1
        self._halo_fresh = False
        # rows of the owned cells: updated x, y, theta + F, A, P, L, z.  The halo copies stay until the next exchange
        # retires them (pack_exchange), which costs nothing extra; drop_halo() is the explicit form.
        return self.store.results() if want_results else None

    def compute_build(self, connect: bool = False) -> torch.Tensor:
        from . import _lib
        self._h.call("afv_build", _lib.BUILD_FORCES | (_lib.BUILD_CONNECT if connect else 0))
        return self.store.results().clone()

    # -- the per-rank drivers (one process per GPU) ---------------------------------------------------------
    def step(self, dt: float, mu: float = 1.0, v0: float = 0.0, Dr: float = 0.0, *, seed: int = 0, n_steps: int = 1) -> None:
        """``n_steps`` of: one exchange with the ring neighbours (cells that left the slab during the previous update
        change owner, halo cells are refreshed), then the fused build + force + update.  With ``overlap`` the exchange
        travels while the step sorts its cells and builds its interior geometry."""
        for _ in range(n_steps):
            if not self.overlap:
                if self._sent is not None and self._inbox is None and not self._halo_fresh:
                    # the usual case inside a loop: finish the exchange issued behind the last update and step, in one call
                    inbox, self._sent = self._sent, None
                    self.store.finish_exchange_step(inbox[0], inbox[1], self._caps()[0], dt, mu, v0, Dr, seed, self._step_index)
                    self._step_index += 1
                else:
                    self._refresh_halo()
                    self.compute_step(dt, mu, v0, Dr, seed)
                # The next exchange is packed and handed to NCCL right away, behind the update on the same stream: the host
                # side of that (torch's P2P batch: >= 0.1 ms) runs while the GPU is still busy with this step, instead of
                # leaving it idle at the start of the next one.  The counts are read (one synchronisation) when the next
                # step, or a build, begins.
                if self.world > 1 and self.comm:
                    self._sent = self.comm.exchange_fixed(*self.exchange_out())
                continue
            if self._inbox is None:       # first step of a run: pack and transport on the handle's stream
                self._halo_fresh = False
                self._inbox = self.comm.exchange_fixed(*self.exchange_out())
            self.late_begin()
            self.late_finish()
            msgs = self.late_end(dt, mu, v0, Dr, seed)
            with torch.cuda.stream(self._xs):
                self._inbox = self.comm.exchange_fixed(*msgs)

    def build(self) -> dict[str, np.ndarray]:
        """Forces and geometry of the owned cells (keys of the reference's merged dict, ``parallel_voronoi.py:596-611``,
        plus ``gids`` and ``pts``)."""
        self._refresh_halo()
        return results_dict(self.compute_build())

    def e2e_measure(self, k: int, step: dict) -> dict:
        """End-to-end leg of bench.py for N > 1, the same traffic per cell as the single-GPU leg: every step uploads the
        positions of this rank's cells from pinned host memory (16 B / cell; angle, A0 and id stay resident), exchanges
        transfers and halos, advances one step, and reads back position and id of the cells the rank owns afterwards
        (24 B / cell)."""
        import time
        import torch.distributed as dist
        res = self.build()
        order = np.argsort(res["gids"])
        n = len(order)
        host = torch.from_numpy(res["pts"][order]).contiguous().pin_memory()
        rows = torch.empty((n, ROW), dtype=torch.float64, device=self.device)
        rows[:, 2] = 0.0
        rows[:, 3] = self.phys.A0
        rows[:, 4] = torch.from_numpy(res["gids"][order].astype(np.float64)).to(self.device)
        dev_pos = torch.empty((n, 2), dtype=torch.float64, device=self.device)
        cols = torch.tensor([0, 1, 4], device=self.device)
        cap_out = int(n * 1.25) + 64
        dev_out = torch.empty((cap_out, 3), dtype=torch.float64, device=self.device)
        out = torch.empty((cap_out, 3), dtype=torch.float64).pin_memory()
        h2d = d2h = 0

        def one_step():
            nonlocal h2d, d2h
            dev_pos.copy_(host, non_blocking=True)
            rows[:, 0:2] = dev_pos
            self.store.set_cells(rows, n)
            self._halo_fresh = False
            self._sent = None
            self._refresh_halo()
            r = self.compute_step(step["dt"], step["mu"], step["v0"], step["Dr"], 7, want_results=True)
            m = r.shape[0]
            torch.index_select(r, 1, cols, out=dev_out[:m])
            out[:m].copy_(dev_out[:m], non_blocking=True)
            torch.cuda.synchronize()
            h2d += host.numel() * 8
            d2h += m * 24

        one_step()
        dist.barrier(); torch.cuda.synchronize()
        h2d = d2h = 0
        t0 = time.perf_counter()
        for _ in range(k):
            one_step()
        dist.barrier()
        dt = time.perf_counter() - t0
        tot = torch.tensor([float(n), float(h2d) / k, float(d2h) / k], device=self.device, dtype=torch.float64)
        dist.all_reduce(tot)
        tmax = torch.tensor([dt], device=self.device, dtype=torch.float64)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        return {"value": float(tot[0].item()) * k / float(tmax.item()), "unit": "cell-steps/s", "h2d_bytes_per_step": int(tot[1].item()),
                "d2h_bytes_per_step": int(tot[2].item()), "steps": k,
                "api": "SlabShardedSimulator, per rank: H2D positions (pinned) + set_cells + exchange + step + D2H position and id of the owned cells"}


def results_dict(res: torch.Tensor) -> dict[str, np.ndarray]:
    a = res.cpu().numpy()
    return dict(pts=a[:, 0:2].copy(), theta=a[:, 2].copy(), gids=a[:, 4].astype(np.int64), forces=a[:, 5:7].copy(),
                areas=a[:, 7].copy(), perimeters=a[:, 8].copy(), arclens=a[:, 9].copy(), coord_nums=a[:, 10].astype(np.int64))


# ------------------------------------------------------------------------------------------------------
# all ranks inside one process (tests, and machines with fewer GPUs than slabs)
# ------------------------------------------------------------------------------------------------------
class LoopbackShardedSimulator:
    """``world`` slabs driven from ONE process; same phases as the one-process-per-GPU run.

    ``devices`` (default ``[device]``): CUDA ordinals the slabs are dealt to round-robin.  With one device all slabs share
    it (tests, machines with fewer GPUs than slabs); with several, every slab runs on its own GPU -- the kernels of the
    ranks are issued back to back and run concurrently, the messages of the exchange are copied device to device.  This
    is what :class:`pyafv_b200.ParallelFiniteVoronoiSimulator` drives.  (The late-cell overlap needs all slabs on one
    device here; the one-process-per-GPU form has no such limit.)"""

    def __init__(self, pts, theta, phys, L: float, world: int, *, device: int = 0, A0=None, halo_width=None, overlap: bool = False,
                 devices=None):
        self.devices = [int(d) for d in (devices if devices else [device])]
        self.world, self.L, self.phys = world, float(L), phys
        self.overlap = bool(overlap) and world > 1 and len(set(self.devices)) == 1
        self._inbox = None                # overlapped form: messages packed for the current positions, on their way, not yet appended
        dev0 = torch.device("cuda", self.devices[0])
        xs = torch.cuda.Stream(dev0, priority=-1) if self.overlap else None   # one exchange stream (single device)
        self.ranks = []
        pts = np.asarray(pts, dtype=np.float64)
        parts = self._deal(pts, theta, A0)
        for rk in range(world):
            dv = self.devices[rk % len(self.devices)]
            stream = torch.cuda.current_stream(torch.device("cuda", dv)).cuda_stream   # one stream per device, shared by its ranks
            p, th, a0, g = parts[rk]
            self.ranks.append(SlabShardedSimulator(p, th, phys, L, rank=rk, world=world, device=dv, gids=g, A0=a0, halo_width=halo_width,
                                                   comm=False, stream=stream, overlap=self.overlap, exchange_stream=xs))
        self.N = pts.shape[0]
        self._same_capacity()

    def _same_capacity(self) -> None:
        """Every rank must use the same message size (the halo rows are addressed from the back of the buffer): the
        largest of the per-rank estimates, as the all-reduce of the one-process-per-GPU form gives."""
        cap = max(halo_capacity(s.store.n_owned, s.geo) for s in self.ranks)
        for s in self.ranks:
            s._cap = (cap,)

    def _deal(self, pts, theta, A0):
        """(pts, theta, A0, gids) of every slab: ownership by x (right-sided, the last slab takes the upper face)."""
        N = pts.shape[0]
        theta = np.zeros(N) if theta is None else np.asarray(theta, dtype=np.float64)
        A0 = np.full(N, self.phys.A0) if A0 is None else np.broadcast_to(np.asarray(A0, dtype=np.float64), (N,))
        W = self.L / self.world
        faces = np.arange(1, self.world) * W                                # the very numbers SlabGeometry uses for lo / hi
        owner = np.searchsorted(faces, pts[:, 0], side="right") if N else np.zeros(0, dtype=int)
        return [(pts[owner == rk], theta[owner == rk], A0[owner == rk], np.flatnonzero(owner == rk)) for rk in range(self.world)]

    def set_cells(self, pts, theta=None, A0=None) -> None:
        """New positions (and angles / preferred areas) for the whole system: cells are dealt to the slabs again."""
        pts = np.asarray(pts, dtype=np.float64)
        for sim, (p, th, a0, g) in zip(self.ranks, self._deal(pts, theta, A0)):
            sim.set_cells(p, th, a0, g)
        self.N = pts.shape[0]
        self._inbox = None
        self._same_capacity()

    def _move(self, t: torch.Tensor, sim) -> torch.Tensor:
        return t if t.device == sim.device else t.to(sim.device, non_blocking=True)

    def _refresh_halo(self) -> None:
        if self.world == 1 or all(s._halo_fresh for s in self.ranks):
            return
        if all(s._inbox is not None for s in self.ranks):      # left over from the overlapped form
            for sim in self.ranks:
                sim._refresh_halo()
        elif all(s._sent is not None for s in self.ranks):     # issued behind the last update
            for sim in self.ranks:
                inbox, sim._sent = sim._sent, None
                sim.exchange_in(self._move(inbox[0], sim), self._move(inbox[1], sim))
        else:
            for sim, inbox in zip(self.ranks, loopback_exchange([s.exchange_out() for s in self.ranks])):
                sim.exchange_in(self._move(inbox[0], sim), self._move(inbox[1], sim))

    def step(self, dt, mu=1.0, v0=0.0, Dr=0.0, *, seed=0, n_steps=1) -> None:
        for _ in range(n_steps):
            if not self.overlap:
                if self.world > 1 and all(s._sent is not None and not s._halo_fresh for s in self.ranks):
                    for sim in self.ranks:     # the exchange issued behind the last update: finish it and step, in one call
                        inbox, sim._sent = sim._sent, None
                        sim.store.finish_exchange_step(self._move(inbox[0], sim), self._move(inbox[1], sim), sim._caps()[0], dt, mu, v0, Dr, seed,
                                                       sim._step_index)
                        sim._step_index += 1
                else:
                    self._refresh_halo()
                    for sim in self.ranks:
                        sim.compute_step(dt, mu, v0, Dr, seed)
                if self.world > 1:             # as SlabShardedSimulator.step: the next exchange right behind the update
                    for sim, inbox in zip(self.ranks, loopback_exchange([s.exchange_out() for s in self.ranks])):
                        sim._sent = inbox
                continue
            if any(s._inbox is None for s in self.ranks):
                for sim in self.ranks:
                    sim._halo_fresh = False
                for sim, inbox in zip(self.ranks, loopback_exchange([s.exchange_out() for s in self.ranks])):
                    sim._inbox = inbox
            for sim in self.ranks:
                sim.late_begin()
            for sim in self.ranks:
                sim.late_finish()
            msgs = [sim.late_end(dt, mu, v0, Dr, seed) for sim in self.ranks]
            for sim, inbox in zip(self.ranks, loopback_exchange(msgs)):
                sim._inbox = inbox

    def build(self, connect: bool = False) -> dict[str, np.ndarray]:
        """Merged dict in global cell order (the reference's ``_merge_results``, ``parallel_voronoi.py:590-633``); with
        ``connect`` also ``connections``: the union of the per-rank contact lists, rows sorted."""
        self._refresh_halo()
        res = [sim.compute_build(connect) for sim in self.ranks]             # issued on every device before anything is read back
        parts = [results_dict(r) for r in res]
        out = {}
        gids = np.concatenate([p["gids"] for p in parts])
        order = np.argsort(gids)
        for k in parts[0]:
            out[k] = np.concatenate([p[k] for p in parts])[order]
        if connect:
            conn = np.concatenate([sim.connections() for sim in self.ranks])
            out["connections"] = conn[np.lexsort((conn[:, 1], conn[:, 0]))] if len(conn) else conn.reshape(0, 2)
        return out

    def close(self) -> None:
        for s in self.ranks:
            s.close()
