"""Slab decomposition of the SPH step across the GPUs of one box (new in this library; the reference
is single-GPU).

The grid is cut along z into contiguous cell-layer ranges, one per rank (one process per GPU).  The
cell hash is z-major, so every layer is a contiguous run of a rank's sorted arrays and a one-layer
halo reproduces the reference's +-1-cell search exactly.  One step (see include/sph_b200.h,
"slab decomposition"):

    integrate owned
    pack: particles that left [zLo,zHi) + copies of the first/last owned layer       (exchange 1: one fixed-size
          message per neighbour; leavers become owned there, layer copies become ghosts; a rank's own
          leavers are also its ghosts on that side)
    local stable sort by (cell hash, ORIGINAL index)  -- same order as the single-GPU sort
    density of owned           -> (pos,p) and (vel,rho) rows of the boundary layers   (exchange 2: sizes known)
    force of owned

Every exchange is a pair of nearest-neighbour send/recv (torch.distributed batch_isend_irecv: NCCL over
NVLink, ordered on the solver's CUDA stream; gloo for the CPU tests).  No collective sits on the data path and
the host synchronises once per step (the read-back after the sort).

This module is the PROTOCOL MODEL of the decomposition, kept for the tests: the production driver is C++
(pibiti_b200/csrc/sph_multi.cu, `sph_multi_*` in include/sph_b200.h, `lib.MultiSystem`), which runs the same phases
with device-resident bookkeeping, its own NCCL communicator and no host synchronisation; bench.py uses that one.
`slab_step` is written over a list of ranks plus a communicator so that the same code drives
  * one rank per process with `DistComm` (gloo on the CPU with the oracle backend; NCCL also works), and
  * all ranks inside one process with `LocalComm` (tests on a single GPU / on the CPU).
The per-rank work is delegated to a backend object; `GpuSlabBackend` wraps one sph_t handle in slab mode.
"""
from __future__ import annotations

import ctypes as C
import os
import time

import numpy as np

from . import lib as _lib
from .lib import SphError

REC = 12            # floats per particle record: pos xyzw, vel xyzw, (originalIndex, rho, p, 0) as raw words


# -------------------------------------------------------------------------------------------------
# partitioning
def z_cells(pos: np.ndarray, par: np.ndarray) -> np.ndarray:
    """z cell layer of each particle, in the same float32 arithmetic as the kernels."""
    z = pos[:, 2].astype(np.float32)
    wmin = np.float32(par["worldMin"][0][2])
    cs = np.float32(par["cellSize"][0][2])
    return np.floor((z - wmin) / cs).astype(np.int64)


def cut_layers(zc: np.ndarray, grid_z: int, ranks: int, min_layers: int = 2) -> list[int]:
    """Layer boundaries [c0=0, c1, ..., cR=grid_z] giving each rank about the same particle count."""
    hist = np.bincount(np.clip(zc, 0, grid_z - 1), minlength=grid_z)
    cum = np.cumsum(hist)
    total = int(cum[-1])
    cuts = [0]
    for r in range(1, ranks):
        target = total * r / ranks
        c = int(np.searchsorted(cum, target, "left")) + 1
        c = max(c, cuts[-1] + min_layers)
        c = min(c, grid_z - min_layers * (ranks - r))
        cuts.append(c)
    cuts.append(grid_z)
    if any(b - a < min_layers for a, b in zip(cuts, cuts[1:])):
        raise SphError(f"grid has too few z layers ({grid_z}) for {ranks} slabs")
    return cuts


def make_records(pos: np.ndarray, vel: np.ndarray, ids: np.ndarray) -> np.ndarray:
    rec = np.zeros((pos.shape[0], REC), np.float32)
    rec[:, 0:4] = pos
    rec[:, 4:8] = vel
    rec[:, 8] = ids.astype(np.uint32).view(np.float32)
    return rec


def record_ids(rec: np.ndarray) -> np.ndarray:
    return np.ascontiguousarray(rec[:, 8]).view(np.uint32)


# -------------------------------------------------------------------------------------------------
class SlabCaps:
    """Message geometry every rank agrees on: a particle message is (1 + leavers + boundary) records."""

    def __init__(self, leavers: int, boundary: int):
        self.leavers, self.boundary = int(leavers), int(boundary)

    @property
    def rows(self) -> int:
        return 1 + self.leavers + self.boundary

    @staticmethod
    def for_state(par, pos, cuts, safety: float = 4.0, floor: int = 16384) -> "SlabCaps":
        """Sized from the initial state: `safety` x the fullest z layer (every rank computes the same numbers)."""
        zc = z_cells(pos, par)
        gz = int(par["gridSize"][0][2])
        fullest = int(np.bincount(np.clip(zc, 0, gz - 1), minlength=gz).max())
        boundary = int(fullest * safety) + floor
        return SlabCaps(max(boundary // 2, floor), boundary)


class GpuSlabBackend:
    """One GPU's slab: an sph_t handle in slab mode plus its message buffers (torch CUDA tensors)."""

    def __init__(self, params: np.ndarray, capacity: int, z_lo: int, z_hi: int, has_lower: bool, has_upper: bool,
                 device: int, caps: SlabCaps):
        import torch
        self.torch = torch
        self.device = torch.device("cuda", device)
        par = np.ascontiguousarray(params).copy()
        par["numParticles"] = capacity
        self.sys = _lib.SphSystem(par, device)
        self.L, self.h = self.sys.lib, self.sys.h
        self.capacity, self.caps = capacity, caps
        self.has_lower, self.has_upper = has_lower, has_upper
        self._check(self.L.sph_slab_configure(self.h, z_lo, z_hi, int(has_lower), int(has_upper)), "sph_slab_configure")
        with torch.cuda.device(self.device):
            mk = lambda: torch.zeros((caps.rows, REC), dtype=torch.float32, device=self.device)
            self.msg_down, self.msg_up = mk(), mk()
            self.dp_down = torch.zeros((caps.boundary, 8), dtype=torch.float32, device=self.device)
            self.dp_up = torch.zeros((caps.boundary, 8), dtype=torch.float32, device=self.device)
        self.stream = torch.cuda.ExternalStream(self.sys.stream(), device=self.device)
        self.n_owned = 0

    def _check(self, rc, what):
        if rc != 0:
            raise SphError(f"{what} failed ({rc}): {self.L.sph_last_error(self.h).decode()}")

    def _dev(self, x):
        """Any array (numpy / torch on any device) -> contiguous float32 CUDA tensor on this device."""
        t = self.torch
        if isinstance(x, np.ndarray):
            x = t.from_numpy(np.ascontiguousarray(x, np.float32))
        if x.is_cuda and x.device == self.device and x.dtype == t.float32 and x.is_contiguous():
            return x                    # already in place (NCCL path: produced on the solver's stream)
        y = x.to(self.device, dtype=t.float32).contiguous()
        t.cuda.current_stream(self.device).synchronize()    # the copy ran on torch's stream, the solver has its own
        return y

    @staticmethod
    def _p(tensor):
        return C.c_void_p(tensor.data_ptr())

    def set_params(self, params: np.ndarray):
        par = np.ascontiguousarray(params).copy()
        par["numParticles"] = self.capacity
        self.sys.set_params(par)

    def set_owned(self, records):
        r = self._dev(records)
        self.torch.cuda.current_stream(self.device).synchronize()
        self._check(self.L.sph_slab_set_owned(self.h, self._p(r), r.shape[0]), "sph_slab_set_owned")
        self.sys.sync()
        self.n_owned = r.shape[0]

    def get_owned(self):
        out = self.torch.empty((max(self.n_owned, 1), REC), dtype=self.torch.float32, device=self.device)
        n = C.c_int(0)
        self._check(self.L.sph_slab_get_owned(self.h, self._p(out), out.shape[0], C.byref(n)), "sph_slab_get_owned")
        return out[: n.value]

    def integrate(self):
        self._check(self.L.sph_slab_integrate(self.h), "sph_slab_integrate")

    def pack(self):
        self._check(self.L.sph_slab_pack(self.h, self._p(self.msg_down), self._p(self.msg_up), self.caps.leavers, self.caps.boundary),
                    "sph_slab_pack")
        return self.msg_down, self.msg_up

    def integrate_pack(self):
        """integrate() + pack() as one kernel."""
        self._check(self.L.sph_slab_integrate_pack(self.h, self._p(self.msg_down), self._p(self.msg_up), self.caps.leavers,
                                                   self.caps.boundary), "sph_slab_integrate_pack")
        return self.msg_down, self.msg_up

    def unpack(self, below, above):
        b, a = self._dev(below), self._dev(above)
        self._keep = (b, a)             # the launch is asynchronous: keep the sources alive
        self._check(self.L.sph_slab_unpack(self.h, self._p(b), self._p(a), self._p(self.msg_down), self._p(self.msg_up),
                                           self.caps.leavers, self.caps.boundary), "sph_slab_unpack")

    def sort(self):
        c = (C.c_int * 3)()
        self._check(self.L.sph_slab_sort(self.h, c), "sph_slab_sort")
        self.n_owned = c[1]
        self.ghosts = (c[0], c[2])
        return c[0], c[1], c[2]

    def density(self):
        self._check(self.L.sph_slab_density(self.h), "sph_slab_density")

    def pack_dp(self):
        c = (C.c_int * 2)()
        self._check(self.L.sph_slab_pack_dp(self.h, self._p(self.dp_down), self._p(self.dp_up), self.caps.boundary, c), "sph_slab_pack_dp")
        return self.dp_down.view(-1)[: 8 * c[0]].view(-1, 8), self.dp_up.view(-1)[: 8 * c[1]].view(-1, 8)

    def expected_dp(self):
        return self.ghosts

    def unpack_dp(self, below, above):
        b, a = self._dev(below), self._dev(above)
        self._keep_dp = (b, a)
        self._check(self.L.sph_slab_unpack_dp(self.h, self._p(b), b.shape[0], self._p(a), a.shape[0]), "sph_slab_unpack_dp")

    def force(self):
        self._check(self.L.sph_slab_force(self.h), "sph_slab_force")

    def force_interior(self):
        """Particles without ghost neighbours: needs density() only, so it can run while rho,p rows are exchanged."""
        self._check(self.L.sph_slab_force_part(self.h, 1), "sph_slab_force_part(interior)")

    def force_boundary(self):
        self._check(self.L.sph_slab_force_part(self.h, 2), "sph_slab_force_part(boundary)")

    def sync(self):
        self.sys.sync()

    def stats(self) -> dict:
        c = (C.c_int * 6)()
        self._check(self.L.sph_slab_stats(self.h, c), "sph_slab_stats")
        return dict(zip(("max_cell", "work", "ghosts_below", "owned", "ghosts_above", "retired"), list(c)))

    def empty_message(self):
        return self.torch.zeros((self.caps.rows, REC), dtype=self.torch.float32, device=self.device)

    def empty_dp(self):
        return self.torch.zeros((0, 8), dtype=self.torch.float32, device=self.device)


# -------------------------------------------------------------------------------------------------
class LocalComm:
    """All ranks live in this process (tests): a neighbour exchange is a hand-over."""

    def exchange(self, backends, outs, empty, expected=None):
        R = len(outs)
        res = []
        for r in range(R):
            for b in backends:
                b.sync()
            below = _copy(outs[r - 1][1]) if r > 0 else empty(backends[r])
            above = _copy(outs[r + 1][0]) if r < R - 1 else empty(backends[r])
            res.append((below, above))
        return res

    def exchange_begin(self, backends, outs, empty, expected=None):
        return self.exchange(backends, outs, empty, expected)

    def exchange_end(self, pending):
        return pending


class DistComm:
    """One rank per process: nearest-neighbour send/recv over torch.distributed.
    NCCL: tensors stay on the GPU and the operations are ordered on the solver's stream (no host sync).
    gloo: tensors are staged through the host (CPU tests, or several ranks sharing one GPU)."""

    def __init__(self, rank: int, world: int):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank, self.world = rank, world
        self.nccl = dist.get_backend() == "nccl"
        self.bytes_sent = 0

    def exchange_begin(self, backends, outs, empty, expected=None):
        """Start the exchange without making the solver's stream wait for it.  NCCL: the operations run on a second
        stream that first waits for everything enqueued on the solver's stream so far (the packing kernels); whatever
        the caller enqueues on the solver's stream before exchange_end overlaps the transfer.  gloo: synchronous."""
        if not self.nccl:
            return self.exchange(backends, outs, empty, expected)
        t, be = self.torch, backends[0]
        if getattr(be, "comm_stream", None) is None:
            with t.cuda.device(be.device):
                be.comm_stream = t.cuda.Stream(device=be.device)
        ready = t.cuda.Event()
        ready.record(be.stream)
        be.comm_stream.wait_event(ready)
        res = self.exchange(backends, outs, empty, expected, stream=be.comm_stream)
        done = t.cuda.Event()
        done.record(be.comm_stream)
        return (res, done, be)

    def exchange_end(self, pending):
        if not self.nccl:
            return pending
        res, done, be = pending
        be.stream.wait_event(done)
        return res

    def exchange(self, backends, outs, empty, expected=None, stream=None):
        """outs[0] = (to lower, to upper).  Fixed-size messages when `expected` is None (receive buffers have the
        senders' shape); otherwise expected[0] = (rows from below, rows from above)."""
        t, dist = self.torch, self.dist
        be, (down, up) = backends[0], outs[0]
        lower, upper = self.rank - 1, self.rank + 1
        has_lower, has_upper = lower >= 0, upper < self.world
        as_numpy = isinstance(down, np.ndarray)
        if as_numpy:
            down, up = t.from_numpy(np.ascontiguousarray(down)), t.from_numpy(np.ascontiguousarray(up))
        width = down.shape[1]
        n_below, n_above = (down.shape[0], up.shape[0]) if expected is None else expected[0]
        if self.nccl:
            ctx = t.cuda.stream(stream if stream is not None else be.stream)
            dev = down.device
        else:
            ctx = _NullCtx()
            if down.is_cuda:
                be.sync()
            down, up, dev = down.cpu(), up.cpu(), t.device("cpu")
        with ctx:
            below = t.zeros((n_below if has_lower else (n_below if expected is None else 0), width), dtype=t.float32, device=dev)
            above = t.zeros((n_above if has_upper else (n_above if expected is None else 0), width), dtype=t.float32, device=dev)
            ops = []
            if has_lower and down.shape[0]:
                ops.append(dist.P2POp(dist.isend, down.contiguous(), lower))
            if has_lower and below.shape[0]:
                ops.append(dist.P2POp(dist.irecv, below, lower))
            if has_upper and up.shape[0]:
                ops.append(dist.P2POp(dist.isend, up.contiguous(), upper))
            if has_upper and above.shape[0]:
                ops.append(dist.P2POp(dist.irecv, above, upper))
            if ops:
                for w in dist.batch_isend_irecv(ops):
                    w.wait()            # NCCL: orders the stream, does not block the host
        self.bytes_sent += (down.numel() * has_lower + up.numel() * has_upper) * 4
        if as_numpy:
            return [(below.numpy(), above.numpy())]
        return [(below, above)]


def _gather_profile(dist, prof, steps, world, extra=None):
    mine = {k: round(v * 1e3 / steps, 3) for k, v in prof.items()}
    mine.update(extra or {})
    out = [None] * world
    dist.all_gather_object(out, mine)
    return out


def _copy(x):
    return x.copy() if isinstance(x, np.ndarray) else x.clone()


class _NullCtx:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def slab_step(backends, comm, prof: dict | None = None):
    """One SPH step of every rank in `backends` (a single rank under DistComm).
    `prof`, if given, accumulates host wall time per phase (with a stream sync after each: profiling only)."""
    t = [time.perf_counter()]

    def lap(name):
        if prof is not None:
            for b in backends:
                b.sync()
            now = time.perf_counter()
            prof[name] = prof.get(name, 0.0) + now - t[0]
            t[0] = now

    if all(hasattr(b, "integrate_pack") for b in backends) and not os.environ.get("SPH_SLAB_SPLIT_PACK"):
        outs = [b.integrate_pack() for b in backends]
    else:
        for b in backends:
            b.integrate()
        outs = [b.pack() for b in backends]
    lap("integrate+pack")
    inc = comm.exchange(backends, outs, lambda b: b.empty_message())
    lap("exchange particles")
    for b, (below, above) in zip(backends, inc):
        b.unpack(below, above)
        b.sort()
    lap("unpack+sort")
    for b in backends:
        b.density()
    outs = [b.pack_dp() for b in backends]
    lap("density+pack rho,p")
    # the rho,p rows only matter to particles in the first / last owned layer: everything else is evaluated while
    # the rows are in flight
    pending = comm.exchange_begin(backends, outs, lambda b: b.empty_dp(), [b.expected_dp() for b in backends])
    for b in backends:
        b.force_interior()
    inc = comm.exchange_end(pending)
    lap("exchange rho,p + interior force")
    for b, (below, above) in zip(backends, inc):
        b.unpack_dp(below, above)
    for b in backends:
        b.force_boundary()
    lap("boundary force")


def split_initial_state(par, pos, vel, ranks, min_layers=2):
    """Cut the grid by particle count and return (cuts, [records of rank r])."""
    zc = z_cells(pos, par)
    gz = int(par["gridSize"][0][2])
    cuts = cut_layers(zc, gz, ranks, min_layers)
    ids = np.arange(pos.shape[0], dtype=np.uint32)
    parts = []
    for r in range(ranks):
        m = (zc >= cuts[r]) & (zc < cuts[r + 1])
        parts.append(make_records(pos[m], vel[m], ids[m]))
    return cuts, parts


def gather_by_id(record_arrays, n):
    """Concatenate owned records of all ranks and order them by original index."""
    rec = np.concatenate([np.asarray(r) for r in record_arrays], 0)
    ids = record_ids(rec)
    assert rec.shape[0] == n and np.array_equal(np.sort(ids), np.arange(n, dtype=np.uint32)), "particles lost or duplicated"
    out = np.empty_like(rec)
    out[ids] = rec
    return out
