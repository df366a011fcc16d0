"""Peer-memory plumbing: one process per GPU on one node, device buffers of every rank mapped into every
other rank with CUDA IPC (sr_ipc_*), and the two things that run over them:

* the training exchange (csrc/exchange.cu `sr_exchange_adam_step`): gradient arena, parameter arena and a signal pad of
  every rank -> ONE fused reduce-scatter + Adam + all-gather kernel per rank;
* the tile-sharded image (BASELINE config 5): rank 0's uint8 image mapped into every rank, written by every rank's tail
  convs, closed by a stream-ordered `sr_peer_barrier`.

torch.distributed is only the side channel that carries the 64-byte IPC handles (all_gather_object) and the barrier
that orders set-up; no collective runs on the data path."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib as L


def export_handle(lib, tensor):
    """(handle bytes, byte offset inside the allocation) of a device tensor's storage."""
    h = C.create_string_buffer(64)
    off = C.c_size_t()
    L.check(lib.sr_ipc_export(L.ptr(tensor), h, C.byref(off)))
    return bytes(h.raw), int(off.value)


def open_handle(lib, handle, offset):
    p = C.c_void_p()
    L.check(lib.sr_ipc_open(handle, offset, C.byref(p)))
    return int(p.value)


def shard(lib, n, rank, world):
    lo, hi = C.c_size_t(), C.c_size_t()
    L.check(lib.sr_exchange_shard(n, rank, world, C.byref(lo), C.byref(hi)))
    return int(lo.value), int(hi.value)


class Exchange:
    """The exchange object of one rank over explicit per-rank pointer tables (own entry = local pointer)."""

    def __init__(self, lib, rank, world, n, grad_ptrs, param_ptrs, signal_ptrs, keep=()):
        self.lib, self.rank, self.world, self.n = lib, rank, world, n
        self.keep = list(keep)                  # tensors / mappings that must outlive the object
        arr = lambda ptrs: (C.c_void_p * world)(*[C.c_void_p(p) for p in ptrs])  # noqa: E731
        self.handle = C.c_void_p()
        L.check(lib.sr_exchange_create(rank, world, n, arr(grad_ptrs), arr(param_ptrs), arr(signal_ptrs),
                                       C.byref(self.handle)))
        self.lo, self.hi = shard(lib, n, rank, world)

    def set_timeout_ms(self, ms):
        L.check(self.lib.sr_exchange_set_timeout_ms(self.handle, float(ms)))

    def adam_step(self, m, v, t, lr, beta_1, beta_2, epsilon, grad_scale, max_blocks=0, stream=None):
        L.check(self.lib.sr_exchange_adam_step(self.handle, L.ptr(m), L.ptr(v), int(t), lr, beta_1, beta_2, epsilon,
                                               grad_scale, max_blocks, stream if stream is not None else L.stream_ptr()))

    def timed_out(self):
        flag = C.c_int()
        L.check(self.lib.sr_exchange_status(self.handle, L.stream_ptr(), C.byref(flag)))
        return bool(flag.value)

    def __del__(self):
        h, self.handle = getattr(self, "handle", None), None
        if h:
            self.lib.sr_exchange_destroy(h)


def connect(lib, grads, params, group=None):
    """Build the Exchange of this rank over the ranks' `grads` / `params` tensors (flat fp32, same length everywhere).
    Collective: every rank of the (default) process group calls it.  Raises if the ranks are not on one node or the
    buffers cannot be mapped (the caller then keeps the NCCL all-reduce)."""
    import socket
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    if world > 8:
        raise RuntimeError("the peer-memory exchange covers one node (<= 8 ranks), got world size %d" % world)
    n = grads.numel()
    signal = torch.zeros(lib.sr_exchange_signal_bytes() // 4, dtype=torch.int32, device=grads.device)
    torch.cuda.synchronize(grads.device)
    mine = dict(host=socket.gethostname(), n=n, grads=export_handle(lib, grads), params=export_handle(lib, params),
                signal=export_handle(lib, signal))
    table = [None] * world
    dist.all_gather_object(table, mine, group=group)
    if any(t["host"] != mine["host"] for t in table) or any(t["n"] != n for t in table):
        raise RuntimeError("peer-memory exchange: ranks are on different hosts or hold different arena sizes")
    ptrs = {"grads": [], "params": [], "signal": []}
    own = {"grads": grads.data_ptr(), "params": params.data_ptr(), "signal": signal.data_ptr()}
    err = None
    try:
        for r, t in enumerate(table):
            for key in ptrs:
                ptrs[key].append(own[key] if r == rank else open_handle(lib, *t[key]))
    except L.SrError as e:          # keep the collective structure: every rank reports, every rank decides the same
        err = str(e)
    errs = [None] * world
    dist.all_gather_object(errs, err, group=group)
    if any(errs):
        raise RuntimeError("peer-memory exchange: cannot map a peer buffer (%s)" % next(e for e in errs if e))
    ex = Exchange(lib, rank, world, n, ptrs["grads"], ptrs["params"], ptrs["signal"], keep=(grads, params, signal))
    dist.barrier(group=group)       # every pad is zeroed and mapped before the first step can signal
    return ex


class PeerBarrier:
    """Stream-ordered barrier over the ranks' signal pads (sr_peer_barrier_*)."""

    def __init__(self, lib, rank, world, signal_ptrs, keep=()):
        self.lib, self.rank, self.world = lib, rank, world
        self.keep = list(keep)
        self.handle = C.c_void_p()
        arr = (C.c_void_p * world)(*[C.c_void_p(p) for p in signal_ptrs])
        L.check(lib.sr_peer_barrier_create(rank, world, arr, C.byref(self.handle)))

    def set_timeout_ms(self, ms):
        L.check(self.lib.sr_peer_barrier_set_timeout_ms(self.handle, float(ms)))

    def arrive_wait(self, stream=None):
        L.check(self.lib.sr_peer_barrier_arrive_wait(self.handle, stream if stream is not None else L.stream_ptr()))

    def timed_out(self):
        flag = C.c_int()
        L.check(self.lib.sr_peer_barrier_status(self.handle, L.stream_ptr(), C.byref(flag)))
        return bool(flag.value)

    def __del__(self):
        h, self.handle = getattr(self, "handle", None), None
        if h:
            self.lib.sr_peer_barrier_destroy(h)


def _gather_and_open(lib, mine, keys, rank, world, group):
    """all ranks' handle records -> per key the list of THIS process's pointers (own entry: the local pointer given in
    mine['_own'][key]); collective, raises the same error on every rank."""
    import socket
    import torch.distributed as dist
    own = mine.pop("_own")
    mine["host"] = socket.gethostname()
    table = [None] * world
    dist.all_gather_object(table, mine, group=group)
    err = None
    ptrs = {k: [] for k in keys}
    if any(t["host"] != mine["host"] for t in table):
        err = "ranks are on different hosts"
    else:
        try:
            for r, t in enumerate(table):
                for k in keys:
                    if t.get(k) is None:
                        ptrs[k].append(0)
                    else:
                        ptrs[k].append(own[k] if r == rank else open_handle(lib, *t[k]))
        except L.SrError as e:
            err = str(e)
    errs = [None] * world
    dist.all_gather_object(errs, err, group=group)
    if any(errs):
        raise RuntimeError("peer memory: %s" % next(e for e in errs if e))
    return ptrs


def connect_canvas(lib, device, nbytes, group=None):
    """Rank 0's uint8 canvas of `nbytes` mapped into every rank + a PeerBarrier over fresh signal pads.
    Returns (pointer to the canvas valid in THIS process, canvas tensor on rank 0 / None elsewhere, barrier)."""
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    if world > 8:
        raise RuntimeError("peer memory covers one node (<= 8 ranks), got world size %d" % world)
    signal = torch.zeros(lib.sr_exchange_signal_bytes() // 4, dtype=torch.int32, device=device)
    canvas = torch.zeros(nbytes, dtype=torch.uint8, device=device) if rank == 0 else None
    torch.cuda.synchronize(device)
    mine = dict(signal=export_handle(lib, signal), canvas=export_handle(lib, canvas) if canvas is not None else None,
                _own=dict(signal=signal.data_ptr(), canvas=canvas.data_ptr() if canvas is not None else 0))
    ptrs = _gather_and_open(lib, mine, ("signal", "canvas"), rank, world, group)
    bar = PeerBarrier(lib, rank, world, ptrs["signal"], keep=(signal, canvas))
    dist.barrier(group=group)
    return ptrs["canvas"][0], canvas, bar
