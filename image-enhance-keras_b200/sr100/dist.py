"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL on the box, gloo in CPU tests).

Inference shards naturally by tile / image with NO data-path collective: every rank holds
a full weight replica and processes a contiguous range of the column-major tile index (or images
round-robin).  Only the training step has an exchange (gradient all-reduce, sr100.train)."""
from __future__ import annotations

import os


def shard_range(n_items, rank, world):
    """Contiguous [lo, hi) of `n_items` for `rank`; the first n %% world ranks get one extra item."""
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_round_robin(n_items, rank, world):
    return list(range(rank, n_items, world))


def env_rank():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def init_process_group(backend=None):
    """Initialise torch.distributed from RANK / WORLD_SIZE / MASTER_* (torchrun contract); no-op for world 1."""
    import torch
    import torch.distributed as dist
    rank, local_rank, world = env_rank()
    if world == 1 or dist.is_initialized():
        return rank, local_rank, world
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    if backend == "nccl":
        torch.cuda.set_device(local_rank)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, local_rank, world


def max_over_ranks(value):
    """Max of a python float over all ranks (device-side timing is reported as the slowest rank)."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized():
        return float(value)
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor([float(value)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value):
    import torch
    import torch.distributed as dist
    if not dist.is_initialized():
        return float(value)
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor([float(value)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def all_reduce_sum_(tensor):
    """In-place sum over ranks (the training path's one exchange step: the flat fp32 gradient arena, NCCL over
    NVLink on the box; the caller folds 1/world into the optimizer step).  Returns the world size."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return 1
    dist.all_reduce(tensor, op=dist.ReduceOp.SUM)
    return dist.get_world_size()


def gather_objects(obj):
    """All ranks' python objects, in rank order (host-side result assembly; not on the data path)."""
    import torch.distributed as dist
    if not dist.is_initialized():
        return [obj]
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, obj)
    return out


def barrier():
    import torch.distributed as dist
    if dist.is_initialized():
        dist.barrier()
