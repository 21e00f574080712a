"""Data-parallel plumbing (one process per GPU, torch.distributed; NCCL on the B200 box, gloo in CPU tests).

The reference is single-process; this is new.  Training replicates the weights,
feature tables and the positive hash set; rank r takes positives r::world of every global batch, draws its
own negatives, and ONE all-reduce per step sums the flat fp32 gradient buffer, whose tail carries the
per-chromosome activity flags so no second collective is needed.  Inference shards by contiguous pair /
batch range with no communication.
"""
from __future__ import annotations

import os

import torch


def init_from_env(backend=None):
    """(rank, world, local_rank); initialises torch.distributed when WORLD_SIZE > 1 (torchrun env)."""
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend)
    return rank, world, local


def shard_rows(n: int, rank: int, world: int):
    """Indices rank::world of a global batch of n rows (interleaved so every rank sees every size class)."""
    return slice(rank, n, world)


def range_shard(total: int, rank: int, world: int):
    """Contiguous [begin, end) share of `total` work items; shares differ by at most one item."""
    return total * rank // world, total * (rank + 1) // world


def allreduce_grads_and_flags(gflat: torch.Tensor, n_flat: int, active: torch.Tensor, world: int):
    """Sum-reduce gradients and OR-reduce activity flags with a single collective.

    gflat  fp32 [n_flat + len(active)]: gradients followed by a scratch tail
    active int32 flags; on return it holds the OR over ranks.  Returns the scale (1 / world) the optimizer
    must apply so the update uses the mean gradient of the global batch."""
    if world <= 1:
        return 1.0
    import torch.distributed as dist
    gflat[n_flat:].copy_(active)
    dist.all_reduce(gflat)
    active.copy_(gflat[n_flat:] > 0)
    return 1.0 / world
