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


_OPENED = {}        # (owner rank, handle bytes) -> mapped base address: an allocation can be mapped only once per process


def _allocation_base(ptr: int) -> int:
    """Base address of the cudaMalloc allocation that contains `ptr` (torch's caching allocator sub-allocates)."""
    import ctypes as C
    libcuda = C.CDLL("libcuda.so.1")
    base, size = C.c_uint64(0), C.c_size_t(0)
    rc = libcuda.cuMemGetAddressRange_v2(C.byref(base), C.byref(size), C.c_uint64(ptr))
    if rc != 0:
        raise RuntimeError(f"cuMemGetAddressRange failed ({rc})")
    return int(base.value)


class PeerBuffers:
    """The same-shaped CUDA tensor of every rank, mapped into this process (CUDA IPC, peer access over NVLink).

    `local` is this rank's tensor (storage from cudaMalloc: torch's default caching allocator); `ptrs[r]` is a device
    pointer through which kernels of THIS rank's device read / write rank r's tensor.  Each rank exports the IPC handle of
    the allocation that holds its tensor plus the tensor's offset inside it; handles travel through torch.distributed
    (all_gather_object) and are opened here with this rank's device current (cudaIpcMemLazyEnablePeerAccess) -- torch's
    own tensor sharing opens them under the OWNER's device, which leaves them unreachable from this device's kernels."""

    def __init__(self, local: torch.Tensor, rank: int, world: int):
        import ctypes as C
        import torch.distributed as dist
        from ._lib import check, load
        lib = load()
        base = _allocation_base(local.data_ptr())
        h = (C.c_ubyte * 64)()
        check(lib.matcha_ipc_get_handle(C.c_void_p(base), h), "matcha_ipc_get_handle")
        gathered = [None] * world
        dist.all_gather_object(gathered, (bytes(h), local.data_ptr() - base, local.numel(), str(local.dtype)))
        self.local, self.ptrs = local, []
        for r in range(world):
            handle, offset, numel, dtype = gathered[r]
            if numel != local.numel() or dtype != str(local.dtype):
                raise RuntimeError(f"peer buffer of rank {r} has a different shape")
            if r == rank:
                self.ptrs.append(local.data_ptr())
                continue
            key = (r, handle)
            if key not in _OPENED:
                mapped = C.c_void_p()
                hb = (C.c_ubyte * 64).from_buffer_copy(handle)
                check(lib.matcha_ipc_open(hb, C.byref(mapped)), "matcha_ipc_open")
                _OPENED[key] = int(mapped.value)
            self.ptrs.append(_OPENED[key] + offset)


def peer_access_available(world: int) -> bool:
    """True when every local GPU pair can map each other's memory (NVLink / NVSwitch boxes)."""
    if not torch.cuda.is_available() or torch.cuda.device_count() < world:
        return False
    return all(torch.cuda.can_device_access_peer(a, b) for a in range(world) for b in range(world) if a != b)
