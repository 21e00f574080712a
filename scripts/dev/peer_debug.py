"""2-rank check of the CUDA-IPC peer mapping used by the fused data-parallel step:
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 scripts/dev/peer_debug.py"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from matcha_b200 import _lib  # noqa: E402
from matcha_b200.parallel import PeerBuffers, init_from_env, peer_access_available  # noqa: E402

rank, world, local = init_from_env()
lib = _lib.load()
print(rank, "peer access available", peer_access_available(world), flush=True)
big = torch.full((1 << 20,), float(rank + 1), device="cuda")
small = torch.full((64,), rank + 10, dtype=torch.int32, device="cuda")
torch.cuda.synchronize()
dist.barrier()
for name, t in (("big", big), ("small", small)):
    pb = PeerBuffers(t, rank, world)
    print(rank, name, "peer ptr", hex(pb.ptrs[1 - rank]), "own", hex(t.data_ptr()), flush=True)
    if name == "big":
        idx = torch.arange(8, device="cuda", dtype=torch.int64)
        out = torch.zeros(8, device="cuda")
        _lib.check(lib.matcha_gather_f32(pb.ptrs[1 - rank], _lib.ptr(idx), 8, _lib.ptr(out), _lib.stream_ptr()), "gather")
        torch.cuda.synchronize()
        print(rank, "kernel on my device read peer memory:", out.tolist(), flush=True)
dist.barrier()
dist.destroy_process_group()
