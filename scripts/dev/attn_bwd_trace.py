"""Timeline of the first stages of CTA 0 of the fused attention backward (library built with
MATCHA_NVCC_EXTRA=-DMATCHA_ATTNB_TRACE):  python scripts/dev/attn_bwd_trace.py"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from matcha_b200 import _lib  # noqa: E402
from matcha_b200.synthetic import build_model, make_dataset  # noqa: E402

ds = make_dataset("cfg2", kmers_per_size=50_000, seed=0)
model = build_model(ds, seed=1)
model.train()
x = torch.from_numpy(ds["positives"][:16384]).cuda()
y = torch.ones(len(x), 1, device="cuda")
for _ in range(2):
    model.zero_grad(set_to_none=True)
    torch.nn.functional.binary_cross_entropy_with_logits(model(x), y).backward()
torch.cuda.synchronize()
lib = _lib.load()
buf = (C.c_ulonglong * 4096)()
lib.matcha_attnb_trace.argtypes = [C.c_void_p]
assert lib.matcha_attnb_trace(buf) == 0
t = np.asarray(buf, dtype=np.int64).reshape(-1, 16)
base = t[0, 0]
print("stage n (3 per tile: G, K, Q): compute warp 2 [top, r_full seen, shuffles done (G only), before d_empty, d_empty seen, stored+arrived] | mma [top, d_full seen, issued]")
for i in range(36):
    if t[i, 0] == 0:
        break
    f = lambda v: int(v - base) if v else -1
    print(i, "GKQ"[i % 3], [f(v) for v in t[i, 0:6]], [f(v) for v in t[i, 8:11]])
