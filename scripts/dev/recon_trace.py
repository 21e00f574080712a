"""Timeline of the first units of CTA 0 of the pipelined reconstruction head (library built with
MATCHA_NVCC_EXTRA=-DMATCHA_RECON_TRACE):  python scripts/dev/recon_trace.py [workload]"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from matcha_b200 import _lib  # noqa: E402
from matcha_b200.synthetic import build_model, make_dataset  # noqa: E402

workload = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
ds = make_dataset(workload, kmers_per_size=100_000, seed=0)
np.random.choice = lambda a, size=None: np.asarray([0])          # recon on chr1
model = build_model(ds, seed=1)
model.train()
x = torch.from_numpy(ds["positives"][:16384]).cuda()
for _ in range(3):
    pred, rl = model(x, return_recon=True)
torch.cuda.synchronize()
lib = _lib.load()
buf = (C.c_ulonglong * 4096)()
lib.matcha_recon_trace.argtypes = [C.c_void_p]
assert lib.matcha_recon_trace(buf) == 0
t = np.asarray(buf, dtype=np.int64).reshape(-1, 16)
base = t[0, 0]
print("unit: compute warp 0 [top, p_full seen, P done, d_full(prev) seen, arrived, scatter done] | mma [before g_full, g_full seen, issued]")
for i in range(40):
    if t[i, 0] == 0:
        break
    print(i, [int(v - base) for v in t[i, 0:6]], [int(v - base) for v in t[i, 8:11]])
