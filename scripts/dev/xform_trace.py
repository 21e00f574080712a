"""Timeline of one tile of the X-form attention forward (library built with MATCHA_NVCC_EXTRA=-DMATCHA_XFORM_TRACE)."""
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
model.eval()
x = torch.from_numpy(ds["positives"][:16384]).cuda()
with torch.no_grad():
    for _ in range(3):
        model(x)
torch.cuda.synchronize()
lib = _lib.load()
buf = (C.c_ulonglong * 4096)()
lib.matcha_xform_trace.argtypes = [C.c_void_p]
assert lib.matcha_xform_trace(buf) == 0
t = np.asarray(buf, dtype=np.int64)
base = t[200]
print("compute warp (ns rel.): per head [start, y_full, S done, exchanged, softmax, z_empty, z stored, arrived]")
for h in range(8):
    print(h, [int(v - base) for v in t[200 + h * 8: 208 + h * 8]])
print("u_full wait", int(t[300] - base), int(t[301] - base))
print("mma thread: per head [mma1(h+1) issued, z_full seen, mma2 issued]")
for h in range(8):
    print(h, [int(v - base) for v in t[100 + h * 4: 103 + h * 4]])
