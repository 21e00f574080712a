"""Timeline of the first steps of CTA 0 of the pipelined encoder forward (library built with
MATCHA_NVCC_EXTRA=-DMATCHA_ENC_TRACE):  python scripts/dev/enc_trace.py [workload]"""
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
model = build_model(ds, seed=1)
model.train()
x = torch.from_numpy(ds["positives"][:16384]).cuda()
for _ in range(3):
    pred = model(x)
torch.cuda.synchronize()
lib = _lib.load()
buf = (C.c_ulonglong * 4096)()
lib.matcha_enc_trace.argtypes = [C.c_void_p]
assert lib.matcha_enc_trace(buf) == 0
t = np.asarray(buf, dtype=np.int64).reshape(-1, 16)
base = t[0, 0]
print("step: producer [top, f_free seen, issued] | compute [top, f_full seen, tile stored+arrived] | mma [f_full seen, a_full seen, issued]")
for i in range(60):
    if t[i, 0] == 0:
        break
    f = lambda v: int(v - base) if v else -1
    print(i, [f(v) for v in t[i, 0:3]], [f(v) for v in t[i, 4:7]], [f(v) for v in t[i, 8:11]])
