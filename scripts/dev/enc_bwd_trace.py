"""Timeline of the first work items of CTA 0 of the pipelined encoder backward (library built with
MATCHA_NVCC_EXTRA=-DMATCHA_ENC_TRACE):  python scripts/dev/enc_bwd_trace.py [workload]"""
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
y = torch.ones(len(x), 1, device="cuda")
for _ in range(2):
    model.zero_grad(set_to_none=True)
    torch.nn.functional.binary_cross_entropy_with_logits(model(x), y).backward()
torch.cuda.synchronize()
lib = _lib.load()
buf = (C.c_ulonglong * 4096)()
lib.matcha_enc_trace.argtypes = [C.c_void_p]
assert lib.matcha_enc_trace(buf) == 0
t = np.asarray(buf, dtype=np.int64).reshape(-1, 16)
base = t[0, 0]
print("item: converter [top, item_done seen, flushed, dE stored, dH0 seen, dH0pre stored, chunks done] | mma [e_full seen, s_full seen, item committed]")
for i in range(40):
    if t[i, 0] == 0:
        break
    f = lambda v: int(v - base) if v else -1
    print(i, [f(v) for v in t[i, 0:7]], [f(v) for v in t[i, 8:11]])
