"""Time the tensor-core all-pairs scorer on cfg4's shape (24,897 bins): python scripts/pair_time.py [iters]"""
import os, sys
sys.path.insert(0, os.getcwd())
import torch
from matcha_b200 import _lib
lib = _lib.load()
n, d = 24897, 64
g = torch.Generator(device="cuda").manual_seed(5)
D = torch.randn(n + 1, d, device="cuda", generator=g); S = torch.randn(n + 1, d, device="cuda", generator=g)
cw = torch.rand(d, device="cuda", generator=g); cb = torch.zeros(1, device="cuda")
total = int(lib.matcha_pair_count(1, n + 1, 0))
out = torch.empty(total, dtype=torch.float32, device="cuda")
nbytes = int(lib.matcha_pair_tc_workspace_bytes(1, n + 1))
ws = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
_lib.check(lib.matcha_pair_tc_prepare(_lib.ptr(D), _lib.ptr(S), _lib.ptr(cw), _lib.ptr(cb), d, 1, n + 1, _lib.ptr(ws), nbytes, _lib.stream_ptr()), "prep")
def run(b, e):
    _lib.check(lib.matcha_pair_tc_score_range(_lib.ptr(ws), 1, n + 1, 0, b, e, 1, _lib.ptr(out), _lib.stream_ptr()), "score")
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 10
for name, (b, e) in {"all": (0, total), "first_half": (0, total // 2), "second_half": (total // 2, total)}.items():
    for _ in range(3): run(b, e)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): run(b, e)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print(f"{name}: {ms:.4f} ms  {(e - b) / ms / 1e6:.1f} Gpairs/s  {(e - b) * 4 / ms / 1e6:.0f} GB/s", flush=True)
