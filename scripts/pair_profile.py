import sys, os
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
for _ in range(3):
    _lib.check(lib.matcha_pair_tc_score_range(_lib.ptr(ws), 1, n + 1, 0, 0, total, 1, _lib.ptr(out), _lib.stream_ptr()), "score")
torch.cuda.synchronize()
