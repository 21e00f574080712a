"""Per-parameter gradient error of the fused attention kernels vs the decomposed pipeline (debug aid):
    python scripts/dev/fused_grad_debug.py [L] [B] [p_drop_on]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import model_from_golden  # noqa: E402
from matcha_b200 import _lib as LIB  # noqa: E402

if __name__ == "__main__":
    L = int(sys.argv[1]) if len(sys.argv) > 1 else 5
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 333
    lib = LIB.load()
    g = np.load(os.path.join(ROOT, "tests", "golden", "ref_small.npz"))
    rng = np.random.default_rng(0)
    N = int(g["chrom_range"][-1][1]) - 1
    x = np.zeros((B, L), dtype=np.int64)
    for b in range(B):
        k = int(rng.integers(2, L + 1))
        x[b, :k] = np.sort(rng.choice(np.arange(1, N + 1), size=k, replace=False))
    x = torch.from_numpy(x).cuda()
    y = torch.from_numpy((rng.random((B, 1)) < 0.3).astype("float32")).cuda()
    w = torch.from_numpy(rng.uniform(0.5, 3, (B, 1)).astype("float32")).cuda()
    np.random.choice = lambda a, size=None: np.asarray([1])
    res = {}
    for fused in (0, 1):
        lib.matcha_set_fused(fused)
        model = model_from_golden(g)
        model.train()
        eng = model._engine()
        eng.seed_base = 91
        pred, rl = model(x, return_recon=True)
        (torch.nn.functional.binary_cross_entropy_with_logits(pred, y, weight=w) + 0.1 * rl.sum()).backward()
        res[fused] = {k: p.grad.detach().cpu().numpy().copy() for k, p in model.named_parameters() if p.grad is not None}
    for k, g0 in res[0].items():
        g1 = res[1][k]
        scale = float(np.abs(g0).max()) + 1e-30
        err = float(np.abs(g1 - g0).max()) / scale
        print(f"{'BAD ' if err > 5e-4 else 'ok  '} {k:60s} rel err {err:.3e}  scale {scale:.3e}")
