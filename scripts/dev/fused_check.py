"""Stand-alone timing / agreement check of the fused kernels against the decomposed pipeline (own process):
    python scripts/dev/fused_check.py [B] [L]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import model_from_golden  # noqa: E402
from matcha_b200 import _lib as L  # noqa: E402


def timed(fn, n=10):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


if __name__ == "__main__":
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
    Lw = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    lib = L.load()
    g = np.load(os.path.join(ROOT, "tests", "golden", "ref_small.npz"))
    model = model_from_golden(g)
    model.eval()
    rng = np.random.default_rng(0)
    N = int(g["chrom_range"][-1][1]) - 1
    x = np.zeros((B, Lw), dtype=np.int64)
    for b in range(B):
        k = int(rng.integers(2, Lw + 1))
        x[b, :k] = np.sort(rng.choice(np.arange(1, N + 1), size=k, replace=False))
    x = torch.from_numpy(x).cuda()
    out = {}
    for fused in (0, 1):
        lib.matcha_set_fused(fused)
        with torch.no_grad():
            out[fused] = model(x).cpu().numpy()
            ms = timed(lambda: model(x))
        print(f"fused={fused}: eval forward B={B} L={Lw}: {ms:.3f} ms  ({B / ms * 1e3:.3e} hyperedges/s)", flush=True)
    err = float(np.abs(out[1] - out[0]).max())
    print(f"max |dlogit| fused vs decomposed: {err:.3e}")
    print("FUSED_OK" if err < 1e-4 else "FUSED_FAIL")
    sys.exit(0 if err < 1e-4 else 1)
