#!/usr/bin/env python
"""Achieved HBM GB/s of the CSR SpMM encoder kernel (csr_encoder.cu) on synthetic sparse feature rows at the bin counts of
BASELINE.json configs[2] / configs[4] (whole genome at 100 kb / 50 kb):

    python scripts/bench_csr_encoder.py [--res 50000] [--density 0.1] [--tokens 163840]

Algorithmic bytes per launch = 8 B per gathered nonzero (value + column index) + 16 B of row pointers and 8 B of node
id per token + 256 B of output per token; the transposed weight rows (n_c x 256 B per chromosome) are re-read from L2 /
shared memory and not counted.  Prints one JSON line."""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np
import scipy.sparse as sp
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from matcha_b200 import _lib  # noqa: E402
from matcha_b200 import hyper_sagnn as M  # noqa: E402
from matcha_b200.synthetic import WHOLE_GENOME, chrom_bins  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--res", type=int, default=50_000)
    ap.add_argument("--density", type=float, default=0.1)
    ap.add_argument("--tokens", type=int, default=163_840)
    ap.add_argument("--iters", type=int, default=10)
    a = ap.parse_args()
    rng = np.random.default_rng(0)
    nums, chrom_range = chrom_bins(WHOLE_GENOME, a.res)
    N = int(sum(nums))
    feats = []
    for n in nums:                       # per row: ~density * n nonzeros at sorted random columns
        k = max(1, int(a.density * n))
        cols = np.sort(rng.integers(0, n, size=(n, k)), axis=1).astype(np.int32)
        indptr = np.arange(0, (n + 1) * k, k, dtype=np.int64)
        feats.append(sp.csr_matrix((rng.standard_normal(n * k).astype(np.float32), cols.reshape(-1), indptr), shape=(n, n)))
    C_ = len(nums)
    attr = np.zeros((N + 1, C_ + 1), dtype=np.float32)
    torch.manual_seed(1)
    ne = M.MultipleEmbedding(feats, 64, True, np.cumsum(nums), chrom_range, None)
    model = M.Classifier(n_head=8, d_model=64, d_k=64, d_v=64, node_embedding=ne, diag_mask=True, bottle_neck=64,
                         attribute_dict=attr).to(M.device)
    model.eval()
    eng = model._engine()
    ids = torch.from_numpy(rng.integers(1, N + 1, size=a.tokens).astype(np.int64)).cuda()
    lib = _lib.load()
    nnz_row = np.concatenate([np.diff(f.indptr) for f in feats])
    gathered = int(nnz_row[ids.cpu().numpy() - 1].sum())
    with torch.no_grad():
        for _ in range(3):
            eng.node_embeddings(ids)
        torch.cuda.synchronize()
        lib.matcha_profile_enable(1)
        for _ in range(a.iters):
            eng.node_embeddings(ids)
        torch.cuda.synchronize()
        lib.matcha_profile_enable(0)
    n = lib.matcha_profile_labels()
    tms, calls, kern = (C.c_float * n)(), (C.c_int64 * n)(), (C.c_int64 * n)()
    _lib.check(lib.matcha_profile_read(tms, calls, kern, n), "profile_read")
    prof = {lib.matcha_profile_label_name(i).decode(): tms[i] / max(1, calls[i]) for i in range(n) if calls[i]}
    ms = prof["enc0_gather_gemm"]
    nbytes = gathered * 8 + a.tokens * (16 + 8 + 256)
    csr_mb = sum(f.data.nbytes + f.indices.nbytes for f in feats) / 1e6
    print(json.dumps({"kernel": "enc0_csr_fwd", "bins": N, "max_n_c": int(max(nums)), "density": a.density, "tokens": a.tokens,
                      "csr_MB": csr_mb, "gathered_nonzeros": gathered, "ms_per_launch": ms, "algorithmic_bytes": nbytes,
                      "achieved_GBps": nbytes / (ms * 1e-3) / 1e9, "other_ms": {k: round(v, 4) for k, v in prof.items()}}))


if __name__ == "__main__":
    main()
