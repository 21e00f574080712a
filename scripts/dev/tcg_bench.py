"""Timing of the general tcgen05 contraction kernel (csrc/gemm_tcg.cu) against the SIMT kernel on the embed_dim-128 shapes of
cfg5 (QKG projection, its data and weight gradient; reconstruction head):  python scripts/dev/tcg_bench.py [only_impl]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from matcha_b200 import _lib as L  # noqa: E402

lib = L.load()
T = 81920
shapes = [("qkg_gemm NT", 0, T, 3072, 128), ("qkg_dgrad NN", 1, T, 128, 3072), ("qkg_wgrad TN", 2, 3072, 128, T),
          ("recon_pred NT", 0, 52000, 4980, 128), ("recon_dgrad NN", 1, 52000, 128, 4980), ("recon_wgrad TN", 2, 4980, 128, 52000)]
only = int(sys.argv[1]) if len(sys.argv) > 1 else -1
for name, form, M, N, K in shapes:
    if form == 0:
        A, B = torch.randn(M, K, device="cuda"), torch.randn(N, K, device="cuda")
    elif form == 1:
        A, B = torch.randn(M, K, device="cuda"), torch.randn(K, N, device="cuda")
    else:
        A, B = torch.randn(K, M, device="cuda"), torch.randn(K, N, device="cuda")
    Cm = torch.zeros(M, N, device="cuda")
    for impl in (0, 2):
        if only >= 0 and impl != only:
            continue
        def run():
            L.check(lib.matcha_gemm(form, impl, A.data_ptr(), B.data_ptr(), Cm.data_ptr(), None, M, N, K, A.stride(0), B.stride(0), N,
                                    None, 0, L.stream_ptr()), "matcha_gemm")
        for _ in range(2):
            run()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print(f"{name:16s} impl {impl}: {ms:8.3f} ms  {2.0 * M * N * K / ms / 1e9:8.1f} TFLOP/s  ({(A.numel() + B.numel() + Cm.numel()) * 4 / ms / 1e6:7.0f} GB/s algorithmic)")
