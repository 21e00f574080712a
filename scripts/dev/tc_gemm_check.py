"""Stand-alone check of the tcgen05 contraction kernels against fp64 torch (run in its own process so a
trap cannot poison other tests):  python scripts/dev/tc_gemm_check.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from matcha_b200 import _lib as L  # noqa: E402


def run(form, M, N, K, impl, bias=False, seed=0, v2=True):
    lib = L.load()
    g = torch.Generator(device="cuda").manual_seed(seed)
    if form == 0:
        A, B = torch.randn(M, K, device="cuda", generator=g), torch.randn(N, K, device="cuda", generator=g)
        ref = A.double() @ B.double().t()
    elif form == 1:
        A, B = torch.randn(M, K, device="cuda", generator=g), torch.randn(K, N, device="cuda", generator=g)
        ref = A.double() @ B.double()
    else:
        A, B = torch.randn(K, M, device="cuda", generator=g), torch.randn(K, N, device="cuda", generator=g)
        ref = A.double().t() @ B.double()
    b = torch.randn(N, device="cuda", generator=g) if bias else None
    if b is not None:
        ref = ref + b.double()
    C = torch.zeros(M, N, device="cuda")
    ns = lib.matcha_gemm_scratch_floats(M) if form == 2 else (N * 64 if (form == 0 and v2) else 0)
    scratch = torch.empty(max(ns, 1), device="cuda")
    L.check(lib.matcha_gemm(form, impl, A.data_ptr(), B.data_ptr(), C.data_ptr(), L.ptr(b), M, N, K, A.stride(0),
                            B.stride(0), N, scratch.data_ptr(), ns, L.stream_ptr()), "matcha_gemm")
    torch.cuda.synchronize()
    return (C.double() - ref).abs().max().item() / ref.abs().max().item()


if __name__ == "__main__":
    ok = True
    cases = [(0, 128, 256, 64, True), (0, 1000, 1536, 64, True), (0, 81920, 1536, 64, False),
             (1, 128, 64, 64, False), (1, 333, 64, 1536, False), (1, 81920, 64, 1536, False),
             (2, 512, 64, 4096, False), (2, 1536, 64, 5000, False), (2, 1536, 64, 81920, False)]
    e = run(0, 1000, 1536, 64, 1, True, v2=False)
    print(f"form 0 v1 kernel (no pre-split weights): rel err {e:.2e}", flush=True)
    ok &= e < 3e-5
    for form, M, N, K, bias in cases:
        e1 = run(form, M, N, K, 1, bias)
        e0 = run(form, M, N, K, 0, bias)
        flag = "ok" if e1 < 3e-5 else "FAIL"
        ok &= e1 < 3e-5
        print(f"form {form} M={M} N={N} K={K}: tcgen05 rel err {e1:.2e}   simt {e0:.2e}   {flag}", flush=True)
    print("TC_GEMM_OK" if ok else "TC_GEMM_FAIL")
    sys.exit(0 if ok else 1)
