"""Shared test helpers: build our Classifier from the golden fixture and mirror it into the oracle."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def model_from_golden(g, d=64, feats=None, sparse=False):
    """feats / sparse: override the per-chromosome feature tables (e.g. scipy CSR matrices with sparse=True)."""
    from matcha_b200 import hyper_sagnn as M
    cr = g["chrom_range"]
    nums = [int(v) for v in g["nums"]]
    if feats is None:
        feats = [g[f"feat/{c}"] for c in range(len(nums))]
    torch.manual_seed(0)
    ne = M.MultipleEmbedding(feats, d, sparse, np.cumsum(nums), cr, None)
    ne.inter_initial = M.SparseEmbedding(g["inter"], False)       # already z-scored by the reference
    model = M.Classifier(n_head=8, d_model=d, d_k=d, d_v=d, node_embedding=ne, diag_mask=True, bottle_neck=d,
                         attribute_dict=g["p/attribute_dict_embedding.weight"])
    sd = {k[2:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("p/")}
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    meta = json.loads(str(g["meta"]))
    assert set(missing) <= set(meta["no_grad_keys"]), set(missing) - set(meta["no_grad_keys"])
    return model.to(M.device)


def oracle_from_model(model):
    """OracleModel sharing the (CPU copies of the) tensors of one of our Classifier instances."""
    from oracle import hypersagnn_oracle as O
    sd = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    ne = model.node_embedding
    feats = [e.embedding.detach().cpu().clone() for e in ne.embeddings]
    inter = ne.inter_initial.embedding.detach().cpu().clone()
    om = O.OracleModel(sd, feats, inter, np.asarray(ne.chrom_range))
    om.p_feature = float(ne.dropout.p)
    om.p_attn = float(model.encode1.mul_head_attn.dropout.p)
    om.p_pff = float(model.encode1.pff_n1.dropout.p)
    return om


def step_seed(eng):
    """The dropout seed the engine used for its most recent forward (see Engine.next_seed)."""
    return (eng.seed_base * 0x9E3779B97F4A7C15 + eng.tape_id * 0xD1B54A32D192ED03) & ((1 << 64) - 1)


def tc_gemm_rel_err(form, M, N, K, impl, bias=False, seed=0, v2=True):
    """matcha_gemm (form 0: A.B^T, 1: A.B, 2: A^T.B) against fp64 torch: max abs error / max |ref|."""
    from matcha_b200 import _lib as L
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
    Cm = torch.zeros(M, N, device="cuda")
    ns = lib.matcha_gemm_scratch_floats(M) if form == 2 else (N * 64 if (form == 0 and v2) else 0)
    scratch = torch.empty(max(ns, 1), device="cuda")
    L.check(lib.matcha_gemm(form, impl, A.data_ptr(), B.data_ptr(), Cm.data_ptr(), L.ptr(b), M, N, K, A.stride(0),
                            B.stride(0), N, scratch.data_ptr(), ns, L.stream_ptr()), "matcha_gemm")
    torch.cuda.synchronize()
    return (Cm.double() - ref).abs().max().item() / ref.abs().max().item()
