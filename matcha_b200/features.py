"""Feature construction on the GPU: the inputs of the node encoder / reconstruction head.

    adjacency_from_pixels      process.py:107-170 (parse_cool_contact's pixel loop)  -> intra_adj, inter_adj
    adjacency_from_clusters    process.py:90-105  (edgelist2adj)                      -> edge_list_adj
    corrcoef_features          main.py:572-577    (np.corrcoef per chromosome block, NaN -> 0)
    zscore_positive_rows_      Modules.py:147-152 (row-wise z-score of the positive inter-contact entries, in place)

The reference runs Python loops over every pixel / cluster member pair / matrix row on the host; here each is one kernel
launch (csrc/features.cu).  Arrays keep the reference's dtypes: adjacency float64, features float32.
"""
from __future__ import annotations

import numpy as np
import torch

from ._lib import MatchaError, check, load, ptr, stream_ptr


def _cuda(a, dtype):
    t = a if torch.is_tensor(a) else torch.from_numpy(np.ascontiguousarray(a))
    return t.to(device="cuda", dtype=dtype).contiguous()


def adjacency_from_pixels(bin1, bin2, count, cool2node, node2chrom, n_nodes):
    """bin1, bin2: cooler bin indices of every pixel; count: its (balanced) weight, NaN = skipped; cool2node[i]: 1-based node
    id of cooler bin i or 0 when its chromosome is not in chrom_list; node2chrom[node] (index 0 unused): chromosome index.
    Returns (intra_adj, inter_adj) float64 CUDA tensors [N, N]."""
    lib = load()
    b1, b2 = _cuda(bin1, torch.int64), _cuda(bin2, torch.int64)
    c = _cuda(count, torch.float64)
    c2n, n2c = _cuda(cool2node, torch.int64), _cuda(node2chrom, torch.int32)
    if n2c.numel() != n_nodes + 1:
        raise MatchaError("node2chrom must have n_nodes + 1 entries (node ids are 1-based)")
    intra = torch.zeros(n_nodes, n_nodes, dtype=torch.float64, device="cuda")
    inter = torch.zeros(n_nodes, n_nodes, dtype=torch.float64, device="cuda")
    check(lib.matcha_adj_from_pixels(ptr(b1), ptr(b2), ptr(c), b1.numel(), ptr(c2n), c2n.numel(), ptr(n2c), n_nodes, ptr(intra), ptr(inter),
                                     stream_ptr()), "matcha_adj_from_pixels")
    return intra, inter


def adjacency_from_clusters(members, offsets, n_nodes):
    """CSR clusters (1-based node ids, as edge_list.npy holds them) -> co-occurrence counts float64 [N, N]."""
    lib = load()
    m, o = _cuda(members, torch.int64), _cuda(offsets, torch.int64)
    adj = torch.zeros(n_nodes, n_nodes, dtype=torch.float64, device="cuda")
    check(lib.matcha_adj_from_clusters(ptr(m), ptr(o), o.numel() - 1, n_nodes, ptr(adj), stream_ptr()), "matcha_adj_from_clusters")
    return adj


def corrcoef_features(adj, chrom_range):
    """[np.corrcoef(adj[s-1:e-1, s-1:e-1]) with NaN -> 0 as float32 for every (s, e) of chrom_range] (main.py:572-577).
    adj: fp32 CUDA tensor or numpy array [N, N]; returns a list of fp32 CUDA tensors [n_c, n_c]."""
    lib = load()
    a = adj if (torch.is_tensor(adj) and adj.is_cuda and adj.dtype == torch.float32) else _cuda(adj, torch.float32)
    if a.stride(1) != 1:
        a = a.contiguous()
    out = []
    ws = None
    for (s, e) in np.asarray(chrom_range):
        s, e = int(s), int(e)
        n = e - s
        blk = a[s - 1:e - 1, s - 1:e - 1]
        cc = torch.empty(n, n, dtype=torch.float32, device=a.device)
        if n < 2:
            cc.fill_(0.0)
            out.append(cc)
            continue
        need = int(lib.matcha_corrcoef_workspace_bytes(n))
        if ws is None or ws.numel() < need:
            ws = None
            ws = torch.empty(need, dtype=torch.uint8, device=a.device)
        check(lib.matcha_corrcoef(ptr(blk), blk.stride(0), n, ptr(cc), n, ptr(ws), ws.numel(), stream_ptr()), "matcha_corrcoef")
        out.append(cc)
    return out


def zscore_positive_rows_(inter):
    """In place on an fp32 CUDA matrix (unit column stride): Modules.py:147-152."""
    if not (torch.is_tensor(inter) and inter.is_cuda and inter.dtype == torch.float32 and inter.dim() == 2 and inter.stride(1) == 1):
        raise MatchaError("zscore_positive_rows_ needs an fp32 CUDA matrix with unit column stride")
    check(load().matcha_zscore_positive_rows(ptr(inter), inter.stride(0), inter.shape[0], inter.shape[1], stream_ptr()),
          "matcha_zscore_positive_rows")
    return inter
