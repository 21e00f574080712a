"""Validation metrics on the device (reference: utils.py:32-72, called once per epoch from main.py:189-195,246-252).

The reference copies every prediction to the host and runs sklearn's roc_auc_score / average_precision_score over all
samples and once per hyperedge size.  Here one call sorts the scores on the GPU (hand-written radix sort,
csrc/metrics.cu) and accumulates both curves per size in fp64; only the (1 + n_sizes) x 4 result table comes back.
"""
from __future__ import annotations

import numpy as np
import torch

from ._lib import MatchaError, check, load, ptr, stream_ptr


def binary_metrics(y_true: torch.Tensor, y_pred: torch.Tensor, size_list: torch.Tensor | None = None, max_size: int = 8):
    """y_true, y_pred: CUDA tensors of n labels (> 0.5 = positive) and n scores; size_list: n hyperedge sizes (ints) or
    None.  Returns {"all": (auroc, aupr, accuracy, count), size: (...), ...} with python floats; rows whose labels are
    all equal carry NaN for the two curve areas (sklearn raises there and the reference swallows it)."""
    if not (y_true.is_cuda and y_pred.is_cuda):
        raise MatchaError("binary_metrics runs on the device: pass CUDA tensors")
    lib = load()
    score = y_pred.detach().reshape(-1).float().contiguous()
    label = y_true.detach().reshape(-1).float().contiguous()
    n = score.numel()
    if n == 0 or label.numel() != n:
        raise MatchaError("binary_metrics: empty input or length mismatch")
    cls, n_classes = None, 0
    if size_list is not None:
        cls = size_list.detach().reshape(-1).to(device=score.device, dtype=torch.int32).contiguous()
        n_classes = int(max_size) + 1
    out = torch.empty((1 + n_classes) * 4, dtype=torch.float64, device=score.device)
    nbytes = int(lib.matcha_metrics_workspace_bytes(n))
    ws = torch.empty(nbytes, dtype=torch.uint8, device=score.device)
    check(lib.matcha_binary_metrics(ptr(score), ptr(label), ptr(cls), n, n_classes, ptr(out), ptr(ws), nbytes, stream_ptr()),
          "matcha_binary_metrics")
    tab = out.cpu().numpy().reshape(1 + n_classes, 4)
    res = {"all": tuple(float(v) for v in tab[0])}
    for c in range(n_classes):
        if tab[1 + c, 3] > 0:
            res[c] = tuple(float(v) for v in tab[1 + c])
    return res


def metric_strings(y_true, y_pred, size_list, max_size=8):
    """The three strings main.py prints / parses: accuracy ("2 0.910 3 0.880 "), AUROC and AUPR ("all 0.912 2 0.901 ...")
    in the formats of utils.py:38-52 and :57-72."""
    m = binary_metrics(y_true, y_pred, size_list, max_size)
    sizes = sorted(k for k in m if k != "all")
    acc = "".join("%s %.3f " % (str(k), m[k][2]) for k in sizes)
    roc, aupr = "all %.3f " % m["all"][0], "all %.3f " % m["all"][1]
    for k in sizes:
        if np.isnan(m[k][0]):
            continue
        roc += "%s %.3f " % (str(k), m[k][0])
        aupr += "%s %.3f " % (str(k), m[k][1])
    return acc, roc[:-1], aupr[:-1]
