"""Host utilities with the reference's names (Code/utils.py): config loading, metrics, k-mer set building.
`build_hash` returns the device hash set (exact membership) instead of a list of Bloom filters."""
import json
import math  # noqa: F401
import os  # noqa: F401

import numpy as np
import torch


def get_config():                                   # utils.py:157-159: ./config.JSON in the working directory
    with open("./config.JSON", "r") as c:
        return json.load(c)


def build_hash(data, compress=True, min_size=2, max_size=5, capacity=None):
    """utils.py:75-97.  data: zero-padded int array [n, max_size] (or ragged list).  One table for all sizes."""
    from matcha_b200.hyper_sagnn import pad_edges
    from matcha_b200.sampler import KmerHashSet
    rows = pad_edges(data, max_size)
    return KmerHashSet(capacity or len(rows), width=max_size).insert(rows)


def roc_auc_cuda(y_true, y_pred, size_list, max_size):
    """utils.py:32-54: 'all <auc> <k> <auc> ...' strings (sklearn on the host, once per epoch)."""
    from sklearn.metrics import average_precision_score, roc_auc_score
    y_t = (y_true > 0.5).float().cpu().numpy().reshape(-1)
    y_p = y_pred.detach().float().cpu().numpy().reshape(-1)
    s = np.asarray(size_list.cpu() if torch.is_tensor(size_list) else size_list).reshape(-1)
    roc_str, aupr_str = "all %.3f " % roc_auc_score(y_t, y_p), "all %.3f " % average_precision_score(y_t, y_p)
    for k in np.unique(s):
        m = s == k
        if y_t[m].min() == y_t[m].max():
            continue
        roc_str += "%s %.3f " % (str(int(k)), roc_auc_score(y_t[m], y_p[m]))
        aupr_str += "%s %.3f " % (str(int(k)), average_precision_score(y_t[m], y_p[m]))
    return roc_str[:-1], aupr_str[:-1]


def accuracy(output, target, size_list=None, max_size=None):
    """utils.py:57-72."""
    out, tgt = output.detach().cpu().reshape(-1), target.detach().cpu().reshape(-1)
    if size_list is None:
        return "%.3f " % float(((out >= 0.5) == (tgt >= 0.5)).float().mean())
    s = np.asarray(size_list.cpu() if torch.is_tensor(size_list) else size_list).reshape(-1)
    acc_str = ""
    for k in np.unique(s):
        m = torch.from_numpy(s == k)
        acc_str += "%s %.3f " % (str(int(k)), float(((out[m] >= 0.5) == (tgt[m] >= 0.5)).float().mean()))
    return acc_str
