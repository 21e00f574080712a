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
    """utils.py:32-54: 'all <auc> <k> <auc> ...' strings.  The curves are computed on the device
    (matcha_b200/metrics.py: radix sort + fp64 curve areas per size); only the result table is copied back."""
    from matcha_b200.metrics import metric_strings
    # (the reference wraps this in `except BaseException: return 0.0, 0.0` because sklearn raises on one-label slices;
    # the device kernel reports those slices as NaN and they are left out of the strings, so nothing is swallowed here)
    _, roc_str, aupr_str = metric_strings(_dev(y_true), _dev(y_pred), _dev(size_list), _max_size(size_list, max_size))
    return roc_str, aupr_str


def accuracy(output, target, size_list=None, max_size=None):
    """utils.py:57-72."""
    from matcha_b200.metrics import binary_metrics, metric_strings
    if size_list is None:
        return "%.3f " % binary_metrics(_dev(target), _dev(output))["all"][2]
    return metric_strings(_dev(target), _dev(output), _dev(size_list), _max_size(size_list, max_size))[0]


def _dev(t):
    t = t if torch.is_tensor(t) else torch.as_tensor(np.asarray(t))
    return t if t.is_cuda else t.cuda()


def _max_size(size_list, max_size):
    return int(max_size) if max_size else int(torch.as_tensor(size_list).max())
