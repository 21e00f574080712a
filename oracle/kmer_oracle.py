"""CPU ORACLE (TEST INFRASTRUCTURE) for k-mer enumeration and counting -- SURVEY.md section 8f, rank 1: the producer of
the hot path's inputs (``all_<k>_counter.npy`` / ``all_<k>_freq_counter.npy``).  No product code imports this file.

Restates ``build_dict`` of the reference's ``Code/generate_kmers.py:8-69`` and the per-k driver at ``:86-141``:

  * only clusters with ``k <= len(cluster) <= max_cluster_size`` take part (``:88-91``);
  * for every node i of a cluster (clusters hold unique, ascending node ids -- ``process.py:72-78``), every (k-1)-subset
    of the members ``> i + min_distance`` forms the k-mer ``(i, subset...)`` (``:16-17``), i.e. every k-subset of the
    cluster is produced exactly once, anchored on its smallest member, with first gap ``> min_distance``;
  * for k > 2 the subset is kept only if every gap between consecutive members of the (k-1)-subset is
    ``> min_distance`` (``:23-32``) -- together: ALL k-1 adjacent gaps of the sorted k-mer exceed ``min_distance``;
  * occurrences are counted over all clusters (``:34-36``) and k-mers seen ``>= min_freq_cutoff`` times are kept
    (``:40``).

The reference emits rows in process-pool completion order (``:116-132``), so the contract is the MULTISET of
(k-mer, frequency) pairs; this oracle returns it sorted lexicographically.  Integer work: parity is bit-exact.
Pinned against the unmodified ``build_dict`` executed in the authoring container: ``oracle/make_kmer_golden.py`` ->
``tests/golden/kmer_small.npz`` (checked by ``tests/test_oracle_golden.py``).
"""
from __future__ import annotations

from itertools import combinations
from typing import Sequence, Tuple

import numpy as np


def count_kmers(clusters: Sequence[Sequence[int]], k: int, min_distance: int = 0, max_cluster_size: int = 25,
                min_freq_cutoff: int = 2) -> Tuple[np.ndarray, np.ndarray]:
    """(rows int64 [n, k] sorted lexicographically, freq int64 [n]) -- see the module docstring."""
    counter = {}
    for cl in clusters:
        cl = np.asarray(cl, dtype=np.int64)
        if not (k <= len(cl) <= max_cluster_size):
            continue
        for comb in combinations(cl.tolist(), k):             # ascending input -> ascending tuples
            if all(comb[j + 1] - comb[j] > min_distance for j in range(k - 1)):
                counter[comb] = counter.get(comb, 0) + 1
    kept = sorted((key, v) for key, v in counter.items() if v >= min_freq_cutoff)
    if not kept:
        return np.zeros((0, k), dtype=np.int64), np.zeros((0,), dtype=np.int64)
    return np.asarray([key for key, _ in kept], dtype=np.int64), np.asarray([v for _, v in kept], dtype=np.int64)


def sort_rows(rows: np.ndarray, freq: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """Canonical (lexicographic) order of an unordered (rows, freq) result, e.g. the reference's own output files."""
    rows = np.asarray(rows, dtype=np.int64).reshape(len(freq), -1)
    if len(freq) == 0:
        return rows, np.asarray(freq, dtype=np.int64)
    order = np.lexsort(rows.T[::-1])
    return rows[order], np.asarray(freq, dtype=np.int64)[order]


def n_subsets(cluster_sizes: np.ndarray, k: int, max_cluster_size: int = 25) -> np.ndarray:
    """C(n, k) per cluster (0 for clusters outside [k, max_cluster_size]): the enumeration work a kernel must cover."""
    from math import comb
    n = np.asarray(cluster_sizes, dtype=np.int64)
    out = np.zeros_like(n)
    ok = (n >= k) & (n <= max_cluster_size)
    out[ok] = [comb(int(v), k) for v in n[ok]]
    return out
