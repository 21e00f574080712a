"""k-mer enumeration and counting on the GPU: ``generate_kmers.py`` of the reference (build_dict, :8-69, and the per-k
driver, :86-141), which produces the ``all_<k>_counter.npy`` / ``all_<k>_freq_counter.npy`` inputs of training.

Every k-subset of every eligible cluster is enumerated by one thread, filtered by the minimum-distance rule and counted
in a device hash table; k-mers seen at least ``min_freq_cutoff`` times are returned.  The reference emits rows in
process-pool completion order, so the contract is the multiset of (k-mer, frequency): rows come back sorted
lexicographically.  Integer work, bit-exact against ``oracle/kmer_oracle.py`` and the reference's own output.
No CPU fallback: without CUDA this raises.
"""
from __future__ import annotations

from math import comb

import numpy as np
import torch

from . import _lib
from ._lib import MatchaError, check, ptr, stream_ptr

MAX_K = 6
MAX_CLUSTER = 64
MAX_ID = (1 << 21) - 1


def clusters_to_csr(clusters):
    """list of clusters (each: unique ascending node ids) -> (members int64 [nnz], offsets int64 [M + 1]), numpy."""
    sizes = np.fromiter((len(c) for c in clusters), dtype=np.int64, count=len(clusters))
    offsets = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    members = np.concatenate([np.asarray(c, dtype=np.int64) for c in clusters]) if len(clusters) else np.zeros(0, np.int64)
    return members, offsets


def _pow2_at_least(n):
    c = 1
    while c < n:
        c <<= 1
    return c


def count_kmers(members, offsets, k, min_distance=0, max_cluster_size=25, min_freq_cutoff=2, capacity=None, device=None):
    """(rows int64 [n, k], freq int64 [n]) as numpy arrays, rows sorted lexicographically.

    members / offsets: CSR of the clusters (numpy or torch int64).  capacity: slots of the counting table (power of two);
    default 2 x the enumeration work, capped at 2^27 (2 GiB of keys) -- pass a larger one if the call reports a full table."""
    if not torch.cuda.is_available():
        raise MatchaError("count_kmers needs a CUDA device (sm_100a): there is no CPU fallback")
    if not (2 <= k <= MAX_K):
        raise MatchaError(f"k={k} unsupported (2..{MAX_K})")
    if max_cluster_size > MAX_CLUSTER:
        raise MatchaError(f"max_cluster_size={max_cluster_size} above {MAX_CLUSTER}")
    lib = _lib.load()
    dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    members = torch.as_tensor(np.asarray(members) if not torch.is_tensor(members) else members, dtype=torch.int64)
    offsets_h = (offsets.cpu().numpy() if torch.is_tensor(offsets) else np.asarray(offsets)).astype(np.int64)
    M = len(offsets_h) - 1
    empty = (np.zeros((0, k), dtype=np.int64), np.zeros((0,), dtype=np.int64))
    if M <= 0:
        return empty
    sizes = np.diff(offsets_h)
    work = np.zeros(M, dtype=np.int64)
    ok = (sizes >= k) & (sizes <= max_cluster_size)
    lut = np.asarray([comb(n, k) for n in range(MAX_CLUSTER + 1)], dtype=np.int64)
    work[ok] = lut[sizes[ok]]
    prefix = np.concatenate([[0], np.cumsum(work)]).astype(np.int64)
    total = int(prefix[-1])
    if total == 0:
        return empty
    if capacity is None:
        capacity = min(_pow2_at_least(max(1024, 2 * total)), 1 << 27)
    if capacity & (capacity - 1):
        raise MatchaError("capacity must be a power of two")
    members_d = members.to(dev).contiguous()
    offsets_d = torch.from_numpy(offsets_h).to(dev)
    prefix_d = torch.from_numpy(prefix).to(dev)
    table = torch.zeros(2 * capacity, dtype=torch.int64, device=dev)
    counts = torch.zeros(capacity, dtype=torch.int32, device=dev)
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    check(lib.matcha_kmer_count(ptr(members_d), ptr(offsets_d), M, ptr(prefix_d), total, int(k), int(min_distance), ptr(table),
                                capacity, ptr(counts), ptr(status), stream_ptr()), "matcha_kmer_count")
    st = int(status.item())
    if st == 1:
        raise MatchaError(f"k-mer counting table full (capacity {capacity}): pass a larger `capacity`")
    if st == 2:
        raise MatchaError("inconsistent cluster CSR / work prefix (or a cluster above 64 members)")
    if st == 3:
        raise MatchaError(f"node ids must lie in [1, {MAX_ID}]")
    n_keep = int((counts >= max(1, int(min_freq_cutoff))).sum().item())
    if n_keep == 0:
        return empty
    rows = torch.empty(n_keep, k, dtype=torch.int64, device=dev)
    freq = torch.empty(n_keep, dtype=torch.int32, device=dev)
    n_out = torch.zeros(1, dtype=torch.int64, device=dev)
    check(lib.matcha_kmer_collect(ptr(table), capacity, ptr(counts), int(k), max(1, int(min_freq_cutoff)), ptr(rows), ptr(freq),
                                  n_keep, ptr(n_out), stream_ptr()), "matcha_kmer_collect")
    if int(n_out.item()) != n_keep:
        raise MatchaError("k-mer collect: slot count changed between the two passes")
    # canonical order: stable sorts from the last column to the first (plumbing; the counting is the kernel's work)
    order = torch.arange(n_keep, device=dev)
    for j in range(k - 1, -1, -1):
        order = order[torch.sort(rows[order, j], stable=True).indices]
    return rows[order].cpu().numpy(), freq[order].to(torch.int64).cpu().numpy()
