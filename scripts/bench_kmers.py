#!/usr/bin/env python
"""k-mer enumeration + counting throughput (generate_kmers.py of the reference).

    python scripts/bench_kmers.py [--clusters 2000000] [--k 3] [--out profiles/r01b_kmers.json]

Synthetic SPRITE-like clusters (sizes 2 + Geometric(0.35) capped at max_cluster_size 25, members at |offset|^-1-like
local offsets around an anchor on a 3.1k / 31k / 62k-bin genome).  Times the two kernels with CUDA events (inputs
resident in HBM), the end-to-end `count_kmers` call from host arrays, and the CPU oracle on a bounded sample of the same
clusters.  Algorithmic bytes per enumerated subset: 8 k B of member ids + one 16 B slot compare-and-swap + one 4 B count.
Prints one JSON line.
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def synth_clusters(m, n_nodes, seed=0, max_size=25):
    """(members, offsets) CSR of m clusters, vectorised: candidates in a padded matrix, row-wise sort + dedupe."""
    rng = np.random.default_rng(seed)
    size = np.minimum(2 + rng.geometric(0.35, size=m), max_size).astype(np.int64)
    anchor = rng.integers(1, n_nodes + 1, size=m)
    mag = np.floor(np.exp(rng.random((m, max_size)) * np.log(200.0))).astype(np.int64)
    off = np.where(rng.random((m, max_size)) < 0.5, -mag, mag)
    cand = np.clip(anchor[:, None] + off, 1, n_nodes)
    cand[:, 0] = anchor
    live = np.arange(max_size)[None, :] < size[:, None]
    cand = np.where(live, cand, np.iinfo(np.int64).max)
    cand.sort(axis=1)
    dup = np.zeros_like(live)
    dup[:, 1:] = cand[:, 1:] == cand[:, :-1]
    keep = (cand != np.iinfo(np.int64).max) & ~dup
    counts = keep.sum(1)
    offsets = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    return cand[keep].astype(np.int64), offsets


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--clusters", type=int, default=2_000_000)
    ap.add_argument("--nodes", type=int, default=30344)
    ap.add_argument("--k", type=int, default=3)
    ap.add_argument("--min-distance", type=int, default=0)
    ap.add_argument("--min-freq", type=int, default=2)
    ap.add_argument("--cpu-sample", type=int, default=20000)
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    line = measure(args)
    print(json.dumps(line))
    if args.out:
        open(args.out, "w").write(json.dumps(line) + "\n")


def measure(args):
    """args: namespace with clusters, nodes, k, min_distance, min_freq, cpu_sample.  Returns the result dict (also used by
    bench.py for its `kmers` sub-object)."""
    import torch
    from math import comb
    from matcha_b200 import _lib
    from matcha_b200.kmers import count_kmers
    from oracle import kmer_oracle as KO
    lib = _lib.load()
    members, offsets = synth_clusters(args.clusters, args.nodes)
    sizes = np.diff(offsets)
    k = args.k
    lut = np.asarray([comb(n, k) for n in range(65)], dtype=np.int64)
    work = np.where((sizes >= k) & (sizes <= 25), lut[np.minimum(sizes, 64)], 0)
    prefix = np.concatenate([[0], np.cumsum(work)]).astype(np.int64)
    total = int(prefix[-1])
    cap = 1
    while cap < 2 * total:
        cap <<= 1
    cap = min(cap, 1 << 27)
    dev = torch.device("cuda", torch.cuda.current_device())
    md, od, pd = (torch.from_numpy(a).to(dev) for a in (members, offsets, prefix))
    table = torch.zeros(2 * cap, dtype=torch.int64, device=dev)
    counts = torch.zeros(cap, dtype=torch.int32, device=dev)
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    s = torch.cuda.current_stream().cuda_stream

    def count_once():
        table.zero_(); counts.zero_(); status.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(lib.matcha_kmer_count(md.data_ptr(), od.data_ptr(), len(sizes), pd.data_ptr(), total, k, args.min_distance,
                                         table.data_ptr(), cap, counts.data_ptr(), status.data_ptr(), s), "kmer_count")
        e1.record()
        torch.cuda.synchronize()
        assert int(status.item()) == 0, int(status.item())
        return e0.elapsed_time(e1)
    for _ in range(2):
        count_once()
    ms = float(np.median([count_once() for _ in range(5)]))
    distinct = int((counts > 0).sum().item())
    kept = int((counts >= args.min_freq).sum().item())
    t0 = time.perf_counter()
    rows, freq = count_kmers(members, offsets, k, args.min_distance, 25, args.min_freq)
    e2e_s = time.perf_counter() - t0
    assert len(freq) == kept
    # CPU oracle (the reference's algorithm, one core) on a bounded sample of the same clusters
    ns = min(args.cpu_sample, len(sizes))
    sample = [members[offsets[i]:offsets[i + 1]] for i in range(ns)]
    t0 = time.perf_counter()
    KO.count_kmers(sample, k, args.min_distance, 25, args.min_freq)
    cpu_s = time.perf_counter() - t0
    cpu_subsets = int(work[:ns].sum())
    peak = 6458.7
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    bytes_per_subset = 8.0 * k + 16.0 + 4.0
    line = {"metric": "kmer_subsets_per_s", "value": total / (ms * 1e-3), "unit": "k-subsets/s", "k": k, "clusters": int(len(sizes)),
            "subsets": total, "distinct_kmers": distinct, "kept_kmers": kept, "ms_count_kernel": ms,
            "e2e": {"value": total / e2e_s, "unit": "k-subsets/s", "seconds": e2e_s,
                    "note": "count_kmers() from host arrays: prefix sums, H2D, both kernels, device sort, D2H"},
            "roofline": {"bound": "hbm", "achieved": total * bytes_per_subset / (ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": total * bytes_per_subset / (ms * 1e-3) / 1e9 / peak,
                         "note": "algorithmic bytes per subset: 8k member ids + 16 B slot + 4 B count; random access"},
            "cpu_baseline": {"value": cpu_subsets / cpu_s, "unit": "k-subsets/s", "cores": 1, "kind": "port",
                             "sample": f"first {ns} clusters ({cpu_subsets} subsets), oracle/kmer_oracle.py"}}
    return line


if __name__ == "__main__":
    main()
