"""k-mer enumeration entry point with the reference's interface (Code/generate_kmers.py): reads ./config.JSON and
`temp_dir/edge_list.npy` (the clusters written by process.py:87: unique ascending node ids per cluster), and writes
`temp_dir/all_<k>_counter.npy` [n, k] and `temp_dir/all_<k>_freq_counter.npy` [n] for every k of "k-mer_size".

The reference enumerates subsets per anchor node in a process pool (generate_kmers.py:103-132, with a 0.2 s sleep per
submitted batch) and concatenates results in completion order; here one kernel launch per k enumerates, filters and
counts every k-subset of every cluster (`matcha_b200.kmers.count_kmers`), and rows are written sorted lexicographically.
The contract -- the multiset of (k-mer, frequency) pairs -- is identical (bit-exact, tests/golden/kmer_small.npz).
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from utils import get_config  # noqa: E402

from matcha_b200.kmers import clusters_to_csr, count_kmers  # noqa: E402


def main():
    config = get_config()
    max_size = config["max_cluster_size"]
    k_list = config["k-mer_size"]
    temp_dir = config["temp_dir"]
    min_dis = config["min_distance"]
    min_freq_cutoff = config["min_freq_cutoff"]
    data = np.load(os.path.join(temp_dir, "edge_list.npy"), allow_pickle=True)
    members, offsets = clusters_to_csr([np.asarray(d, dtype=np.int64) for d in data])
    for k in k_list:
        rows, freq = count_kmers(members, offsets, int(k), int(min_dis), int(max_size), int(min_freq_cutoff))
        if len(rows) > 0:                                           # generate_kmers.py:134-141: nothing is written for an empty set
            print()
            print(rows.shape)
            np.save(os.path.join(temp_dir, "all_%d_counter.npy" % k), rows)
            np.save(os.path.join(temp_dir, "all_%d_freq_counter.npy" % k), freq)
        else:
            print("size %d: no k-mer reaches min_freq_cutoff = %d -- all_%d_counter.npy NOT written (main.py will not find it)"
                  % (k, min_freq_cutoff, k))
        print("Quick summarize")                                    # generate_kmers.py:142-145
        print("total data", len(freq))
        for c in [2, 3, 4, 5, 6, 7, 8]:
            print(">= %d" % c, int(np.sum(freq >= c)))


if __name__ == "__main__":
    main()
