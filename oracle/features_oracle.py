"""CPU ORACLE (TEST INFRASTRUCTURE -- never imported by the product path) of the feature-construction steps:
numpy restatements of process.py:90-105 (edgelist2adj) and :148-170 (the pixel loop of parse_cool_contact), pinned against
the UNMODIFIED functions executed by oracle/make_features_golden.py (tests/golden/features_small.npz).  corrcoef / z-score
need no restatement: the reference calls np.corrcoef (main.py:574) and scipy.stats.mstats.zscore (Modules.py:149) directly
and tests call the same library functions."""
import numpy as np


def edgelist2adj(edge_list, n_nodes):
    """process.py:96-103: adj[i-1, j-1] += 1 for every ordered pair i != j of every cluster."""
    adj = np.zeros((n_nodes, n_nodes))
    for e in edge_list:
        e = np.asarray(e, dtype=np.int64)
        ii, jj = np.meshgrid(e, e, indexing="ij")
        keep = ii != jj
        np.add.at(adj, (ii[keep] - 1, jj[keep] - 1), 1.0)
    return adj


def pixels2adj(bin1, bin2, count, cool2node, node2chrom, n_nodes):
    """process.py:148-170.  cool2node: dict cooler index -> 1-based node (bins of other chromosomes are absent, :130-133);
    node2chrom: dict node -> chromosome."""
    intra, inter = np.zeros((n_nodes, n_nodes)), np.zeros((n_nodes, n_nodes))
    for i1, i2, c in zip(bin1, bin2, count):
        if i1 not in cool2node or i2 not in cool2node:
            continue
        a, b = cool2node[i1] - 1, cool2node[i2] - 1
        c = float(c)
        if np.isnan(c):
            continue
        dst = intra if node2chrom[a + 1] == node2chrom[b + 1] else inter
        dst[a, b] += c
        dst[b, a] += c
    return intra, inter
