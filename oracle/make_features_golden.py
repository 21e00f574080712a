"""Golden arrays for the feature-construction kernels, produced by EXECUTING the reference's own functions (authoring
container only):

    python oracle/make_features_golden.py        # writes tests/golden/features_small.npz

``Code/process.py`` runs its whole pipeline at import (reads config.JSON, cluster files, an .mcool through h5py -- absent
here), so its functions ``edgelist2adj`` (:90-105) and ``parse_cool_contact`` (:107-172) are lifted VERBATIM with ``ast`` and
executed in a namespace that provides the globals they read (``temp_dir``, ``mcool_path``, ``res``, ``chrom_list``) and a
stand-in ``h5py`` whose ``File`` returns the nested dict of arrays a cooler file holds.  Their np.load / np.save go to a
scratch temp_dir.  Also stores np.corrcoef / scipy zscore results for the same data (the reference's calls, main.py:574,
Modules.py:149)."""
import ast
import os
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/Code/process.py"


def lift(names):
    tree = ast.parse(open(REF).read())
    fns = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in names]
    assert len(fns) == len(names)
    return ast.Module(body=fns, type_ignores=[])


def main():
    rng = np.random.default_rng(0)
    chroms = ["chrA", "chrB", "chrC"]
    nums = [30, 22, 17]
    res = 1000
    starts = np.concatenate([[0], np.cumsum(nums)])
    chrom_range = np.stack([starts[:-1] + 1, starts[1:] + 1], 1).astype(np.int64)
    N = int(sum(nums))
    node2chrom, bin2node = {}, {}
    for c, (s, e) in enumerate(chrom_range):
        for i in range(int(s), int(e)):
            node2chrom[i] = c
            bin2node["%s:%d" % (chroms[c], (i - int(s)) * res)] = i
    # a cooler file with one extra chromosome (chrZ, not in chrom_list) whose bins must be skipped
    cool_chrom_names = ["chrA", "chrZ", "chrB", "chrC"]
    cool_nums = [30, 9, 22, 17]
    cool_chrom = np.concatenate([np.full(n, i) for i, n in enumerate(cool_nums)])
    cool_start = np.concatenate([np.arange(n) * res for n in cool_nums])
    nb = len(cool_chrom)
    n_pix = 2500
    b1 = rng.integers(0, nb, n_pix)
    b2 = rng.integers(0, nb, n_pix)
    b1, b2 = np.minimum(b1, b2), np.maximum(b1, b2)
    key = np.unique(b1 * nb + b2)                       # cooler pixels are unique (bin1 <= bin2)
    b1, b2 = key // nb, key % nb
    cnt = rng.gamma(2.0, 1.5, len(b1))
    cnt[rng.random(len(b1)) < 0.05] = np.nan            # unbalanced bins
    clusters = []
    for _ in range(400):
        m = int(min(12, 2 + rng.geometric(0.4)))
        clusters.append(sorted(set(int(v) for v in rng.integers(1, N + 1, m))))
    clusters = [c for c in clusters if len(c) > 1]

    with tempfile.TemporaryDirectory() as td:
        np.save(os.path.join(td, "chrom_range.npy"), chrom_range)
        np.save(os.path.join(td, "node2chrom.npy"), node2chrom, allow_pickle=True)
        np.save(os.path.join(td, "bin2node.npy"), bin2node, allow_pickle=True)
        arr = np.empty(len(clusters), dtype=object)
        for i, c in enumerate(clusters):
            arr[i] = c
        np.save(os.path.join(td, "edge_list.npy"), arr, allow_pickle=True)
        cool = {"resolutions": {str(res): {"bins": {"chrom": cool_chrom, "start": cool_start},
                                           "chroms": {"name": np.asarray(cool_chrom_names).astype("S")},
                                           "pixels": {"bin1_id": b1, "bin2_id": b2, "balanced": cnt}}}}
        h5py = types.SimpleNamespace(File=lambda path, mode: cool)
        ns = dict(np=np, os=os, sys=sys, h5py=h5py, temp_dir=td, mcool_path="synthetic.mcool", res=res, chrom_list=chroms,
                  tqdm=lambda x, *a, **k: x, trange=lambda n, *a, **k: range(n))
        exec(compile(lift(["edgelist2adj", "parse_cool_contact"]), "process.py(functions)", "exec"), ns)
        ns["edgelist2adj"]()
        ns["parse_cool_contact"]()
        edge_adj = np.load(os.path.join(td, "edge_list_adj.npy"))
        intra = np.load(os.path.join(td, "intra_adj.npy"))
        inter = np.load(os.path.join(td, "inter_adj.npy"))
    members = np.concatenate([np.asarray(c, dtype=np.int64) for c in clusters])
    offsets = np.concatenate([[0], np.cumsum([len(c) for c in clusters])]).astype(np.int64)
    cool2node = np.zeros(nb, dtype=np.int64)
    for i in range(nb):
        name = "%s:%d" % (cool_chrom_names[cool_chrom[i]], cool_start[i])
        cool2node[i] = bin2node.get(name, 0)
    n2c = np.zeros(N + 1, dtype=np.int32)
    for k, v in node2chrom.items():
        n2c[k] = v
    out = dict(chrom_range=chrom_range, members=members, offsets=offsets, edge_adj=edge_adj, bin1=b1.astype(np.int64),
               bin2=b2.astype(np.int64), count=cnt, cool2node=cool2node, node2chrom=n2c, intra=intra, inter=inter)
    path = os.path.join(os.path.dirname(HERE), "tests", "golden", "features_small.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes; edge_adj sum", edge_adj.sum(), "intra sum", np.nansum(intra), "inter sum",
          np.nansum(inter))


if __name__ == "__main__":
    main()
