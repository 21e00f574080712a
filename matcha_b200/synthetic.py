"""Seeded synthetic SPRITE-like inputs of the shapes BASELINE.json names.

Nothing is downloaded: bins follow the hg38 chromosome sizes with the reference's binning formula
(process.py:21-36: n_c = ceil(size / res) + 1, ids 1-based and contiguous per chromosome); k-mers are
drawn from a distance-decaying generative model (mostly intra-chromosomal, |offset|^-1), frequencies are
2 + Geometric(1/2); features are built exactly as main.py:572-577 (corrcoef of the intra adjacency block,
NaN -> 0) and Modules.py:147-152 (row-wise z-score of the positive inter-chromosomal entries).
"""
from __future__ import annotations

import math

import numpy as np

HG38 = {
    "chr1": 248956422, "chr2": 242193529, "chr3": 198295559, "chr4": 190214555, "chr5": 181538259,
    "chr6": 170805979, "chr7": 159345973, "chr8": 145138636, "chr9": 138394717, "chr10": 133797422,
    "chr11": 135086622, "chr12": 133275309, "chr13": 114364328, "chr14": 107043718, "chr15": 101991189,
    "chr16": 90338345, "chr17": 83257441, "chr18": 80373285, "chr19": 58617616, "chr20": 64444167,
    "chr21": 46709983, "chr22": 50818468, "chrX": 156040895,
}
WHOLE_GENOME = ["chr%d" % i for i in range(1, 23)] + ["chrX"]

CONFIGS = {
    # name: (chromosomes, resolution, embed_dim)   -- BASELINE.json "configs", in order
    "cfg1": (["chr1", "chr2"], 1_000_000, 64),
    "cfg2": (WHOLE_GENOME, 1_000_000, 64),
    "cfg3": (WHOLE_GENOME, 100_000, 64),
    "cfg4": (["chr1"], 10_000, 64),
    "cfg5": (WHOLE_GENOME, 50_000, 128),
}


def chrom_bins(chroms, res):
    nums = [int(math.ceil(HG38[c] / res)) + 1 for c in chroms]
    starts = np.concatenate([[0], np.cumsum(nums)])
    chrom_range = np.stack([starts[:-1] + 1, starts[1:] + 1], 1).astype(np.int64)
    return nums, chrom_range


def draw_kmers(rng, k, m, nums, chrom_range, p_inter=0.05):
    """m candidate k-mers (rows sorted, unique ids); returns the unique valid rows."""
    nums_a = np.asarray(nums)
    c = rng.choice(len(nums), size=m, p=nums_a / nums_a.sum())
    n_c = nums_a[c]
    anchor = (rng.random(m) * n_c).astype(np.int64)
    cols = [anchor + chrom_range[c, 0]]
    for _ in range(k - 1):
        # |offset| ~ 1/|offset| on [1, n_c): inverse-CDF of the continuous law, floored
        mag = np.floor(np.exp(rng.random(m) * np.log(np.maximum(n_c, 2)))).astype(np.int64)
        off = np.where(rng.random(m) < 0.5, -mag, mag)
        local = np.clip(anchor + off, 0, n_c - 1)
        ids = local + chrom_range[c, 0]
        inter = rng.random(m) < p_inter
        if inter.any():
            c2 = rng.choice(len(nums), size=int(inter.sum()), p=nums_a / nums_a.sum())
            ids[inter] = (rng.random(int(inter.sum())) * nums_a[c2]).astype(np.int64) + chrom_range[c2, 0]
        cols.append(ids)
    rows = np.sort(np.stack(cols, 1), axis=1)
    ok = (np.diff(rows, axis=1) > 0).all(1)
    return np.unique(rows[ok], axis=0)


def rank_quantile(freq, rng):
    """Uniform quantile of each frequency (ties broken randomly) -- the role of QuantileTransformer at main.py:555."""
    order = np.lexsort((rng.random(len(freq)), freq))
    q = np.empty(len(freq), dtype=np.float64)
    q[order] = (np.arange(len(freq)) + 0.5) / len(freq)
    return q


def make_dataset(config="cfg2", kmers_per_size=200_000, sizes=(2, 3, 4, 5), seed=0, with_inter=True,
                 q_pos=0.6, q_dict=0.4, neg_num=3):
    chroms, res, d = CONFIGS[config] if isinstance(config, str) else config
    rng = np.random.default_rng(seed)
    nums, chrom_range = chrom_bins(chroms, res)
    N = int(sum(nums))
    kmers, freq = {}, {}
    for k in sizes:
        rows = draw_kmers(rng, k, int(kmers_per_size * 1.15), nums, chrom_range)[:kmers_per_size]
        kmers[k] = rows
        freq[k] = 2 + rng.geometric(0.5, size=len(rows))
    # co-occurrence adjacency from the k-mers (weighted by frequency)
    adj = np.zeros((N, N), dtype=np.float32)
    for k in sizes:
        r, f = kmers[k] - 1, freq[k].astype(np.float32)
        for a in range(k):
            for b in range(a + 1, k):
                np.add.at(adj, (r[:, a], r[:, b]), f)
    adj = adj + adj.T
    feats = []
    intra_mask = np.zeros((N, N), dtype=bool)
    for (s, e) in chrom_range:
        blk = adj[s - 1:e - 1, s - 1:e - 1]
        with np.errstate(invalid="ignore", divide="ignore"):
            cc = np.corrcoef(blk).astype(np.float32)            # main.py:574
        cc[np.isnan(cc)] = 0.0
        feats.append(cc)
        intra_mask[s - 1:e - 1, s - 1:e - 1] = True
    inter = None
    if with_inter and len(chroms) > 1:
        inter = np.where(intra_mask, 0.0, adj).astype(np.float32)
    # attribute table, main.py:497-512
    C = len(nums)
    attrs = []
    for i, n in enumerate(nums):
        ch = np.zeros((n, C), dtype=np.float32)
        ch[:, i] = 1
        coor = (np.arange(n, dtype=np.float32) / nums[0]).reshape(-1, 1)
        attrs.append(np.concatenate([ch, coor], -1))
    attr = np.concatenate([np.zeros((1, C + 1), dtype=np.float32), np.concatenate(attrs, 0)], 0)
    # positives (quantile > q_pos) with weights as main.py:555-556,594-595; dictionary (quantile > q_dict)
    pos, pos_w, dict_rows = [], [], []
    L = max(sizes)
    for k in sizes:
        q = rank_quantile(freq[k], rng)
        rows = np.concatenate([kmers[k], np.zeros((len(kmers[k]), L - k), dtype=np.int64)], 1)
        pos.append(rows[q > q_pos]); pos_w.append(q[q > q_pos])
        dict_rows.append(rows[q > q_dict])
    pos, pos_w = np.concatenate(pos), np.concatenate(pos_w).astype(np.float32)
    pos_w = pos_w / pos_w.mean() * neg_num
    return {"config": config, "chroms": chroms, "res": res, "d": d, "nums": nums, "chrom_range": chrom_range, "N": N,
            "kmers": kmers, "freq": freq, "features": feats, "inter": inter, "attr": attr,
            "positives": pos, "pos_weight": pos_w, "dict": np.concatenate(dict_rows), "L": L}


def build_model(ds, seed=1):
    """Construct MultipleEmbedding + Classifier exactly as main.py:609-623 does."""
    import torch
    from . import hyper_sagnn as M
    torch.manual_seed(seed)
    inter = None if ds["inter"] is None else ds["inter"].copy()
    ne = M.MultipleEmbedding(ds["features"], ds["d"], False, np.cumsum(ds["nums"]), ds["chrom_range"], inter)
    model = M.Classifier(n_head=8, d_model=ds["d"], d_k=ds["d"], d_v=ds["d"], node_embedding=ne, diag_mask=True,
                         bottle_neck=ds["d"], attribute_dict=ds["attr"])
    return model.to(M.device)


def write_temp_dir(ds, temp_dir, config_path=None):
    """Write `ds` in the reference's on-disk formats (process.py:36-39,175-176; generate_kmers.py:140-141):
    chrom_range.npy, node2chrom.npy, bin2node.npy, node2bin.npy, all_<k>_counter.npy,
    all_<k>_freq_counter.npy, intra_adj.npy, inter_adj.npy -- plus a config.JSON with the reference's keys."""
    import json
    import os
    os.makedirs(temp_dir, exist_ok=True)
    cr, N, res = ds["chrom_range"], ds["N"], ds["res"]
    np.save(os.path.join(temp_dir, "chrom_range.npy"), cr)
    node2chrom, bin2node, node2bin = {}, {}, {}
    for c, (s, e) in enumerate(cr):
        for i in range(int(s), int(e)):
            node2chrom[i] = c
            name = "%s:%d" % (ds["chroms"][c], (i - int(s)) * res)
            bin2node[name], node2bin[i] = i, name
    np.save(os.path.join(temp_dir, "node2chrom.npy"), node2chrom, allow_pickle=True)
    np.save(os.path.join(temp_dir, "bin2node.npy"), bin2node, allow_pickle=True)
    np.save(os.path.join(temp_dir, "node2bin.npy"), node2bin, allow_pickle=True)
    for k, rows in ds["kmers"].items():
        np.save(os.path.join(temp_dir, "all_%d_counter.npy" % k), rows)
        np.save(os.path.join(temp_dir, "all_%d_freq_counter.npy" % k), ds["freq"][k])
    adj = np.zeros((N, N), dtype=np.float32)
    for k, rows in ds["kmers"].items():
        r, f = rows - 1, ds["freq"][k].astype(np.float32)
        for a in range(k):
            for b in range(a + 1, k):
                np.add.at(adj, (r[:, a], r[:, b]), f)
    adj = adj + adj.T
    intra = np.zeros_like(adj)
    for (s, e) in cr:
        intra[s - 1:e - 1, s - 1:e - 1] = adj[s - 1:e - 1, s - 1:e - 1]
    np.save(os.path.join(temp_dir, "intra_adj.npy"), intra)
    np.save(os.path.join(temp_dir, "inter_adj.npy"), adj - intra)
    cfg = {"cluster_path": "synthetic.cluster", "mcool_path": "synthetic.mcool", "resolution": res,
           "chrom_list": list(ds["chroms"]), "chrom_size": "synthetic", "temp_dir": temp_dir, "max_cluster_size": 25,
           "min_distance": 0, "k-mer_size": sorted(int(k) for k in ds["kmers"]), "min_freq_cutoff": 2,
           "quantile_cutoff_for_positive": 0.6, "quantile_cutoff_for_unlabel": 0.4, "embed_dim": ds["d"]}
    if config_path:
        json.dump(cfg, open(config_path, "w"), indent=1)
    return cfg
