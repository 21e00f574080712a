"""Small end-to-end workload for compute-sanitizer (memcheck / racecheck / synccheck / initcheck): every kernel family
once -- 2 training steps (tensor-core path, 1,280 hyperedges = 6,400 tokens, recon on), eval forward, pair tables + all-pairs
scoring of one chromosome, denoise post-processing, metrics, feature construction, k-mer counting, negative sampling, and one
embed_dim-128 training step (general tcgen05 contraction kernel).  MATCHA_ENC_PIPE=1 puts the pipelined encoder forward on
this small data set too (it is the default only for wide feature rows).
    MATCHA_ENC_PIPE=1 compute-sanitizer --tool memcheck python scripts/dev/sanitize_step.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from matcha_b200.denoise import QuantileUniform, denoise_matrix  # noqa: E402
from matcha_b200.features import adjacency_from_clusters, corrcoef_features, zscore_positive_rows_  # noqa: E402
from matcha_b200.kmers import clusters_to_csr, count_kmers  # noqa: E402
from matcha_b200.metrics import binary_metrics  # noqa: E402
from matcha_b200.sampler import KmerHashSet, NegativeSampler  # noqa: E402
from matcha_b200.scorer import PairScorer, pair_count  # noqa: E402
from matcha_b200.synthetic import build_model, make_dataset  # noqa: E402
from matcha_b200.trainer import Trainer  # noqa: E402

ds = make_dataset("cfg1", kmers_per_size=5000, seed=0)
model = build_model(ds, seed=1)
hs = KmerHashSet(len(ds["dict"]), width=5).insert(ds["dict"])
tr = Trainer(model, NegativeSampler(hs, ds["chrom_range"], min_dis=0, neg_num=3, seed=2), alpha=1.0, beta=0.5, seed=3)
P = 320
pos = torch.from_numpy(ds["positives"][:2 * P]).cuda()
w = torch.from_numpy(ds["pos_weight"][:2 * P]).cuda()
tr.step(pos[:P], w[:P], pos[P:], w[P:])
tr.step(pos[P:], w[P:])
torch.cuda.synchronize()
print("train", tr.mean_losses())
model.eval()
with torch.no_grad():
    lg = model(pos[:500])
sc = PairScorer(model)
lo, hi = (int(v) for v in ds["chrom_range"][0])
n = hi - lo
proba = sc.score_range(lo, hi, 0, sigmoid=True)
assert proba.numel() == pair_count(lo, hi, 0)
origin = torch.rand(n, n, device="cuda")
my, pix = denoise_matrix(proba, origin, n, 0, QuantileUniform(1000, random_state=0), want_pixels=True)
m = binary_metrics((torch.rand(5000, device="cuda") < 0.3).float(), torch.rand(5000, device="cuda"), torch.randint(2, 6, (5000,), device="cuda"), 5)
rng = np.random.default_rng(0)
clusters = [np.unique(rng.integers(1, ds["N"] + 1, 6)) for _ in range(500)]
mem, off = clusters_to_csr(clusters)
adj = adjacency_from_clusters(mem, off, ds["N"])
feats = corrcoef_features(adj.float(), ds["chrom_range"])
zscore_positive_rows_(adj.float().contiguous())
rows, freq = count_kmers(mem, off, 3, 0, 25, 1)
# embed_dim 128: one training step with every contraction on the general tcgen05 kernel (csrc/gemm_tcg.cu)
ds128 = dict(ds)
ds128["d"] = 128
model128 = build_model(ds128, seed=1)
model128.train()
x128 = torch.from_numpy(ds["positives"][:600]).cuda()
pred, rl = model128(x128, return_recon=True)
(torch.nn.functional.binary_cross_entropy_with_logits(pred, torch.ones_like(pred)) + 0.5 * rl.sum()).backward()
torch.cuda.synchronize()
print("ok", float(lg.mean()), float(my.mean()), m["all"][:2], len(rows), float(pred.mean()))
