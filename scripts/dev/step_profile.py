"""A few training steps of the bench configuration (cfg2, 16,384 hyperedges) for profiler captures:
    ncu -k regex:<kernel> -c 1 --set full --import-source on -o gpurun_out/x python scripts/dev/step_profile.py [steps] [workload]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from matcha_b200.sampler import KmerHashSet, NegativeSampler  # noqa: E402
from matcha_b200.synthetic import build_model, make_dataset  # noqa: E402
from matcha_b200.trainer import Trainer  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
workload = sys.argv[2] if len(sys.argv) > 2 else "cfg2"
ds = make_dataset(workload, kmers_per_size=100_000, seed=0)
model = build_model(ds, seed=1)
hs = KmerHashSet(len(ds["dict"]), width=5).insert(ds["dict"])
tr = Trainer(model, NegativeSampler(hs, ds["chrom_range"], min_dis=0, neg_num=3, seed=2), alpha=1.0, beta=0.001, seed=3)
P = 4096
perm = np.random.RandomState(7).permutation(len(ds["positives"]))
pos = torch.from_numpy(ds["positives"][perm]).cuda()
w = torch.from_numpy(ds["pos_weight"][perm]).cuda()
for i in range(steps):
    tr.step(pos[i * P:(i + 1) * P], w[i * P:(i + 1) * P])
torch.cuda.synchronize()
print("ok", tr.mean_losses())
