"""Stated-tolerance bf16 mode of the attention block (BASELINE north_star: "scores ... within rtol 1e-4 (fp32) or a stated bf16
tolerance"): the tcgen05 contractions of the fused attention forward / backward issue ONE bf16 product per step (hi * hi)
instead of the three of the fp32-accurate split.  Reports, on the bench configuration (cfg2, 16,384 hyperedges per step):
logit / gradient error of one training step against the fp32-accurate mode on identical inputs and dropout masks, the
attention kernel times of both modes, and validation AUROC / AUPR after the same short training run in both modes.

    python scripts/bf16_mode_report.py [train_steps] > profiles/r02_bf16_mode.json"""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from matcha_b200 import _lib  # noqa: E402
from matcha_b200.metrics import binary_metrics  # noqa: E402
from matcha_b200.sampler import KmerHashSet, NegativeSampler  # noqa: E402
from matcha_b200.synthetic import build_model, make_dataset  # noqa: E402
from matcha_b200.trainer import Trainer  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 300
lib = _lib.load()
ds = make_dataset("cfg2", kmers_per_size=200_000, seed=0)
hs = KmerHashSet(len(ds["dict"]), width=5).insert(ds["dict"])
P = 4096
perm = np.random.RandomState(7).permutation(len(ds["positives"]))
n_val = 8192
val_pos = torch.from_numpy(ds["positives"][perm[:n_val]]).cuda()
pos = torch.from_numpy(ds["positives"][perm[n_val:]]).cuda()
w = torch.from_numpy(ds["pos_weight"][perm[n_val:]]).cuda()
nb = len(pos) // P


def profile_read():
    n = lib.matcha_profile_labels()
    tms, calls, kern = (C.c_float * n)(), (C.c_int64 * n)(), (C.c_int64 * n)()
    _lib.check(lib.matcha_profile_read(tms, calls, kern, n), "profile_read")
    return {lib.matcha_profile_label_name(i).decode(): tms[i] / max(1, calls[i]) for i in range(n) if calls[i]}


out = {}
one = {}
for passes in (3, 1):
    lib.matcha_set_mma_passes(passes)
    model = build_model(ds, seed=1)
    tr = Trainer(model, NegativeSampler(hs, ds["chrom_range"], min_dis=0, neg_num=3, seed=2), alpha=1.0, beta=0.001, seed=3,
                 fused_boundary=False)
    tr.step(pos[:P], w[:P])
    torch.cuda.synchronize()
    one[passes] = (tr.logits.clone(), tr.e.gflat[:tr.e.n_flat].clone(), tr.loss_out.clone())
    # timing + short training run
    for i in range(1, 6):
        tr.step(pos[(i % nb) * P:(i % nb + 1) * P], w[(i % nb) * P:(i % nb + 1) * P])
    lib.matcha_profile_enable(1)
    for i in range(6, 26):
        tr.step(pos[(i % nb) * P:(i % nb + 1) * P], w[(i % nb) * P:(i % nb + 1) * P])
    torch.cuda.synchronize()
    lib.matcha_profile_enable(0)
    prof = profile_read()
    for i in range(26, steps):
        tr.step(pos[(i % nb) * P:(i % nb + 1) * P], w[(i % nb) * P:(i % nb + 1) * P])
    model.eval()
    vs = NegativeSampler(hs, ds["chrom_range"], min_dis=0, neg_num=3, seed=99)
    neg, valid = vs.sample(val_pos.contiguous())
    x = torch.cat([val_pos, neg[valid.bool()]])
    y = torch.cat([torch.ones(len(val_pos)), torch.zeros(int(valid.sum()))]).cuda()
    lib.matcha_set_mma_passes(3)                       # evaluate both trained models with the fp32-accurate forward
    with torch.no_grad():
        pr = torch.sigmoid(model(x)).view(-1)
    m = binary_metrics(y, pr, (x != 0).sum(1), 5)
    out["passes_%d" % passes] = {"attn_fwd_ms": prof.get("attn_fwd"), "attn_bwd_ms": prof.get("attn_bwd"),
                                 "losses_after_run": tr.mean_losses(), "val_auroc": m["all"][0], "val_aupr": m["all"][1],
                                 "val_auroc_by_size": {str(k): v[0] for k, v in m.items() if k != "all"}}
    del tr, model
lib.matcha_set_mma_passes(3)
l3, g3, _ = one[3]
l1, g1, _ = one[1]
out["one_step_bf16_vs_fp32_accurate"] = {
    "max_abs_dlogit": float((l1 - l3).abs().max()), "max_abs_logit": float(l3.abs().max()),
    "rel_l2_logits": float((l1 - l3).norm() / l3.norm()),
    "grad_rel_l2": float((g1 - g3).norm() / g3.norm()), "grad_max_abs_diff_over_max": float((g1 - g3).abs().max() / g3.abs().max())}
out["config"] = {"workload": "cfg2, 4096 positives + 12288 negatives per step, width 5", "train_steps": steps,
                 "validation": "%d held-out positives + their sampled negatives, evaluated with the fp32-accurate forward" % n_val}
t3 = out["passes_3"]["attn_fwd_ms"] + out["passes_3"]["attn_bwd_ms"]
t1 = out["passes_1"]["attn_fwd_ms"] + out["passes_1"]["attn_bwd_ms"]
T = P * 4 * 5
out["roofline"] = {"algorithmic_flops_fwd_bwd": 3 * 2.0 * T * 1536 * 64,
                   "fp32_accurate": {"ms": t3, "TFLOPs": 3 * 2.0 * T * 1536 * 64 / (t3 * 1e-3) / 1e12},
                   "bf16": {"ms": t1, "TFLOPs": 3 * 2.0 * T * 1536 * 64 / (t1 * 1e-3) / 1e12}}
print(json.dumps(out, indent=1))
