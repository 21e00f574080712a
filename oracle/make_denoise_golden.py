"""Golden vectors for the denoise post-processing, produced by EXECUTING the reference's own script body (authoring
container only).

    python oracle/make_denoise_golden.py        # writes tests/golden/denoise_small.npz

``Code/denoise_contact.py`` is a script that imports matplotlib / seaborn / h5py (absent here) and loads a model at
import, so nothing of it can be imported.  Its two functions ``proba2matrix`` (:31-61) and ``generate_pair_wise``
(:67-74) and its per-chromosome ``for`` loop (:147-229) are lifted VERBATIM with ``ast`` and executed in a namespace that
provides the globals they read: a stub ``predict`` returning the logits we choose, recording stubs for ``plt`` / ``sns``
(the heat-map calls receive the final matrices), the real sklearn ``QuantileTransformer`` the script constructs (:107).
"""
import ast
import math
import os
import sys
from unittest import mock

import numpy as np
import torch
import torch.nn.functional as F
from sklearn.preprocessing import QuantileTransformer

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/Code/denoise_contact.py"


def lift():
    tree = ast.parse(open(REF).read())
    fns = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in ("proba2matrix", "generate_pair_wise")]
    loops = [n for n in tree.body if isinstance(n, ast.For) and "generate_pair_wise" in ast.dump(n)]
    assert len(fns) == 2 and len(loops) == 1
    return ast.Module(body=fns, type_ignores=[]), ast.Module(body=loops, type_ignores=[])


def run_reference(chrom_range, min_dis, origin, logits_per_chrom, chrom_name):
    fns, loop = lift()
    heat = []
    sns = mock.MagicMock()
    sns.heatmap.side_effect = lambda m, *a, **k: (heat.append(np.array(m, copy=True)), mock.MagicMock())[1]
    calls = iter(logits_per_chrom)
    ns = dict(np=np, torch=torch, F=F, math=math, tqdm=lambda x, *a, **k: x, chrom_range=chrom_range, min_dis=min_dis,
              origin=origin, transformer=QuantileTransformer(n_quantiles=1000, output_distribution="uniform"),
              task_mode="class", classifier_model=None, predict=lambda model, pw: next(calls).reshape(-1, 1),
              plt=mock.MagicMock(), sns=sns, vmin=-0.0, vmax=1.0, chrom_name=chrom_name, bin_id1=[], bin_id2=[], balanced=[])
    exec(compile(fns, "denoise_contact.py(functions)", "exec"), ns)
    exec(compile(loop, "denoise_contact.py(loop)", "exec"), ns)
    # two heat maps per chromosome: `my` (:198) then `origin_part` (:222)
    return ns["bin_id1"], ns["bin_id2"], ns["balanced"], heat[0::2], heat[1::2]


def main():
    rng = np.random.default_rng(0)
    nums = [40, 55]
    starts = np.concatenate([[0], np.cumsum(nums)])
    chrom_range = np.stack([starts[:-1] + 1, starts[1:] + 1], 1).astype(np.int64)
    N = int(sum(nums))
    origin = np.zeros((N, N), dtype="float32")
    for (s, e) in chrom_range:
        n = e - s
        blk = rng.poisson(2.0, size=(n, n)).astype("float32") * (rng.random((n, n)) < 0.5)
        blk = np.triu(blk) + np.triu(blk, 1).T
        dead = rng.choice(n, size=3, replace=False)            # unmappable bins: empty rows / columns ("gaps", :169-170)
        blk[dead, :] = 0.0
        blk[:, dead] = 0.0
        origin[s - 1:e - 1, s - 1:e - 1] = blk
    out = {"chrom_range": chrom_range, "origin": origin}
    for ci, min_dis in enumerate((0, 2)):
        logits = []
        for (s, e) in chrom_range:
            total = sum(len(range(i + min_dis, e)) for i in range(s, e))          # pairs generate_pair_wise emits (:67-74)
            logits.append(rng.normal(0.0, 2.0, size=total).astype("float32"))
        b1, b2, bal, my, orig = run_reference(chrom_range, min_dis, origin, [l.copy() for l in logits], ["chrA", "chrB"])
        out[f"min_dis/{ci}"] = np.int64(min_dis)
        for c in range(len(nums)):
            out[f"logits/{ci}/{c}"] = logits[c]
            out[f"bin1/{ci}/{c}"], out[f"bin2/{ci}/{c}"] = np.asarray(b1[c]), np.asarray(b2[c])
            out[f"balanced/{ci}/{c}"] = np.asarray(bal[c])
            out[f"my/{ci}/{c}"] = my[c]
        print("min_dis", min_dis, [m.shape for m in my], [float(m.max()) for m in my])
    path = os.path.join(os.path.dirname(HERE), "tests", "golden", "denoise_small.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
