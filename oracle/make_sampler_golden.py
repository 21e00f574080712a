"""Distribution fixture for the negative sampler, produced by EXECUTING the reference's own
``generate_negative`` (authoring container only).

    python oracle/make_sampler_golden.py        # writes tests/golden/sampler_stats.json

``/root/reference/Code/main.py`` cannot be imported (its body reads ./config.JSON and the Temp files at
import, main.py:516-571), so its FunctionDef nodes are lifted verbatim with ``ast`` and executed in a
namespace that provides the module globals they read -- the recipe of SURVEY.md section 8c.  The
reference sampler consumes numpy / Python global RNG streams, so only DISTRIBUTIONS are comparable with
our counter-based sampler; tests/test_sampler_oracle.py checks those.
"""
import ast
import json
import os
import random
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "_stubs"))
sys.path.insert(0, "/root/reference/Code")
sys.path.insert(0, os.path.dirname(HERE))

import torch  # noqa: E402

from oracle import sampler_oracle as SO  # noqa: E402


def lift_main_functions():
    src = open("/root/reference/Code/main.py").read()
    tree = ast.parse(src)
    fns = [n for n in tree.body if isinstance(n, ast.FunctionDef)]
    mod = ast.Module(body=fns, type_ignores=[])
    ns = {}
    exec("import math, random, os, time\nimport numpy as np\nimport torch\nimport torch.nn.functional as F\n"
         "from torch.nn.utils.rnn import pad_sequence\nfrom utils import *\n", ns)
    exec(compile(mod, "main.py(lifted)", "exec"), ns)
    return ns


def toy_problem(seed=0):
    rng = np.random.default_rng(seed)
    nums = [120, 90, 150]
    starts = np.concatenate([[0], np.cumsum(nums)])
    chrom_range = np.stack([starts[:-1] + 1, starts[1:] + 1], 1).astype(np.int64)
    kmers = {}
    for k in (3, 4, 5):
        rows = []
        while len(rows) < 6000:
            c = int(rng.integers(0, 3))
            a = int(rng.integers(chrom_range[c, 0], chrom_range[c, 1]))
            ids = {a}
            while len(ids) < k:
                ids.add(int(np.clip(a + rng.integers(-12, 13), chrom_range[c, 0], chrom_range[c, 1] - 1)))
            rows.append(sorted(ids))
        kmers[k] = np.unique(np.asarray(rows, dtype=np.int64), axis=0)
    return nums, chrom_range, kmers


def changed_count(pos, neg):
    return len(pos) - len(set(int(v) for v in pos) & set(int(v) for v in neg))


def main():
    ns = lift_main_functions()
    nums, chrom_range, kmers = toy_problem()
    node2chrom = {}
    for c, (s, e) in enumerate(chrom_range):
        for i in range(s, e):
            node2chrom[i] = c
    stats = {"nums": nums, "sizes": {}}
    for k, rows in kmers.items():
        from pybloom_live import BloomFilter
        dicts = [BloomFilter(10) for _ in range(6)]
        for r in rows:
            dicts[k].add(tuple(r))
        ns.update(train_dict=dicts, test_dict=dicts, node2chrom=node2chrom, chrom_range=chrom_range, min_size=k,
                  max_size=k, min_dis=0, task_mode="class", neg_num=3, device=torch.device("cpu"))
        np.random.seed(100 + k)
        random.seed(100 + k)
        hist = np.zeros(k + 1, dtype=np.int64)
        same_chrom = 0
        total = 0
        in_set = 0
        pos_all = rows[np.random.permutation(len(rows))[:2880]]
        for b in range(0, len(pos_all), 96):
            pos = pos_all[b:b + 96]
            x, y, w, s = ns["generate_negative"](pos, "train_dict", torch.ones(len(pos)), neg_num=3)
            neg = x[len(pos):].numpy()
            assert neg.shape[0] == 3 * len(pos)
            for g, nrow in enumerate(neg):
                prow = pos[g // 3]
                hist[changed_count(prow, nrow)] += 1
                total += 1
                same_chrom += int(sorted(node2chrom[int(v)] for v in prow) == sorted(node2chrom[int(v)] for v in nrow))
                in_set += int(tuple(int(v) for v in nrow) in dicts[k])
        stats["sizes"][str(k)] = {"changed_hist": hist.tolist(), "n": total, "same_chrom_multiset": same_chrom,
                                  "in_positive_set": in_set}
        print(k, hist / total, same_chrom / total, in_set)
    out = os.path.join(HERE, "..", "tests", "golden", "sampler_stats.json")
    json.dump(stats, open(out, "w"), indent=1)
    print("wrote", out)


if __name__ == "__main__":
    main()
