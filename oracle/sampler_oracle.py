"""CPU ORACLE (TEST INFRASTRUCTURE) for the negative sampler and the positive-set membership test.

Restates, with an exact Python ``set`` and the same splitmix64 candidate streams as
``matcha_b200/csrc/sampler.cu``, the semantics of the reference's ``generate_negative``
(main.py:361-428) and ``neighbor_check`` (main.py:345-346):

  * per negative, the set of positions to corrupt is drawn once: each of the k positions independently
    with probability 1/2, redrawn while empty.  (The reference draws a count from Binomial(k, 1/2)
    conditioned on > 0, main.py:371-372, then a uniform subset of that size, :389 -- the same law.)
  * every candidate round replaces those positions by a uniform bin of the SAME chromosome
    (main.py:402-407), sorts, and is rejected if two ids coincide (:410-414), if an adjacent gap is
    <= min_distance (:416-421) or if the tuple is a known positive (:392);
  * the reference loops forever; after ``max_rounds`` rejected rounds we emit the positive itself and
    flag the row invalid.

Third-party boundary: the reference's set is ``pybloom_live.BloomFilter`` (version unpinned, not in the
tree, not installed) -- **parity unpinned** at that boundary.  The contract adopted is exact membership.
"""
from __future__ import annotations

import numpy as np

from .hypersagnn_oracle import _M64, splitmix64


def kmer_key(row):
    return tuple(int(v) for v in row if int(v) != 0)


def build_set(kmers):
    return {kmer_key(r) for r in np.asarray(kmers)}


def sample_negatives(pos, positive_set, chrom_range, neg_num=3, min_dis=0, seed=0, step=0, max_rounds=64):
    """pos int64 [P, L] zero padded -> (neg [P*neg_num, L], valid [P*neg_num] uint8, rounds [P*neg_num])."""
    pos = np.asarray(pos, dtype=np.int64)
    P, L = pos.shape
    cr = [(int(s), int(e)) for s, e in np.asarray(chrom_range)]
    base = splitmix64((seed ^ splitmix64(step)) & _M64)
    neg = np.zeros((P * neg_num, L), dtype=np.int64)
    valid = np.zeros(P * neg_num, dtype=np.uint8)
    rounds = np.zeros(P * neg_num, dtype=np.int32)
    for g in range(P * neg_num):
        p = [int(v) for v in pos[g // neg_num]]
        k = 0
        for i, v in enumerate(p):
            if v != 0:
                k = i + 1
        key = splitmix64((base + g) & _M64)
        ctr = 0
        cmask = 0
        tries = 0
        while tries < 64 and cmask == 0:
            cmask = splitmix64((key + ctr) & _M64) & ((1 << k) - 1)
            ctr += 1
            tries += 1
        if cmask == 0:
            cmask = 1
        rng = {}
        for i in range(k):
            if (cmask >> i) & 1:
                rng[i] = (0, 0)
                for (s, e) in cr:
                    if s <= p[i] < e:
                        rng[i] = (s, e)
        accepted, rnd, t = False, 0, list(p)
        while rnd < max_rounds and not accepted:
            t = list(p)
            for i in range(k):
                if (cmask >> i) & 1 and rng[i][1] > rng[i][0]:
                    u = splitmix64((key + ctr) & _M64) >> 32
                    ctr += 1
                    s, e = rng[i]
                    t[i] = s + ((u * (e - s)) >> 32)
            head = sorted(t[:k])
            t = head + [0] * (L - k)
            ok = all((head[i + 1] - head[i]) != 0 and (head[i + 1] - head[i]) > min_dis for i in range(k - 1))
            if ok and tuple(head) not in positive_set:
                accepted = True
            rnd += 1
        neg[g] = t if accepted else p
        valid[g] = 1 if accepted else 0
        rounds[g] = rnd
    return neg, valid, rounds


def reference_style_negatives(pos, positive_set, chrom_range, node2chrom, neg_num=3, min_dis=0, rng=None, pyrandom=None):
    """Line-by-line restatement of generate_negative's sampling loop (main.py:369-428) with its own RNG
    calls (numpy binomial/choice + Python random), used only to compare DISTRIBUTIONS with the counter
    sampler above (tests/test_sampler_oracle.py)."""
    import math
    import random as _random
    rng = rng or np.random
    pyrandom = pyrandom or _random
    pos = [np.asarray([v for v in r if v != 0]) for r in np.asarray(pos)]
    sizes = sorted({len(r) for r in pos})
    change_num_list = {}
    for s in sizes:
        cn = rng.binomial(s, 0.5, int(len(pos) * (math.ceil(neg_num) * 2)))
        change_num_list[s] = list(cn[cn != 0])
    out = []
    for sample in pos:
        for _ in range(neg_num):
            change_num = change_num_list[len(sample)].pop()
            changes = rng.choice(np.arange(len(sample)), change_num, replace=False)
            temp = np.copy(sample)
            while tuple(int(v) for v in temp) in positive_set:
                temp = np.copy(sample)
                for change in changes:
                    start, end = chrom_range[node2chrom[int(temp[change])]]
                    temp[change] = int(math.floor((end - start) * pyrandom.random())) + start
                temp = list(set(int(v) for v in temp))
                if len(temp) < len(sample):
                    temp = np.copy(sample)
                    continue
                temp.sort()
                if min(temp[k + 1] - temp[k] for k in range(len(temp) - 1)) <= min_dis:
                    temp = np.copy(sample)
            out.append(list(temp))
    return out
