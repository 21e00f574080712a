"""Streaming batch scorers: all intra-chromosomal pairs (denoise_contact.py:67-88,147-155) and k-way
tuples (predict_multiway.py:74-87, main.py:482-494).

Pairs never exist on the host: per-node tables are computed once (k = 2 closed form, see
csrc/scorer.cu) and the kernel enumerates (i, j) in ``generate_pair_wise`` order.  Work shards across
ranks by contiguous pair range with no communication.
"""
from __future__ import annotations

import numpy as np
import torch

from ._lib import MatchaError, check, ptr, stream_ptr


def pair_count(lo: int, hi: int, min_dis: int = 0) -> int:
    """Number of pairs generate_pair_wise(chrom) emits for ids [lo, hi): i in [lo, hi), j in [i+min_dis, hi)."""
    from . import _lib
    return int(_lib.load().matcha_pair_count(int(lo), int(hi), int(min_dis)))


def pair_index_to_ij(p, lo, hi, min_dis=0):
    """Host helper (tests / post-processing): global pair index -> (i, j), vectorised numpy."""
    n, md = hi - lo, min_dis
    full = n - md
    p = np.asarray(p, dtype=np.int64)
    # largest r with r*full - r(r-1)/2 <= p
    r = np.floor(((2 * full + 1) - np.sqrt((2 * full + 1) ** 2 - 8.0 * p)) / 2).astype(np.int64)
    for _ in range(2):
        pref = r * full - r * (r - 1) // 2
        r = np.where(pref > p, r - 1, r)
        pref = r * full - r * (r - 1) // 2
        nxt = (r + 1) * full - (r + 1) * r // 2
        r = np.where(nxt <= p, r + 1, r)
    pref = r * full - r * (r - 1) // 2
    return lo + r, lo + r + md + (p - pref)


class PairScorer:
    """All-pairs scorer bound to a trained Classifier (eval mode)."""

    def __init__(self, model):
        self.engine = model._engine()
        self.model = model
        self.refresh()

    def refresh(self):
        """Recompute the per-node tables (call after the weights change)."""
        self.engine.ensure_bound()
        self.D, self.S = self.engine.pair_tables() if self.engine.d == 64 else (None, None)
        m = self.model
        self.cls_w = m.pff_classifier.PWF_Conv0.weight.data.reshape(-1)
        self.cls_b = m.pff_classifier.PWF_Conv0.bias.data.reshape(-1)
        self._packed = {}

    def _pack(self, lo, hi):
        """Operand blocks of the tensor-core scorer for chromosome ids [lo, hi), built once per table refresh."""
        key = (int(lo), int(hi))
        ws = self._packed.get(key)
        if ws is None:
            eng = self.engine
            nbytes = int(eng.lib.matcha_pair_tc_workspace_bytes(key[0], key[1]))
            ws = torch.empty(nbytes, dtype=torch.uint8, device=eng.dev)
            check(eng.lib.matcha_pair_tc_prepare(ptr(self.D), ptr(self.S), ptr(self.cls_w), ptr(self.cls_b), eng.d, key[0], key[1],
                                                 ptr(ws), nbytes, stream_ptr()), "matcha_pair_tc_prepare")
            self._packed[key] = ws
        return ws

    def score_range(self, lo, hi, min_dis=0, p_begin=0, p_end=None, sigmoid=False, out=None, impl="tc"):
        """Scores of pairs [p_begin, p_end) of chromosome ids [lo, hi) -> fp32 tensor on the device.
        impl "tc" = tcgen05 contraction form (default), "simt" = fp32 FMA form (cross-check)."""
        total = pair_count(lo, hi, min_dis)
        p_end = total if p_end is None else p_end
        if not (0 <= p_begin <= p_end <= total):
            raise MatchaError("pair range out of bounds")
        eng = self.engine
        if out is None:
            out = torch.empty(p_end - p_begin, dtype=torch.float32, device=eng.dev)
        if p_end == p_begin:        # empty shard / min_distance beyond the chromosome
            return out
        if eng.d != 64:
            return self._score_range_generic(lo, hi, min_dis, p_begin, p_end, sigmoid, out)
        if impl == "tc":
            ws = self._pack(lo, hi)
            check(eng.lib.matcha_pair_tc_score_range(ptr(ws), int(lo), int(hi), int(min_dis), int(p_begin),
                                                     int(p_end), 1 if sigmoid else 0, ptr(out), stream_ptr()),
                  "matcha_pair_tc_score_range")
        else:
            check(eng.lib.matcha_pair_score_range(ptr(self.D), ptr(self.S), ptr(self.cls_w), ptr(self.cls_b), eng.d,
                                                  int(lo), int(hi), int(min_dis), int(p_begin), int(p_end),
                                                  1 if sigmoid else 0, ptr(out), stream_ptr()), "matcha_pair_score_range")
        return out

    def _score_range_generic(self, lo, hi, min_dis, p_begin, p_end, sigmoid, out, batch=1 << 18):
        """embed_dim != 64: the closed-form kernels are specialised for 64, so pairs go through Classifier.forward as width-2
        tuples (exactly what denoise_contact.py:76-88 does), generated on the device in batches."""
        model = self.model
        model.eval()
        n, full = hi - lo, hi - lo - min_dis
        with torch.no_grad():
            for b in range(p_begin, p_end, batch):
                e = min(b + batch, p_end)
                p = torch.arange(b, e, device=self.engine.dev, dtype=torch.float64)
                r = torch.floor(((2 * full + 1) - torch.sqrt((2 * full + 1) ** 2 - 8.0 * p)) / 2).to(torch.int64)
                p = p.to(torch.int64)
                for _ in range(2):                       # exact integer correction of the float row estimate
                    r = torch.where(r * full - r * (r - 1) // 2 > p, r - 1, r)
                    r = torch.where((r + 1) * full - (r + 1) * r // 2 <= p, r + 1, r)
                pref = r * full - r * (r - 1) // 2
                x = torch.stack([lo + r, lo + r + min_dis + (p - pref)], dim=1)
                o = model(x).view(-1)
                out[b - p_begin:e - p_begin] = torch.sigmoid(o) if sigmoid else o
        return out

    def score_chromosome(self, chrom_id, min_dis=0, sigmoid=False, rank=0, world=1):
        """This rank's contiguous share of chromosome `chrom_id`'s pairs: (p_begin, scores)."""
        lo, hi = (int(v) for v in self.engine.chrom_range[chrom_id])
        total = pair_count(lo, hi, min_dis)
        b, e = total * rank // world, total * (rank + 1) // world
        return b, self.score_range(lo, hi, min_dis, b, e, sigmoid)


def score_tuples(model, samples, batch_size=int(1e4), sigmoid=False, rank=0, world=1):
    """predict() of predict_multiway.py:74-87: batches of `batch_size`, each padded to the longest tuple IN
    THAT BATCH (the score depends on the padded width), sharded by whole batches."""
    model.eval()
    n = len(samples)
    nb = (n + batch_size - 1) // batch_size
    outs = []
    with torch.no_grad():
        for j in range(rank, nb, world):
            chunk = samples[j * batch_size:min((j + 1) * batch_size, n)]
            L = max(len(s) for s in chunk)
            x = np.zeros((len(chunk), L), dtype=np.int64)
            for i, s in enumerate(chunk):
                x[i, :len(s)] = np.asarray(s, dtype=np.int64)
            xt = torch.from_numpy(x).pin_memory().cuda(non_blocking=True)
            o = model(xt)
            outs.append((j, torch.sigmoid(o) if sigmoid else o))
    return outs
