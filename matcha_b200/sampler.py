"""Device hash set of positive k-mers and the GPU negative sampler.

Replaces, on the device, ``utils.build_hash`` (utils.py:75-97), ``neighbor_check`` (main.py:345-346) and
``generate_negative`` (main.py:361-459).  Membership is exact (the reference uses a Bloom filter with a
1e-3 false-positive rate); the candidate streams are counter-based (splitmix64) so
``oracle/sampler_oracle.py`` reproduces every sampled id bit for bit.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import MatchaError, check, ptr, stream_ptr

MAX_WIDTH = 6


def _pow2_at_least(n: int) -> int:
    c = 1
    while c < n:
        c <<= 1
    return c


class KmerHashSet:
    """Open-addressing table of 16-byte slots (packed k-mer, 21 bits per id), load factor <= 0.5."""

    def __init__(self, capacity_hint: int, width: int = 5, device=None):
        if not torch.cuda.is_available():
            raise MatchaError("KmerHashSet needs a CUDA device")
        if not (1 <= width <= MAX_WIDTH):
            raise MatchaError(f"width {width} unsupported (1..{MAX_WIDTH})")
        self.lib = _lib.load()
        self.dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.width = width
        self.capacity = _pow2_at_least(max(1024, 2 * int(capacity_hint)))
        self.table = torch.zeros((self.capacity + 1) * 2, dtype=torch.int64, device=self.dev)   # +1 status slot
        self.count = 0

    def _as_rows(self, kmers):
        if isinstance(kmers, np.ndarray):
            kmers = torch.from_numpy(np.ascontiguousarray(kmers.astype(np.int64)))
        kmers = kmers.to(self.dev, torch.int64)
        if kmers.dim() != 2 or kmers.shape[1] > self.width:
            raise MatchaError(f"k-mers must be [n, <= {self.width}]")
        if kmers.shape[1] < self.width:
            kmers = torch.cat([kmers, kmers.new_zeros(kmers.shape[0], self.width - kmers.shape[1])], 1)
        return kmers.contiguous()

    def insert(self, kmers):
        """kmers [n, k] int64, rows sorted ascending, zero padded to the right (k <= width)."""
        rows = self._as_rows(kmers)
        if rows.shape[0] == 0:
            return self
        if (self.count + rows.shape[0]) * 2 > self.capacity:
            raise MatchaError("hash set over capacity: construct it with a larger capacity_hint")
        check(self.lib.matcha_hashset_insert(ptr(self.table), self.capacity, ptr(rows), rows.shape[0], self.width,
                                             stream_ptr()), "matcha_hashset_insert")
        self.count += rows.shape[0]
        return self

    def overflowed(self) -> bool:
        return bool(self.table[2 * self.capacity].item() != 0)

    def contains(self, kmers):
        rows = self._as_rows(kmers)
        out = torch.empty(rows.shape[0], dtype=torch.uint8, device=self.dev)
        check(self.lib.matcha_hashset_contains(ptr(self.table), self.capacity, ptr(rows), rows.shape[0], self.width,
                                               ptr(out), stream_ptr()), "matcha_hashset_contains")
        return out.bool()


class NegativeSampler:
    """``neg_num`` corrupted copies per positive, same-chromosome replacement, rejected while they are
    positives (main.py:361-428).  ``sample`` returns negatives in the reference's order (positive-major)."""

    def __init__(self, hashset: KmerHashSet, chrom_range, min_dis=0, neg_num=3, seed=0, max_rounds=64):
        self.hs = hashset
        cr = np.asarray(chrom_range, dtype=np.int64)
        self.n_chrom = len(cr)
        self._cs = (C.c_int64 * self.n_chrom)(*[int(v) for v in cr[:, 0]])
        self._ce = (C.c_int64 * self.n_chrom)(*[int(v) for v in cr[:, 1]])
        self.min_dis, self.neg_num, self.seed, self.max_rounds = int(min_dis), int(neg_num), int(seed), int(max_rounds)
        self.step = 0

    def sample(self, pos, out=None, valid=None, rounds=None, step=None):
        """pos int64 [P, L] on the device (L == hash-set width) -> (neg [P*neg_num, L], valid uint8 [P*neg_num])."""
        hs = self.hs
        if pos.dim() != 2 or pos.shape[1] != hs.width or pos.dtype != torch.int64 or not pos.is_contiguous():
            raise MatchaError(f"pos must be a contiguous int64 [P, {hs.width}] tensor")
        P, L = pos.shape
        n = P * self.neg_num
        if out is None:
            out = torch.empty(n, L, dtype=torch.int64, device=hs.dev)
        if valid is None:
            valid = torch.empty(n, dtype=torch.uint8, device=hs.dev)
        if step is None:
            step = self.step
            self.step += 1
        check(hs.lib.matcha_neg_sample(ptr(hs.table), hs.capacity, ptr(pos), P, L, self.neg_num, self._cs, self._ce,
                                       self.n_chrom, self.min_dis, self.seed, step, self.max_rounds, ptr(out),
                                       ptr(valid), ptr(rounds), stream_ptr()), "matcha_neg_sample")
        return out, valid
