"""Fused training step: the loop body of main.py:155-188 without host round trips.

    negatives (GPU sampler) -> x = [pos; neg] -> Classifier forward (train mode) -> weighted BCE-with-logits
    -> alpha * bce + beta * recon -> backward -> (NCCL all-reduce of the flat gradient buffer) -> AdamW

Every stage is one or a few launches of our own kernels on the current stream; nothing reads back to the
host (the reference does two `.item()` per step plus one sync per chromosome).  Data-parallel: every rank
holds a replica, takes its own slice of positives, and the flat fp32 gradient buffer is all-reduced once
per step (torch.distributed / NCCL over NVLink); the activity flags ride in the tail of the same buffer.
"""
from __future__ import annotations

import numpy as np
import torch

from ._lib import check, ptr, stream_ptr
from .engine import FlatAdamW
from .parallel import allreduce_grads_and_flags
from .sampler import NegativeSampler


class Trainer:
    def __init__(self, model, sampler: NegativeSampler, alpha=1.0, beta=0.001, lr=1e-3, weight_decay=0.01,
                 seed=0, world_size=1, rank=0, recon_rng=None):
        model.train()
        self.model = model
        self.e = model._engine()
        self.e.ensure_bound()
        # dropout streams differ per rank; the recon-chromosome stream (below) is shared by all ranks
        self.e.seed_base = (int(seed) + 0x632BE59BD9B4E019 * int(rank)) & ((1 << 64) - 1)
        self.sampler = sampler
        self.alpha, self.beta = float(alpha), float(beta)
        self.opt = FlatAdamW(self.e, lr=lr, weight_decay=weight_decay)
        self.world = int(world_size)
        self.neg_num = sampler.neg_num
        # the per-step chromosome draw of Modules.py:192 -- one shared stream so all ranks draw the same one
        self.recon_rng = recon_rng or np.random.RandomState(seed)
        self.loss_out = torch.zeros(3, dtype=torch.float32, device=self.e.dev)
        self.loss_sum = torch.zeros(3, dtype=torch.float32, device=self.e.dev)
        self.steps = 0
        self._buf_P = -1

    def _buffers(self, P, L):
        if self._buf_P == (P, L):
            return
        dev, n = self.e.dev, P * (1 + self.neg_num)
        self.x = torch.zeros(n, L, dtype=torch.int64, device=dev)
        self.y = torch.zeros(n, dtype=torch.float32, device=dev)
        self.y[:P] = 1.0
        self.w = torch.ones(n, dtype=torch.float32, device=dev)
        self.valid = torch.empty(n - P, dtype=torch.uint8, device=dev)
        self.logits = torch.empty(n, dtype=torch.float32, device=dev)
        self.dlogit = torch.empty(n, dtype=torch.float32, device=dev)
        self.recon = torch.zeros(1, dtype=torch.float32, device=dev)
        self._buf_P = (P, L)

    def step(self, pos, pos_w):
        """pos int64 [P, L] (device, rows sorted, zero padded), pos_w fp32 [P] (device).  Returns nothing:
        running losses accumulate on the device in `loss_sum` ({bce, recon, total})."""
        e, lib = self.e, self.e.lib
        P, L = pos.shape
        self._buffers(P, L)
        n = P * (1 + self.neg_num)
        self.x[:P].copy_(pos)
        self.w[:P].copy_(pos_w)
        self.sampler.sample(pos, out=self.x[P:], valid=self.valid)
        self.w[P:].copy_(self.valid)               # exhausted rows (emitted as the positive itself) get weight 0
        rchrom = int(self.recon_rng.randint(0, e.C)) if (self.beta != 0.0 and e.desc.inter) else -1
        seed = e.next_seed()
        e.prepare()
        e.run_forward(self.x, True, seed, rchrom, logits=self.logits, recon=self.recon)
        check(lib.matcha_bce_loss(ptr(self.logits), ptr(self.y), ptr(self.w), n, self.alpha, self.beta, ptr(self.recon),
                                  ptr(self.dlogit), ptr(self.loss_out), stream_ptr()), "matcha_bce_loss")
        e.gflat.zero_()
        e.run_backward(self.x, seed, rchrom, self.dlogit, self.beta)
        # one collective per step: gradients + activity flags (no-op when world == 1)
        scale = allreduce_grads_and_flags(e.gflat, e.n_flat, e.active, self.world)
        self.opt.step(grad_scale=scale)
        self.loss_sum += self.loss_out
        self.steps += 1
        return n

    def mean_losses(self):
        """Host read of the running means (one sync; call once per epoch, as main.py:197 does)."""
        v = (self.loss_sum / max(1, self.steps)).cpu().numpy()
        return {"bce": float(v[0]), "recon": float(v[1]), "loss": float(v[2])}
