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
        # the derived-weight folds of matcha_prepare depend on the weights only: they run on a side stream under the
        # batch assembly + negative sampling of the same step (both are small launches that leave most SMs idle)
        self._side = torch.cuda.Stream(device=self.e.dev)
        self._copy = None

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
        main = torch.cuda.current_stream()
        self._side.wait_stream(main)               # the previous step's AdamW (weights) and backward (derived buffers)
        with torch.cuda.stream(self._side):
            e.prepare()
        self.x[:P].copy_(pos)
        self.w[:P].copy_(pos_w)
        self.sampler.sample(pos, out=self.x[P:], valid=self.valid)
        self.w[P:].copy_(self.valid)               # exhausted rows (emitted as the positive itself) get weight 0
        rchrom = int(self.recon_rng.randint(0, e.C)) if (self.beta != 0.0 and e.desc.inter) else -1
        seed = e.next_seed()
        main.wait_stream(self._side)
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

    def run_host_batches(self, pos_host, w_host, P, n_steps, start=0):
        """Steps over positives held in PINNED HOST memory (int64 [n, L], fp32 [n]): every step copies its own batch
        host -> device (issued one step ahead on a copy stream, two device slots) and reads its {bce, recon, loss}
        back device -> host (asynchronously into a pinned [n_steps, 3] array, complete when this returns), so the host
        never blocks on the step it has just launched.  Batch i is rows [b * P, (b + 1) * P), b = (start + i) mod
        (n // P).  Returns (losses fp32 [n_steps, 3] pinned, h2d bytes per step, d2h bytes per step)."""
        dev = self.e.dev
        main = torch.cuda.current_stream()
        if self._copy is None:
            self._copy = torch.cuda.Stream(device=dev)
        L = pos_host.shape[1]
        nb = len(pos_host) // P
        xs = [torch.empty(P, L, dtype=torch.int64, device=dev) for _ in range(2)]
        ws = [torch.empty(P, dtype=torch.float32, device=dev) for _ in range(2)]
        ready = [torch.cuda.Event() for _ in range(2)]
        consumed = [torch.cuda.Event() for _ in range(2)]
        out = torch.empty(n_steps, 3, dtype=torch.float32).pin_memory()

        def issue(i):
            slot, b = i & 1, (start + i) % nb
            self._copy.wait_event(consumed[slot])          # the step that last used this slot has finished (no-op at first)
            with torch.cuda.stream(self._copy):
                xs[slot].copy_(pos_host[b * P:(b + 1) * P], non_blocking=True)
                ws[slot].copy_(w_host[b * P:(b + 1) * P], non_blocking=True)
                ready[slot].record(self._copy)

        self._copy.wait_stream(main)
        issue(0)
        for i in range(n_steps):
            if i + 1 < n_steps:
                issue(i + 1)
            slot = i & 1
            main.wait_event(ready[slot])
            self.step(xs[slot], ws[slot])
            consumed[slot].record(main)
            out[i].copy_(self.loss_out, non_blocking=True)
        main.synchronize()
        return out, P * L * 8 + P * 4, 12

    def mean_losses(self):
        """Host read of the running means (one sync; call once per epoch, as main.py:197 does)."""
        v = (self.loss_sum / max(1, self.steps)).cpu().numpy()
        return {"bce": float(v[0]), "recon": float(v[1]), "loss": float(v[2])}
