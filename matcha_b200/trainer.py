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
                 seed=0, world_size=1, rank=0, recon_rng=None, fused_boundary=True):
        """fused_boundary=False keeps the step boundary as separate launches (all-reduce, AdamW) and leaves the step's
        gradients in the flat buffer afterwards (tests that inspect them); the default clears the buffer inside the fused
        boundary kernel."""
        model.train()
        self.model = model
        self.e = model._engine()
        self.e.ensure_bound()
        # dropout streams differ per rank; the recon-chromosome stream (below) is shared by all ranks
        self.e.seed_base = (int(seed) + 0x632BE59BD9B4E019 * int(rank)) & ((1 << 64) - 1)
        self.sampler = sampler
        self.alpha, self.beta = float(alpha), float(beta)
        if self.beta != 0.0 and self.e.C > 1:
            self.e.require_inter()            # beta * recon with no target would silently train on 0 (main.py phase 1!)
        self.opt = FlatAdamW(self.e, lr=lr, weight_decay=weight_decay)
        self.world = int(world_size)
        self.rank = int(rank)
        self._peer = None
        if self.world == 1 and fused_boundary and self.e.n_flat % 4 == 0 and self.e.n_always % 4 == 0:
            # single replica: the same kernel without barriers = AdamW + gradient-buffer clear in ONE launch
            import ctypes as C
            self.active_red = torch.zeros_like(self.e.active)
            self._peer = {"g": (C.c_void_p * 1)(self.e.gflat.data_ptr()), "a": (C.c_void_p * 1)(self.e.active.data_ptr()), "b": None,
                          "epoch": 0}
        if self.world > 1:
            # every replica must start from the SAME weights: rank 0's parameters win (a launcher that forgot to seed the
            # constructors identically would otherwise all-reduce gradients taken at different points forever)
            import torch.distributed as dist
            dist.broadcast(self.e.flat, src=0)
            if fused_boundary:
                self._setup_peer_step()
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
        self._side2 = torch.cuda.Stream(device=self.e.dev)     # assembles the next step's batch (see step())
        self._copy = None

    def _setup_peer_step(self):
        """Map every rank's gradient buffer / activity flags / barrier flags into this process so the step boundary runs as
        ONE kernel over NVLink peer memory (csrc/dp_fused.cu).  Falls back to the NCCL all-reduce + AdamW launches when the
        GPUs cannot map each other (MATCHA_DP_FUSED=0 forces that path).  The decision is made collectively."""
        import ctypes as C
        import os
        import torch.distributed as dist
        from .parallel import PeerBuffers, peer_access_available
        e = self.e
        ok = os.environ.get("MATCHA_DP_FUSED", "1") != "0" and dist.get_backend() == "nccl" and self.world <= 8 \
            and peer_access_available(self.world) and e.n_flat % 4 == 0 and e.n_always % 4 == 0
        flag = torch.tensor([1 if ok else 0], device=e.dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) == 0:
            return
        bar = torch.zeros(int(e.lib.matcha_dp_barrier_bytes()) // 4, dtype=torch.int32, device=e.dev)
        self.active_red = torch.zeros_like(e.active)
        torch.cuda.synchronize()
        dist.barrier()
        peers = {"g": PeerBuffers(e.gflat, self.rank, self.world), "a": PeerBuffers(e.active, self.rank, self.world),
                 "b": PeerBuffers(bar, self.rank, self.world)}
        arr = lambda pb: (C.c_void_p * self.world)(*pb.ptrs)
        self._peer = {"bufs": peers, "g": arr(peers["g"]), "a": arr(peers["a"]), "b": arr(peers["b"]), "epoch": 0}
        dist.barrier()

    def _peer_step(self):
        """all-reduce + mean + AdamW + gradient-buffer clear in one launch (every rank calls it once per step)."""
        e, o, pr = self.e, self.opt, self._peer
        pr["epoch"] += 1
        o.t += 1
        if self.world == 1 and pr["g"][0] != e.gflat.data_ptr():      # the engine re-bound its buffers (model.to(...))
            pr["g"][0], pr["a"][0] = e.gflat.data_ptr(), e.active.data_ptr()
        check(e.lib.matcha_dp_reduce_adamw(self.world, self.rank, pr["g"], pr["a"], pr["b"], ptr(e.flat), ptr(o.m1), ptr(o.m2),
                                           ptr(self.active_red), e.n_always, e.n_flat, len(e.segments), e.active.numel(),
                                           ptr(o.seg_begin), ptr(o.seg_end), ptr(o.seg_flag), ptr(o.seg_step), o.t, pr["epoch"],
                                           o.lr, o.betas[0], o.betas[1], o.eps, o.wd, stream_ptr()), "matcha_dp_reduce_adamw")

    def _buffers(self, P, L):
        if self._buf_P == (P, L):
            return
        dev, n = self.e.dev, P * (1 + self.neg_num)
        # two batch slots: while step i computes on one, the batch of step i + 1 is assembled in the other
        self.xb = [torch.zeros(n, L, dtype=torch.int64, device=dev) for _ in range(2)]
        self.wb = [torch.ones(n, dtype=torch.float32, device=dev) for _ in range(2)]
        self.vb = [torch.empty(n - P, dtype=torch.uint8, device=dev) for _ in range(2)]
        self.y = torch.zeros(n, dtype=torch.float32, device=dev)
        self.y[:P] = 1.0
        self.logits = torch.empty(n, dtype=torch.float32, device=dev)
        self.dlogit = torch.empty(n, dtype=torch.float32, device=dev)
        self.recon = torch.zeros(1, dtype=torch.float32, device=dev)
        self._assembled = [torch.cuda.Event() for _ in range(2)]
        self._released = [torch.cuda.Event() for _ in range(2)]
        self._slot, self._pending = 0, None
        self._buf_P = (P, L)

    def _assemble(self, slot, pos, pos_w):
        """x = [pos; negatives], w = [pos_w; valid] into batch slot `slot`, on the current stream."""
        P = pos.shape[0]
        x, w, valid = self.xb[slot], self.wb[slot], self.vb[slot]
        x[:P].copy_(pos)
        w[:P].copy_(pos_w)
        self.sampler.sample(pos, out=x[P:], valid=valid)
        w[P:].copy_(valid)                         # exhausted rows (emitted as the positive itself) get weight 0

    def step(self, pos, pos_w, next_pos=None, next_w=None, next_ready=None):
        """pos int64 [P, L] (device, rows sorted, zero padded), pos_w fp32 [P] (device).  Returns nothing:
        running losses accumulate on the device in `loss_sum` ({bce, recon, total}).

        next_pos / next_w (optional): the positives of the FOLLOWING step.  Their negatives are sampled and the batch is
        assembled on a side stream while this step computes (the sampler depends on the positives and the hash set only),
        so the next call -- which must pass exactly these tensors as pos / pos_w -- finds its batch ready.  next_ready: a
        CUDA event after which next_pos / next_w hold their data (e.g. the H2D copy of a host-fed loop)."""
        e, lib = self.e, self.e.lib
        P, L = pos.shape
        self._buffers(P, L)
        n = P * (1 + self.neg_num)
        main = torch.cuda.current_stream()
        self._side.wait_stream(main)               # the previous step's AdamW (weights) and backward (derived buffers)
        with torch.cuda.stream(self._side):
            e.prepare()
        slot = self._slot
        hit = False
        if self._pending is not None:
            main.wait_event(self._assembled[slot])  # assembled under the previous step
            hit = self._pending == (pos.data_ptr(), pos_w.data_ptr(), P, L)
            self._pending = None
        if not hit:
            self._assemble(slot, pos, pos_w)
        x, w = self.xb[slot], self.wb[slot]
        rchrom = int(self.recon_rng.randint(0, e.C)) if (self.beta != 0.0 and e.desc.inter) else -1
        seed = e.next_seed()
        # no main.wait_stream(side) here: matcha_forward itself waits for the prepare it depends on, after the token
        # bucketing (which needs no weights) has been queued, so bucketing and prepare overlap too
        e.run_forward(x, True, seed, rchrom, logits=self.logits, recon=self.recon)
        if next_pos is not None and tuple(next_pos.shape) == (P, L):
            other = slot ^ 1
            self._side2.wait_event(self._released[other])      # the step that last read that slot (no-op at first)
            if next_ready is not None:
                self._side2.wait_event(next_ready)
            else:
                self._side2.wait_stream(main)
            with torch.cuda.stream(self._side2):
                self._assemble(other, next_pos, next_w)
                self._assembled[other].record(self._side2)
            self._pending = (next_pos.data_ptr(), next_w.data_ptr(), P, L)
        check(lib.matcha_bce_loss(ptr(self.logits), ptr(self.y), ptr(w), n, self.alpha, self.beta, ptr(self.recon),
                                  ptr(self.dlogit), ptr(self.loss_out), stream_ptr()), "matcha_bce_loss")
        if self._peer is None or self._peer["epoch"] == 0:
            e.gflat.zero_()                     # (the fused data-parallel step clears the buffer itself afterwards)
        e.run_backward(x, seed, rchrom, self.dlogit, self.beta)
        self._released[slot].record(main)
        if self._peer is not None:
            self._peer_step()                   # barrier + P2P all-reduce + mean + AdamW + clear: one launch over NVLink
        else:
            # one collective per step: gradients + activity flags (no-op when world == 1)
            scale = allreduce_grads_and_flags(e.gflat, e.n_flat, e.active, self.world)
            self.opt.step(grad_scale=scale)
        self.loss_sum += self.loss_out
        self.steps += 1
        self._slot = slot ^ 1 if self._pending is not None else slot
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
            if i + 1 < n_steps:
                self.step(xs[slot], ws[slot], xs[slot ^ 1], ws[slot ^ 1], ready[slot ^ 1])
            else:
                self.step(xs[slot], ws[slot])
            consumed[slot].record(main)
            out[i].copy_(self.loss_out, non_blocking=True)
        main.synchronize()
        return out, P * L * 8 + P * 4, 12

    def mean_losses(self):
        """Host read of the running means (one sync; call once per epoch, as main.py:197 does)."""
        v = (self.loss_sum / max(1, self.steps)).cpu().numpy()
        return {"bce": float(v[0]), "recon": float(v[1]), "loss": float(v[2])}
