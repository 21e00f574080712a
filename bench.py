#!/usr/bin/env python
"""bench.py -- hyperedges/s trained (pos + neg, fwd + bwd + AdamW, negative sampling included) on the
synthetic whole-genome 1 Mb SPRITE-like workload (BASELINE.json configs[1]: ~3.1k bins, k = 2..5, d = 64).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One process per GPU (torchrun env for N > 1).  A step = one pass of the hot path over one batch:
GPU negative sampler -> Classifier forward (train mode, dropout on) -> weighted BCE + beta * recon ->
backward -> (NCCL all-reduce) -> AdamW.  `value` is measured with the positives already resident in HBM;
`e2e` repeats the measurement with HOST (pinned) positives copied in and the loss copied out every step.
`--impl reference` times the CPU oracle port of the reference path (the reference is pure Python and does
not travel to the GPU box) on this box's host cores, on the SAME configuration (same data set, same 4096 positives +
12288 negatives per step).  The line also carries `cfg3` (the same step on BASELINE configs[2]'s 30,344 bins, with
HBM rooflines of the node-encoder and reconstruction-head kernels) and `pair_scores` (configs[3]: all-pairs scoring,
kernel-only and end to end through table build + scoring + on-device denoise post-processing + D2H).
Prints ONE JSON line.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "hyperedges_per_s_trained"
UNIT = "hyperedges/s"
WORKLOADS = {
    # BASELINE.json configs[1] -- the single-GPU configuration the metric is quoted on (the default)
    "cfg2": "cfg2: synthetic SPRITE-like clusters, whole-genome 1 Mb (3067 bins, 23 chromosomes), k=2..5 mixed, embed_dim 64",
    # configs[0] (the reference's CPU-runnable case) and configs[2] (the 8-GPU training configuration; per-GPU work is
    # the same at any N, so `--workload cfg3 --gpus 1` measures one rank of it): parity cases, selectable here
    "cfg1": "cfg1: synthetic SPRITE-like clusters, chr1+chr2 at 1 Mb (494 bins), k=2..5 mixed, embed_dim 64",
    "cfg3": "cfg3: synthetic SPRITE-like clusters, whole-genome 100 kb (30344 bins, 23 chromosomes), k=2..5 mixed, embed_dim 64",
    "cfg5": "cfg5: synthetic SPRITE-like clusters, whole-genome 50 kb (60653 bins, 23 chromosomes), k=2..5 mixed, embed_dim 128",
}
POS_PER_STEP = 4096          # positives per GPU per step; x3 negatives -> 16384 hyperedges / GPU / step
NEG_NUM = 3
KMERS_PER_SIZE = 400_000


def load_peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return {"hbm_gbs": float(p["hbm_gbs"]), "bf16_tflops": float(p["bf16_tflops"]),
                "bf16_tflops_sustained": float(p.get("bf16_tflops_sustained", p["bf16_tflops"])), "src": "measured"}
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "src": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.proc, self.path = gpu_index, None, None

    def start(self):
        try:
            f = tempfile.NamedTemporaryFile("w", suffix=".csv", delete=False)
            self.path = f.name
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"], stdout=f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        try:
            rows = [r.strip().split(", ") for r in open(self.path) if r.strip()]
            sm = [float(r[1]) for r in rows if len(r) >= 9]
            if sm:
                out["sm_mhz"], out["sm_max_mhz"], out["samples"] = float(np.median(sm)), float(rows[0][2]), len(sm)
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            for i, n in enumerate(names):
                if any(len(r) >= 9 and r[5 + i].strip().lower() == "active" for r in rows):
                    out["reasons"].append(n)
            os.unlink(self.path)
        except Exception:
            pass
        return out


# ------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference path (sampler + fwd + bwd + AdamW), reference batch size
# ------------------------------------------------------------------------------------------
def cpu_reference_arm(ds, steps, warmup, seed=0, P=96):
    import torch
    from oracle import hypersagnn_oracle as O
    from oracle import sampler_oracle as SO
    torch.set_num_threads(os.cpu_count())
    # parameters come from our own module constructor (same shapes / init laws as main.py:609-623), moved to the CPU
    from matcha_b200.synthetic import build_model
    model = build_model(ds, seed=1)
    sd = {k: v.detach().cpu().clone().float() for k, v in model.state_dict().items()}
    feats = [e.embedding.detach().cpu().clone() for e in model.node_embedding.embeddings]
    inter = model.node_embedding.inter_initial.embedding.detach().cpu().clone() if ds["inter"] is not None else None
    del model
    om = O.OracleModel(sd, feats, inter, ds["chrom_range"])
    names = O.live_param_names(om)
    pset = SO.build_set(ds["dict"])
    rng = np.random.RandomState(seed)
    m1 = {n: torch.zeros_like(om.params[n]) for n in names}
    m2 = {n: torch.zeros_like(om.params[n]) for n in names}
    pos_all, w_all = ds["positives"], ds["pos_weight"]
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        idx = rng.randint(0, len(pos_all), P)
        pos, pw = pos_all[idx], w_all[idx]
        neg, valid, _ = SO.sample_negatives(pos, pset, ds["chrom_range"], NEG_NUM, 0, seed=seed, step=it)
        x = torch.from_numpy(np.concatenate([pos, neg]))
        y = torch.cat([torch.ones(P, 1), torch.zeros(P * NEG_NUM, 1)])
        w = torch.cat([torch.from_numpy(pw).view(-1, 1), torch.from_numpy(valid.astype(np.float32)).view(-1, 1)])
        out = O.loss_and_grads(om, x, y, w, 1.0, 0.001, random_chrom=int(rng.randint(0, len(ds["nums"]))), train=True,
                               seed=it)
        with torch.no_grad():
            for n in names:
                p, a, b = O.adamw_step(om.params[n].detach(), out["grads"][n], m1[n], m2[n], it + 1)
                om.params[n], m1[n], m2[n] = p, a, b
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    per_step = float(np.mean(times))
    return {"value": P * (1 + NEG_NUM) / per_step, "ms_per_step": per_step * 1e3, "cores": os.cpu_count(),
            "sample": f"{steps} steps of {P} positives + {P * NEG_NUM} negatives after {warmup} warm-up "
                      f"(sampler + forward + backward + AdamW of the oracle port, torch CPU ops on {os.cpu_count()} threads)"}


# ------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--kmers-per-size", type=int, default=KMERS_PER_SIZE)
    ap.add_argument("--pos-per-step", type=int, default=POS_PER_STEP)
    ap.add_argument("--cpu-baseline-steps", type=int, default=8)
    ap.add_argument("--no-cfg3", action="store_true", help="skip the configs[2] (30,344 bins) sub-measurement")
    ap.add_argument("--cfg5", action="store_true", help="also measure configs[4] (60,653 bins, embed_dim 128: general tcgen05 contractions); "
                                                        "needs ~60 GB of host memory per rank for the synthetic dense matrices")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--gemm-impl", type=int, default=-1, help="-1 library default, 0 SIMT, 1 tcgen05")
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-pairs", action="store_true", help="skip the all-pairs scorer measurement")
    args = ap.parse_args()
    warmup = max(3, args.warmup) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries exactly ONE JSON line: anything libraries print while we run (NCCL's version banner goes to fd 1)
    # is sent to stderr instead; the saved descriptor is restored for the final print
    sys.stdout.flush()
    stdout_fd = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        sys.stdout.flush()
        os.dup2(stdout_fd, 1)
        print(json.dumps(obj), flush=True)
    config = {"workload": WORKLOADS[args.workload], "hyperedges_per_gpu_per_step": args.pos_per_step * (1 + NEG_NUM),
              "positives_per_gpu_per_step": args.pos_per_step, "neg_num": NEG_NUM, "padded_width": 5,
              "kmers_per_size": args.kmers_per_size, "parallelism": f"dp{world}",
              "l2": "per-step activation stream (~1.4 GB) exceeds the 126 MB L2; no explicit flush"}

    from matcha_b200.synthetic import make_dataset
    if args.impl == "reference":
        if rank != 0:
            return 0
        # same configuration as our arm (same data set, same batch); every step is one batch of the workload
        ds = make_dataset(args.workload, kmers_per_size=args.kmers_per_size, seed=0)
        r = cpu_reference_arm(ds, max(1, args.steps), max(1, args.warmup), P=args.pos_per_step)
        line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
                "note": "CPU oracle port of the reference path (oracle/hypersagnn_oracle.py + oracle/sampler_oracle.py, pinned to the "
                        "unmodified Modules.py by tests/golden); the reference itself is Python that cannot travel to the GPU box",
                "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": r["sample"]},
                "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        emit(line)
        return 0

    import torch
    import torch.distributed as dist
    from matcha_b200 import _lib
    from matcha_b200.parallel import init_from_env
    from matcha_b200.sampler import KmerHashSet, NegativeSampler
    from matcha_b200.synthetic import build_model
    from matcha_b200.trainer import Trainer

    assert torch.cuda.is_available(), "bench.py needs a CUDA device"
    torch.cuda.set_device(local)
    init_from_env()
    lib = _lib.load()
    if args.gemm_impl >= 0:
        lib.matcha_set_gemm_impl(args.gemm_impl)
    peaks = load_peaks()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world > 1:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    def read_profile():
        n = lib.matcha_profile_labels()
        tms, calls, kern = (C.c_float * n)(), (C.c_int64 * n)(), (C.c_int64 * n)()
        _lib.check(lib.matcha_profile_read(tms, calls, kern, n), "profile_read")
        return {lib.matcha_profile_label_name(i).decode(): (tms[i], calls[i], kern[i]) for i in range(n) if calls[i]}

    def measure_training(ds, P, steps, warm, want_e2e, want_big, clocks=None):
        """One training configuration: device-timed steps (inputs resident in HBM), per-call-site profile, optionally the
        host-fed end-to-end loop and the 4x batch.  Every rank executes exactly the same sequence of steps."""
        model = build_model(ds, seed=1)
        hs = KmerHashSet(len(ds["dict"]), width=5).insert(ds["dict"])
        sampler = NegativeSampler(hs, ds["chrom_range"], min_dis=0, neg_num=NEG_NUM, seed=2 + rank)
        trainer = Trainer(model, sampler, alpha=1.0, beta=0.001, seed=3, world_size=world, rank=rank)
        g = np.random.RandomState(7)
        perm = g.permutation(len(ds["positives"]))
        pos_host = torch.from_numpy(ds["positives"][perm][rank::world].copy()).pin_memory()     # rows rank::world of a shuffled pool
        w_host = torch.from_numpy(ds["pos_weight"][perm][rank::world].copy()).pin_memory()
        pos_dev, w_dev = pos_host.cuda(), w_host.cuda()
        st = {"P": P, "nb": len(pos_host) // P}
        assert st["nb"] >= 1, "not enough positives for one step"

        def run(n_steps, host_io, start):
            P_, nb = st["P"], st["nb"]
            if host_io:
                # the package's own host-fed loop (matcha_b200/trainer.py): per step one H2D copy of the batch from pinned
                # memory (prefetched one step ahead on a copy stream) and one D2H read of the step's losses
                losses, _, _ = trainer.run_host_batches(pos_host, w_host, P_, n_steps, start)
                assert bool(torch.isfinite(losses).all())
                return
            for i in range(n_steps):
                b, b2 = (start + i) % nb, (start + i + 1) % nb
                if i + 1 < n_steps:      # the next batch's positives: its negatives are sampled under this step
                    trainer.step(pos_dev[b * P_:(b + 1) * P_], w_dev[b * P_:(b + 1) * P_], pos_dev[b2 * P_:(b2 + 1) * P_],
                                 w_dev[b2 * P_:(b2 + 1) * P_])
                else:
                    trainer.step(pos_dev[b * P_:(b + 1) * P_], w_dev[b * P_:(b + 1) * P_])

        def timed(n_steps, host_io, start, profile=False):
            barrier()
            if profile:
                lib.matcha_profile_enable(1)
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
            run(n_steps, host_io, start)
            ev1.record()
            barrier()
            ms = ev0.elapsed_time(ev1)
            prof = None
            if profile:
                lib.matcha_profile_enable(0)
                prof = read_profile()
            return max_over_ranks(ms), prof

        if clocks is not None:
            clocks.start()                           # started before the warm-up: nvidia-smi takes a while to produce samples
        run(warm, False, 0)
        # the extra untimed steps on both sides give nvidia-smi time to sample the clocks under the same load
        run(max(20, steps), False, 0)
        ms_dev, _ = timed(steps, False, warm)                        # the reported value: no per-call-site events
        run(max(20, steps), False, 0)
        torch.cuda.synchronize()
        clk = clocks.stop() if clocks is not None else None
        _, prof = timed(steps, False, warm, profile=True)            # same steps again with CUDA events per call site
        out = {"ms_dev": ms_dev, "prof": prof, "clk": clk, "P": P}
        if want_e2e:
            run(2, True, 0)
            out["ms_e2e"], _ = timed(steps, True, warm + steps)
        out["losses"] = trainer.mean_losses()
        if want_big and len(pos_host) // (4 * P) >= 2:
            # supplementary: the same step at 4x the batch (a part of a step does not depend on the batch)
            st["P"], st["nb"] = 4 * P, len(pos_host) // (4 * P)
            run(3, False, 0)
            k_big = max(5, steps // 3)
            ms_big, _ = timed(k_big, False, 3)
            out["big"] = {"hyperedges_per_gpu_per_step": 4 * P * (1 + NEG_NUM), "value": 4 * P * (1 + NEG_NUM) * world * k_big / (ms_big * 1e-3),
                          "unit": UNIT, "ms_per_step": ms_big / k_big, "steps": k_big}
        # HBM accounting of the node encoder / reconstruction head (4 * n_c bytes per real token + 512 B; n_c averaged over
        # the real tokens of the positives pool: negatives stay on their positive's chromosomes)
        node_nc = np.zeros(ds["N"] + 1, dtype=np.float64)
        for (cs, ce) in ds["chrom_range"]:
            node_nc[int(cs):int(ce)] = float(ce - cs)
        real = ds["positives"] != 0
        out["row_bytes"] = 4.0 * float(node_nc[ds["positives"][real]].mean())
        out["real_tokens"] = P * (1 + NEG_NUM) * float(real.sum()) / len(ds["positives"])
        del trainer, model, hs, sampler, pos_dev, w_dev
        torch.cuda.empty_cache()
        return out

    def hbm_kernels(ds, m):
        """achieved HBM GB/s of the encoder / reconstruction-head call sites of one measured configuration"""
        n_chrom = len(ds["nums"])
        alg = {"enc0_gather_gemm": m["real_tokens"] * (m["row_bytes"] + 512.0), "enc0_wgrad": m["real_tokens"] * (m["row_bytes"] + 512.0),
               # every eligible token (real, outside the uniformly drawn chromosome, Modules.py:192) reads its 4 * n_r byte target
               # row and its E row and adds a dtE row
               "recon_pred_gemm": m["real_tokens"] * (1.0 - 1.0 / n_chrom) * (4.0 * ds["N"] / n_chrom + 512.0)}
        res = {}
        for k, work in alg.items():
            if k in m["prof"]:
                per = m["prof"][k][0] / m["prof"][k][1]
                res[k] = {"GB/s": work / (per * 1e-3) / 1e9, "ms_per_launch": per, "frac_of_peak": work / (per * 1e-3) / 1e9 / peaks["hbm_gbs"],
                          "algorithmic_bytes_per_launch": work}
        return res

    def pair_scorer_bench(iters=5):
        """Second metric of BASELINE.json: pair-scores/s of the denoise all-pairs scorer (denoise_contact.py:67-88) on
        configs[3]'s shape -- chr1 at 10 kb, 24,897 bins, n(n+1)/2 = 3.1e8 pairs generated on the device -- sharded by
        contiguous pair range across ranks with no communication.  `value` times the scoring kernel alone on resident
        tables; `e2e` is the whole denoise path of one chromosome from a trained-shape model: per-node table build
        (matcha_pair_tables: the encoder over 24,897 rows of 100 KB) + table packing + scoring (sharded) + gather of the
        packed scores to rank 0 + on-device post-processing (denoise_contact.py:160-192) + D2H of the finished matrix."""
        from matcha_b200 import hyper_sagnn as M
        from matcha_b200.denoise import QuantileUniform, denoise_matrix
        from matcha_b200.scorer import PairScorer
        n, d = 24897, 64
        gen = torch.Generator(device="cuda").manual_seed(5)
        total = int(lib.matcha_pair_count(1, n + 1, 0))
        b, e = total * rank // world, total * (rank + 1) // world
        # a model of the cfg4 shape: ONE chromosome of 24,897 bins; feature values do not change the cost
        feats = (torch.randn(n, n, device="cuda", generator=gen) * (1.0 / np.sqrt(n))).cpu().numpy()
        cr = np.asarray([[1, n + 1]], dtype=np.int64)
        attr = np.concatenate([np.zeros((1, 2), np.float32), np.stack([np.ones(n, np.float32), np.arange(n, dtype=np.float32) / n], 1)], 0)
        torch.manual_seed(1)
        ne = M.MultipleEmbedding([feats], d, False, np.asarray([n]), cr, None)
        model = M.Classifier(n_head=8, d_model=d, d_k=d, d_v=d, node_embedding=ne, diag_mask=True, bottle_neck=d,
                             attribute_dict=attr).to(M.device)
        model.eval()
        del feats
        sc = PairScorer(model)
        out = torch.empty(e - b, dtype=torch.float32, device="cuda")
        origin = (torch.rand(n, n, device="cuda", generator=gen) < 0.02).float()
        host_my = torch.empty(n, n, dtype=torch.float32).pin_memory() if rank == 0 else None
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]

        def kernel_only():
            sc.score_range(1, n + 1, 0, b, e, sigmoid=True, out=out)

        def whole(record):
            if record: ev[0].record()
            sc.refresh()                                   # per-node tables from the model (encoder over every node)
            sc._pack(1, n + 1)                             # operand blocks of the tensor-core scorer
            if record: ev[1].record()
            sc.score_range(1, n + 1, 0, b, e, sigmoid=True, out=out)
            if record: ev[2].record()
            if world > 1:
                width = (total + world - 1) // world + 1
                send = torch.zeros(width, dtype=torch.float32, device="cuda")
                send[:e - b] = out
                bufs = [torch.empty_like(send) for _ in range(world)] if rank == 0 else None
                dist.gather(send, bufs, dst=0)
                if rank == 0:
                    full = torch.empty(total, dtype=torch.float32, device="cuda")
                    for r in range(world):
                        rb, re_ = total * r // world, total * (r + 1) // world
                        full[rb:re_] = bufs[r][:re_ - rb]
            else:
                full = out
            if record: ev[3].record()
            if rank == 0:
                my = denoise_matrix(full, origin, n, 0, QuantileUniform(1000, random_state=0))
                if record: ev[4].record()
                host_my.copy_(my, non_blocking=True)
            elif record:
                ev[4].record()
            if record: ev[5].record()

        for _ in range(3):
            kernel_only()
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(iters):
            kernel_only()
        ev1.record()
        barrier()
        ms = max_over_ranks(ev0.elapsed_time(ev1) / iters)
        whole(False)
        barrier()
        t0 = time.perf_counter()
        whole(True)
        barrier()
        wall_ms = (time.perf_counter() - t0) * 1e3
        stages = {"tables_ms": ev[0].elapsed_time(ev[1]), "score_ms": ev[1].elapsed_time(ev[2]), "gather_ms": ev[2].elapsed_time(ev[3]),
                  "postprocess_ms": ev[3].elapsed_time(ev[4]), "d2h_ms": ev[4].elapsed_time(ev[5])}
        e2e_ms = max_over_ranks(ev[0].elapsed_time(ev[5]))
        del out, origin, sc, model
        torch.cuda.empty_cache()
        return total, ms, e2e_ms, wall_ms, stages

    ds = make_dataset(args.workload, kmers_per_size=args.kmers_per_size, seed=0)      # same data on every rank
    P = args.pos_per_step
    T = P * (1 + NEG_NUM) * 5
    m = measure_training(ds, P, args.steps, warmup, True, True, ClockSampler(local) if rank == 0 else None)
    ms_dev, prof, clk, ms_e2e, losses, big = m["ms_dev"], m["prof"], m["clk"], m["ms_e2e"], m["losses"], m.get("big")
    # configs[2] (whole genome at 100 kb, 30,344 bins): the configuration BASELINE names for 8-GPU training, measured at
    # every N with the same per-GPU batch (weak scaling) -- the encoder and the reconstruction head stream real HBM here
    cfg3 = None
    if not args.no_cfg3 and args.workload == "cfg2":
        ds3 = make_dataset("cfg3", kmers_per_size=min(args.kmers_per_size, 100_000), seed=0)
        k3 = max(5, args.steps // 2)
        m3 = measure_training(ds3, P, k3, 3, False, False)
        cfg3 = {"workload": WORKLOADS["cfg3"], "value": P * (1 + NEG_NUM) * world * k3 / (m3["ms_dev"] * 1e-3), "unit": UNIT,
                "ms_per_step": m3["ms_dev"] / k3, "steps": k3, "n_gpus": world, "hyperedges_per_gpu_per_step": P * (1 + NEG_NUM),
                "kernel_ms_per_step": {k: round(v[0] / k3, 4) for k, v in sorted(m3["prof"].items(), key=lambda kv: -kv[1][0])},
                "hbm_kernels": hbm_kernels(ds3, m3), "losses": m3["losses"]}
        del ds3
    # configs[4] (whole genome at 50 kb, 60,653 bins, embed_dim 128): the contractions run on the general tcgen05 kernel
    # (csrc/gemm_tcg.cu), the row-wise kernels on their embed_dim-128 SIMT instantiations; opt-in because the synthetic dense N x N matrices need ~60 GB of host memory per rank
    cfg5 = None
    if args.cfg5:
        import psutil
        avail = psutil.virtual_memory().available
        if avail < 70e9 * max(1, world):
            cfg5 = {"skipped": "host memory: %.0f GB available, ~60 GB per rank needed" % (avail / 1e9)}
        else:
            ds5 = make_dataset("cfg5", kmers_per_size=min(args.kmers_per_size, 50_000), seed=0)
            k5 = max(3, args.steps // 4)
            m5 = measure_training(ds5, P, k5, 3, False, False)
            cfg5 = {"workload": WORKLOADS["cfg5"], "value": P * (1 + NEG_NUM) * world * k5 / (m5["ms_dev"] * 1e-3), "unit": UNIT,
                    "ms_per_step": m5["ms_dev"] / k5, "steps": k5, "n_gpus": world, "hyperedges_per_gpu_per_step": P * (1 + NEG_NUM),
                    "contractions": "general tcgen05 kernel (csrc/gemm_tcg.cu, bf16x3 split, fp32 accumulate); row-wise kernels fp32 SIMT",
                    "kernel_ms_per_step": {k: round(v[0] / k5, 4) for k, v in sorted(m5["prof"].items(), key=lambda kv: -kv[1][0])},
                    "losses": m5["losses"]}
            del ds5
    pair_total, pair_ms, pair_e2e_ms, pair_wall_ms, pair_stages = (0, 1.0, 1.0, 1.0, {}) if args.no_pairs else pair_scorer_bench()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0
    per_step = P * (1 + NEG_NUM) * world
    value = per_step * args.steps / (ms_dev * 1e-3)
    e2e = per_step * args.steps / (ms_e2e * 1e-3)
    # dominant call site and its roofline (algorithmic flops / bytes per launch)
    d, qkg = 64, 1536
    impl_env = os.environ.get("MATCHA_GEMM_IMPL", "1") if args.gemm_impl < 0 else str(args.gemm_impl)
    fused = os.environ.get("MATCHA_FUSED", "1") != "0" and impl_env != "0"
    if fused:
        # fused hyperedge-tile kernels: QKG never leaves the SM, so the bound is the tensor pipe.  ALGORITHMIC flops only:
        # forward = the QKG projection; backward = its data + weight gradients (the in-kernel recompute of QKG and the three
        # bf16 passes of the fp32-accurate split are implementation cost, not counted)
        alg = {"attn_fwd": ("tensor", 2.0 * T * qkg * d), "attn_bwd": ("tensor", 2.0 * 2.0 * T * qkg * d)}
    else:
        alg = {
            "qkg_gemm": ("tensor", 2.0 * T * qkg * d), "qkg_wgrad": ("tensor", 2.0 * T * qkg * d), "qkg_dgrad": ("tensor", 2.0 * T * qkg * d),
            "attn_fwd": ("hbm", T * (qkg + d) * 4.0), "attn_bwd": ("hbm", T * (2 * qkg + d) * 4.0),
        }
    enc = hbm_kernels(ds, m)
    for k, v in enc.items():
        alg[k] = ("hbm", v["algorithmic_bytes_per_launch"])
    top = max(prof.items(), key=lambda kv: kv[1][0])
    tot_ms = sum(v[0] for v in prof.values())
    name, (tms, calls, _) = top
    per_launch_ms = tms / calls
    if name in alg:
        bound, work = alg[name]
    else:
        bound, work = "hbm", None
    if work is None:
        roof = {"bound": bound, "achieved": None, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": None, "traffic": None}
    elif bound == "tensor":
        ach = work / (per_launch_ms * 1e-3) / 1e12
        roof = {"bound": "tensor", "achieved": ach, "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s",
                "frac": ach / peaks["bf16_tflops_sustained"], "traffic": None}
    else:
        ach = work / (per_launch_ms * 1e-3) / 1e9
        roof = {"bound": "hbm", "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": ach / peaks["hbm_gbs"],
                "traffic": None}
    impl_used = args.gemm_impl if args.gemm_impl >= 0 else int(os.environ.get("MATCHA_GEMM_IMPL", "1") or 1)
    # DRAM traffic of the same kernel from the committed `ncu --set full` capture (bytes per launch), if one exists
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "top_kernel_traffic.json")))
        ent = tr.get(name if not fused else "fused_" + name)
        if ent and ent.get("tokens") == T:
            roof["traffic"] = ent["dram_read_bytes"] + ent["dram_write_bytes"]
    except Exception:
        pass
    roof.update({"kernel": ("fused_" + name) if fused and name.startswith("attn_") else name, "ms_per_launch": per_launch_ms,
                 "share_of_step": tms / tot_ms, "peak_source": peaks["src"],
                 "contractions": "tcgen05 bf16x3 split, fp32 accumulate" if impl_used == 1 else "fp32 SIMT"})
    pair_rate = pair_total / (pair_ms * 1e-3)
    pair = {"value": pair_rate, "unit": "pair-scores/s", "ms_per_pass": pair_ms, "pairs": pair_total,
            "workload": "cfg4 shape: chr1 at 10 kb, 24,897 bins, all n(n+1)/2 pairs generated on the device, sigmoid applied, "
                        "sharded by pair range over the ranks",
            "roofline": {"bound": "hbm", "achieved": pair_rate * 4.0 / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                         "frac": pair_rate * 4.0 / 1e9 / peaks["hbm_gbs"] / world, "peak_source": peaks["src"],
                         "note": "4 bytes written per pair; tables (12.7 MB) stay in L2"},
            "e2e": {"value": pair_total / (pair_e2e_ms * 1e-3), "unit": "pair-scores/s", "ms": pair_e2e_ms, "wall_ms": pair_wall_ms,
                    "stages_ms_rank0": {k: round(v, 3) for k, v in pair_stages.items()}, "d2h_bytes": 24897 * 24897 * 4,
                    "what": "model -> per-node tables (encoder over 24,897 rows of 100 KB) -> packed operands -> all-pairs scores (sharded) "
                            "-> gather to rank 0 -> denoise post-processing on the device (denoise_contact.py:160-192, quantile map "
                            "included) -> finished [n, n] matrix copied to pinned host memory"}}
    # SURVEY 8f rank 1 (generate_kmers.py): k-mer enumeration + counting on synthetic clusters, the producer of this path's inputs
    kmers = None
    if not args.no_pairs:
        try:
            import importlib.util
            import types
            spec = importlib.util.spec_from_file_location("bench_kmers", os.path.join(ROOT, "scripts", "bench_kmers.py"))
            bk = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(bk)
            kmers = bk.measure(types.SimpleNamespace(clusters=500_000, nodes=30344, k=3, min_distance=0, min_freq=2, cpu_sample=5000))
        except Exception as exc:      # a secondary measurement must not take the headline line down
            kmers = {"error": repr(exc)}
    launches = int(sum(v[2] for v in prof.values()))
    breakdown = {k: round(v[0] / args.steps, 4) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0])}

    cpu = None
    if not args.no_cpu_baseline and world == 1:      # the CPU baseline is reported at N = 1 only: a bounded sample (~15 s)
        r = cpu_reference_arm(ds, args.cpu_baseline_steps, 2, P=args.pos_per_step)
        cpu = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": r["sample"]}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warmup,
            "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": config, "roofline": roof, "cpu_baseline": cpu,
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": int(P * 5 * 8 + P * 4), "d2h_bytes_per_step": 12,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches, "clocks": clk, "kernel_ms_per_step": breakdown, "losses": losses,
            "pair_scores": None if args.no_pairs else pair, "large_batch": big, "cfg3": cfg3, "cfg5": cfg5, "kmers": kmers,
            "bloom_agreement": "not measured: the reference's positive set is pybloom_live.BloomFilter (third party, version unpinned, "
                               "absent from this image); membership here is exact (bit-exact vs a CPU set, tests/test_gpu_parity.py)"}
    # achieved HBM GB/s of the node-encoder kernels (the north star's evidence for the sparse-row encoder)
    line["encoder_hbm"] = {k: {kk: vv for kk, vv in v.items() if kk != "algorithmic_bytes_per_launch"} for k, v in enc.items()
                           if k in ("enc0_gather_gemm", "enc0_wgrad")}
    emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
