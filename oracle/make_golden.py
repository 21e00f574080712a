"""Generate golden vectors by EXECUTING THE UNMODIFIED REFERENCE (authoring container only).

    python oracle/make_golden.py            # writes tests/golden/ref_small.npz and ref_small_d128.npz (embed_dim 128)

Imports /root/reference/Code/Modules.py untouched (only a stub `pybloom_live` package is put on
sys.path because the real one is not installed; it is needed solely by `utils.py:9`'s import).  The
reference has no tests or golden vectors of its own, so these files ARE the pin
for the oracle restatement (oracle/hypersagnn_oracle.py) and, through it, for the CUDA path.

The GPU box has no /root/reference: nothing under tests/, bench.py or smoke() runs this script.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/Code"
sys.path.insert(0, os.path.join(HERE, "_stubs"))
sys.path.insert(0, REF)

import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

import Modules  # noqa: E402  (the reference, unmodified)


def build_reference(nums, d=64, seed=1):
    rng = np.random.default_rng(seed)
    C = len(nums)
    starts = np.concatenate([[0], np.cumsum(nums)])
    chrom_range = np.stack([starts[:-1] + 1, starts[1:] + 1], axis=1).astype(np.int64)   # process.py:21-36
    num_list = np.cumsum(nums)
    N = int(num_list[-1])
    # intra features exactly as main.py:572-577: corrcoef of the intra adjacency block, NaN -> 0
    feats = []
    for n in nums:
        adj = rng.poisson(2.0, size=(n, n)).astype("float32")
        adj = adj + adj.T
        adj[rng.integers(0, n)] = 0.0                       # an empty row -> NaN corrcoef row -> 0 (main.py:575)
        with np.errstate(invalid="ignore", divide="ignore"):
            t = np.corrcoef(adj).astype("float32")
        t[np.isnan(t)] = 0.0
        feats.append(t)
    inter = (rng.poisson(0.6, size=(N, N)) * (rng.random((N, N)) < 0.5)).astype("float32")
    for c in range(C):                                       # inter-chromosomal only
        s, e = chrom_range[c] - 1
        inter[s:e, s:e] = 0.0
    # attribute table as main.py:497-512
    attrs = []
    for i, n in enumerate(nums):
        ch = np.zeros((n, C)); ch[:, i] = 1
        coor = np.arange(n).reshape(-1, 1).astype("float32") / nums[0]
        attrs.append(np.concatenate([ch, coor], -1))
    attr = np.concatenate([np.zeros((1, C + 1)), np.concatenate(attrs, 0)], 0).astype("float32")

    build_reference.inter_raw = inter.copy()                 # kept to pin our own z-score restatement
    torch.manual_seed(seed)
    ne = Modules.MultipleEmbedding(feats, d, False, torch.as_tensor(num_list), chrom_range, inter.copy())
    model = Modules.Classifier(n_head=8, d_model=d, d_k=d, d_v=d, node_embedding=ne, diag_mask=True,
                               bottle_neck=d, attribute_dict=attr)
    # perturb the parameters that torch initialises to exactly 0/1 so the golden vectors exercise them
    g = torch.Generator().manual_seed(seed + 7)
    with torch.no_grad():
        for k, p in model.named_parameters():
            if "layer_norm" in k or k.endswith("bias"):
                p.add_(0.1 * torch.randn(p.shape, generator=g))
    return model, feats, chrom_range, N


def make_inputs(rng, N, L, B):
    x = np.zeros((B, L), dtype=np.int64)
    for b in range(B):
        k = L if b < B // 2 else int(rng.integers(2, L + 1))   # half full-width, half padded
        x[b, :k] = np.sort(rng.choice(np.arange(1, N + 1), size=k, replace=False))
    return x


def main(out=os.path.join(HERE, "..", "tests", "golden", "ref_small.npz"), d=64, widths=(2, 3, 4, 5), full_grads=(3, 5), full_grad_max_elems=None):
    nums = [40, 28, 33]
    model, feats, chrom_range, N = build_reference(nums, d=d)
    C = len(nums)
    sd = model.state_dict()
    arrays = {"chrom_range": chrom_range, "nums": np.asarray(nums)}
    for c, f in enumerate(model.node_embedding.embeddings):
        arrays[f"feat/{c}"] = f.embedding.numpy()
    arrays["inter"] = model.node_embedding.inter_initial.embedding.numpy()      # z-scored by Modules.py:147-152
    arrays["inter_raw"] = build_reference.inter_raw
    grads_any = {}

    rng = np.random.default_rng(5)
    real_choice = np.random.choice
    try:
        for L in widths:
            x = make_inputs(rng, N, L, 16)
            xt = torch.from_numpy(x)
            arrays[f"x/L{L}"] = x
            model.eval()
            with torch.no_grad():
                arrays[f"logits_eval/L{L}"] = model(xt).numpy()
                for r in range(C):
                    np.random.choice = lambda a, size=None, r=r: np.asarray([r])   # pin Modules.py:192's draw
                    out_l, rl = model(xt, return_recon=True)
                    arrays[f"recon_eval/L{L}/r{r}"] = rl.numpy()
            # gradients: train mode with every dropout probability forced to 0 (torch's dropout stream
            # cannot be reproduced elsewhere); loss as main.py:166 with alpha = 1, beta = 0.5
            model.train()
            for mod in model.modules():
                if isinstance(mod, torch.nn.Dropout):
                    mod.p = 0.0
            model.node_embedding.dropout.p = 0.0
            r = L % C
            np.random.choice = lambda a, size=None, r=r: np.asarray([r])
            y = torch.from_numpy((rng.random((16, 1)) < 0.4).astype("float32"))
            w = torch.from_numpy(rng.uniform(0.5, 3.0, size=(16, 1)).astype("float32"))
            model.zero_grad(set_to_none=True)
            pred, rl = model(xt, return_recon=True)
            bce = F.binary_cross_entropy_with_logits(pred, y, weight=w)
            loss = bce * 1.0 + rl * 0.5
            loss.backward()
            arrays[f"y/L{L}"], arrays[f"w/L{L}"] = y.numpy(), w.numpy()
            arrays[f"train_logits/L{L}"] = pred.detach().numpy()
            arrays[f"train_bce/L{L}"] = bce.detach().numpy().reshape(1)
            arrays[f"train_recon/L{L}"] = rl.detach().numpy()
            arrays[f"train_rchrom/L{L}"] = np.asarray([r])
            live = []
            for k, p in model.named_parameters():
                if p.grad is not None:
                    if L in full_grads and (full_grad_max_elems is None or p.numel() <= full_grad_max_elems):
                        arrays[f"grad/L{L}/{k}"] = p.grad.numpy().copy()     # full gradients for some widths / tensors keep the fixture small
                    else:
                        arrays[f"gradnorm/L{L}/{k}"] = np.asarray([p.grad.double().norm().item()])
                    live.append(k)
            grads_any[L] = live
    finally:
        np.random.choice = real_choice

    model.eval()
    with torch.no_grad():
        ids = torch.arange(1, N + 1).view(-1, 1)
        arrays["embeddings"] = model.get_node_embeddings(ids).numpy()[:, 0, :]          # main.py:462-476
        # known-answer facts of SURVEY.md section 3.4
        t = torch.tensor([[3, 47, 90]])
        arrays["kat/pad_width"] = np.asarray([model(t).item(),
                                              model(torch.tensor([[3, 47, 90, 0]])).item(),
                                              model(torch.tensor([[3, 47, 90, 0, 0]])).item()], dtype="float32")
        arrays["kat/permuted"] = np.asarray([model(torch.tensor([[90, 3, 47]])).item()], dtype="float32")
        pairs = np.stack(np.triu_indices(N, 0), 1)[::37] + 1
        arrays["kat/pairs"] = pairs.astype(np.int64)
        arrays["kat/pair_logits"] = model(torch.from_numpy(pairs)).numpy()

    dead = {}
    live_all = set(sum(grads_any.values(), []))
    for k, v in sd.items():
        if k in live_all or k.startswith("attribute_dict") or "Embedding_recon" in k:   # recon heads feed the eval recon loss
            arrays[f"p/{k}"] = v.numpy()             # values only for tensors that influence outputs
        else:
            dead[k] = list(v.shape)
    arrays["meta"] = np.asarray(json.dumps({
        "state_dict_keys": {k: list(v.shape) for k, v in sd.items()},
        "no_grad_keys": sorted(dead),
        "live_keys_by_L": grads_any,
        "torch": torch.__version__, "numpy": np.__version__,
        "reference": "ma-compbio/MATCHA Code/Modules.py (unmodified)",
    }))
    np.savez_compressed(out, **arrays)
    print("wrote", out, "%.1f KB" % (os.path.getsize(out) / 1024), "keys", len(arrays))


if __name__ == "__main__":
    main()
    # embed_dim 128 (BASELINE.json configs[4]): a smaller fixture, widths 3 and 5, full gradients of the tensors up to 20 000 elements (norms for the rest) at width 5
    main(out=os.path.join(HERE, "..", "tests", "golden", "ref_small_d128.npz"), d=128, widths=(3, 5), full_grads=(5,), full_grad_max_elems=20000)
