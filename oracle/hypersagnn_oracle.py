"""CPU ORACLE (TEST INFRASTRUCTURE — never imported by the product path).

A from-scratch functional restatement, in plain torch CPU ops, of the *live* arithmetic of the
reference Hyper-SAGNN hyperedge-scoring path (ma-compbio/MATCHA).  It takes nothing but a
``state_dict``-style mapping of tensors plus the feature tables, so it can run on the GPU box where
``/root/reference`` does not exist.  It is pinned against golden vectors produced by executing the
UNMODIFIED reference ``Code/Modules.py`` in the authoring container (``oracle/make_golden.py`` ->
``tests/golden/*.npz``; checked by ``tests/test_oracle_golden.py``).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import this file.

Reference lines restated (all under /root/reference/Code):
  encoder ............ Modules.py:176-188 (MultipleEmbedding.forward), :104-113 (TiedAutoEncoder.forward)
  recon loss ......... Modules.py:192-199, z-scored target built at :147-152
  attribute + mix .... Modules.py:261-270 (Classifier.get_embedding)
  attention .......... Modules.py:513-575 (MultiHeadAttention.forward), :448-460, :424-446
  pff_n1 ............. Modules.py:353-376 via EncoderLayer.forward :611-617
  scorer ............. Modules.py:278-318 (Classifier.forward)
  loss ............... main.py:56 (BCE-with-logits, weight=w), main.py:166 (alpha*bce + beta*recon)
  AdamW .............. main.py:630 (torch.optim.AdamW(lr=1e-3) defaults)

Dead compute of the reference (tied decoder :115-122, fc2 :573, pff_n2 :615, encode2, the never-applied
key-padding mask :281) is omitted: it has no effect on outputs or gradients.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

# ---------------------------------------------------------------------------------------------
# counter-based RNG shared (bit-exactly) with the CUDA kernels (matcha_b200/csrc/rng.cuh)
# ---------------------------------------------------------------------------------------------
_M64 = (1 << 64) - 1
GOLDEN = 0x9E3779B97F4A7C15
SITE_FEATURE, SITE_ATTN, SITE_PFF = 1, 2, 3
SITE_MULT = 0xD1B54A32D192ED03


def splitmix64(z: int) -> int:
    """One splitmix64 output for state z (python ints, exact)."""
    z = (z + GOLDEN) & _M64
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _M64
    return z ^ (z >> 31)


def splitmix64_np(z: np.ndarray) -> np.ndarray:
    z = z.astype(np.uint64)
    with np.errstate(over="ignore"):
        z = z + np.uint64(GOLDEN)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def site_key(seed: int, site: int) -> int:
    return splitmix64((seed ^ ((site * SITE_MULT) & _M64)) & _M64)


def dropout_threshold16(p: float) -> int:
    return int(round(p * 65536.0))


def dropout_keep_mask(seed: int, site: int, rows: np.ndarray, ncols: int, p: float) -> np.ndarray:
    """keep[i, j] for element (row id rows[i], column j): one splitmix64 word per 4 columns, 16 bits
    each; keep iff bits >= round(p * 2^16).  Same formula as csrc/rng.cuh::dropout_keep."""
    key = np.uint64(site_key(seed, site))
    rows = np.asarray(rows, dtype=np.uint64).reshape(-1, 1)
    j = np.arange(ncols, dtype=np.uint64).reshape(1, -1)
    with np.errstate(over="ignore"):
        ctr = key + (rows << np.uint64(24)) + (j >> np.uint64(2))
    r = splitmix64_np(ctr)
    v = (r >> ((j & np.uint64(3)) * np.uint64(16))) & np.uint64(0xFFFF)
    return v >= np.uint64(dropout_threshold16(p))


# ---------------------------------------------------------------------------------------------
# model description
# ---------------------------------------------------------------------------------------------
@dataclass
class OracleModel:
    """Everything the forward needs, as plain tensors.  ``params`` uses the reference's state_dict keys."""
    params: Dict[str, torch.Tensor]
    features: List[torch.Tensor]          # per chromosome [n_c, n_c] (SparseEmbedding.embedding, Modules.py:48-52)
    inter: Optional[torch.Tensor]         # [N, N] z-scored inter-chromosomal matrix (Modules.py:147-154) or None
    chrom_range: np.ndarray               # [C, 2] 1-based [start, end)  (process.py:21-36)
    n_head: int = 8
    p_feature: float = 0.2                # Modules.py:174
    p_attn: float = 0.3                   # Modules.py:226
    p_pff: float = 0.4                    # Modules.py:227
    extra: dict = field(default_factory=dict)

    @property
    def dtype(self):
        return self.params["next_w.FF_Linear0.weight"].dtype

    def to(self, dtype):
        return OracleModel({k: v.to(dtype) if v.is_floating_point() else v for k, v in self.params.items()},
                           [f.to(dtype) for f in self.features],
                           None if self.inter is None else self.inter.to(dtype),
                           self.chrom_range, self.n_head, self.p_feature, self.p_attn, self.p_pff, self.extra)

    def requires_grad_(self, names: Sequence[str]):
        for n in names:
            self.params[n] = self.params[n].detach().clone().requires_grad_(True)
        return self


def _ln(v, g, b, eps=1e-5):
    mu = v.mean(-1, keepdim=True)
    var = ((v - mu) ** 2).mean(-1, keepdim=True)
    return (v - mu) / torch.sqrt(var + eps) * g + b


def live_param_names(m: OracleModel) -> List[str]:
    """state_dict keys that receive a gradient."""
    C = len(m.chrom_range)
    names = []
    for c in range(C):
        names += [f"node_embedding.Embedding_Linear{c}.tied weight_0", f"node_embedding.Embedding_Linear{c}.tied weight_1",
                  f"node_embedding.Embedding_recon{c}.FF_Linear0.weight", f"node_embedding.Embedding_recon{c}.FF_Linear0.bias"]
    names += ["attribute_nn.weight", "attribute_nn.bias", "next_w.FF_Linear0.weight", "next_w.FF_Linear0.bias"]
    a = "encode1.mul_head_attn."
    names += [a + "layer_norm1.weight", a + "layer_norm1.bias", a + "layer_norm2.weight", a + "layer_norm2.bias",
              a + "layer_norm3.weight", a + "layer_norm3.bias", a + "w_qs.weight", a + "w_ks.weight", a + "w_vs.weight",
              a + "fc1.weight", a + "fc1.bias"]
    p = "encode1.pff_n1."
    names += [p + "PWF_Conv0.weight", p + "PWF_Conv0.bias", p + "PWF_Conv1.weight", p + "PWF_Conv1.bias",
              p + "layer_norm.weight", p + "layer_norm.bias"]
    names += ["layer_norm1.weight", "layer_norm1.bias", "layer_norm2.weight", "layer_norm2.bias",
              "pff_classifier.PWF_Conv0.weight", "pff_classifier.PWF_Conv0.bias"]
    return names


# ---------------------------------------------------------------------------------------------
# forward pieces
# ---------------------------------------------------------------------------------------------
def chrom_of(ids: torch.Tensor, chrom_range: np.ndarray) -> torch.Tensor:
    """chromosome index per id (-1 for pad id 0)."""
    out = torch.full_like(ids, -1)
    for c, (s, e) in enumerate(chrom_range):
        out[(ids >= int(s)) & (ids < int(e))] = c
    return out


def encoder(m: OracleModel, ids: torch.Tensor, train: bool = False, seed: int = 0,
            feature_masks: Optional[torch.Tensor] = None, row_base: int = 0):
    """Modules.py:176-188.  ids [T] int64 -> (E [T, d], H0 [T, d]).  In train mode the gathered feature
    row of token t is multiplied by keep/(1-p) with keep from the shared counter RNG (row id = row_base+t)."""
    P = m.params
    d = P["next_w.FF_Linear0.weight"].shape[0]
    T = ids.numel()
    E = torch.zeros(T, d, dtype=m.dtype)
    H0 = torch.zeros(T, d, dtype=m.dtype)
    ch = chrom_of(ids, m.chrom_range)
    for c, (s, e) in enumerate(m.chrom_range):
        sel = (ch == c).nonzero().flatten()
        if sel.numel() == 0:
            continue
        f = m.features[c][ids[sel] - int(s)]                       # Modules.py:184 (x - num_list[i] - 1)
        if train and m.p_feature > 0:
            keep = dropout_keep_mask(seed, SITE_FEATURE, (sel.numpy() + row_base), f.shape[1], m.p_feature)
            f = f * torch.from_numpy(keep).to(m.dtype) * (1.0 / (1.0 - m.p_feature))
        W0 = P[f"node_embedding.Embedding_Linear{c}.tied weight_0"]
        W1 = P[f"node_embedding.Embedding_Linear{c}.tied weight_1"]
        h0 = torch.tanh(f @ W0.t())                                 # Modules.py:111-113
        H0 = H0.index_put((sel,), h0)
        E = E.index_put((sel,), h0 @ W1.t())
    return E, H0


def recon_loss(m: OracleModel, ids: torch.Tensor, E: torch.Tensor, random_chrom: int):
    """Modules.py:192-199 with the chromosome draw made explicit."""
    P = m.params
    s, e = (int(v) for v in m.chrom_range[random_chrom])
    other = ((ids < s) | (ids >= e)) & (ids != 0)
    if m.inter is None or other.sum() == 0:
        return torch.zeros(1, dtype=m.dtype)
    sel = other.nonzero().flatten()
    target = m.inter[ids[sel] - 1][:, s - 1:e - 1]
    R = P[f"node_embedding.Embedding_recon{random_chrom}.FF_Linear0.weight"]
    rb = P[f"node_embedding.Embedding_recon{random_chrom}.FF_Linear0.bias"]
    pred = torch.tanh(E[sel]) @ R.t() + rb
    return ((target - pred) ** 2).mean(dim=-1).mean().reshape(1) * 100


def mix(m: OracleModel, ids: torch.Tensor, E: torch.Tensor):
    """Modules.py:263-270: X = tanh(next_w(E + attribute_nn(attr[id])))."""
    P = m.params
    attr = P["attribute_dict_embedding.weight"][ids]
    a = attr @ P["attribute_nn.weight"].t() + P["attribute_nn.bias"]
    return torch.tanh((E + a) @ P["next_w.FF_Linear0.weight"].t() + P["next_w.FF_Linear0.bias"])


def attention_block(m: OracleModel, x_ids: torch.Tensor, X: torch.Tensor, train=False, seed=0, row_base=0):
    """Modules.py:513-575 + :611-614 + :353-376.  X [B, L, d] -> dyn2 [B, L, d] (pff_n1 output, masked)."""
    P = m.params
    B, L, d = X.shape
    H = m.n_head
    a = "encode1.mul_head_attn."
    q = _ln(X, P[a + "layer_norm1.weight"], P[a + "layer_norm1.bias"]) @ P[a + "w_qs.weight"].t()
    k = _ln(X, P[a + "layer_norm2.weight"], P[a + "layer_norm2.bias"]) @ P[a + "w_ks.weight"].t()
    v = _ln(X, P[a + "layer_norm3.weight"], P[a + "layer_norm3.bias"]) @ P[a + "w_vs.weight"].t()
    dk = q.shape[-1] // H
    q = q.view(B, L, H, dk).permute(0, 2, 1, 3)
    k = k.view(B, L, H, dk).permute(0, 2, 1, 3)
    v = v.view(B, L, H, dk).permute(0, 2, 1, 3)
    s = q @ k.transpose(-1, -2) / math.sqrt(dk)                    # temperature = d_k ** 0.5 (:493)
    eye = torch.eye(L, dtype=torch.bool)
    s = s.masked_fill(eye, -1e32)                                  # diagonal only (:443-445); pads stay live keys
    A = torch.softmax(s, dim=-1)
    o = (A @ v).permute(0, 2, 1, 3).reshape(B, L, H * dk)
    dyn = o @ P[a + "fc1.weight"].t() + P[a + "fc1.bias"]
    npm = (x_ids != 0).to(X.dtype).unsqueeze(-1)
    rows = (np.arange(B * L) + row_base)
    if train and m.p_attn > 0:
        keep = dropout_keep_mask(seed, SITE_ATTN, rows, d, m.p_attn).reshape(B, L, d)
        dyn = dyn * torch.from_numpy(keep).to(X.dtype) * (1.0 / (1.0 - m.p_attn))
    u = dyn * npm
    p = "encode1.pff_n1."
    h1 = torch.tanh(u @ P[p + "PWF_Conv0.weight"][:, :, 0].t() + P[p + "PWF_Conv0.bias"])
    if train and m.p_pff > 0:
        keep = dropout_keep_mask(seed, SITE_PFF, rows, d, m.p_pff).reshape(B, L, d)
        h1 = h1 * torch.from_numpy(keep).to(X.dtype) * (1.0 / (1.0 - m.p_pff))
    h2 = h1 @ P[p + "PWF_Conv1.weight"][:, :, 0].t() + P[p + "PWF_Conv1.bias"] + u
    return _ln(h2, P[p + "layer_norm.weight"], P[p + "layer_norm.bias"]) * npm


def score(m: OracleModel, x_ids: torch.Tensor, dyn2: torch.Tensor, X: torch.Tensor):
    """Modules.py:290-311 -> logits [B, 1]."""
    P = m.params
    D = _ln(dyn2, P["layer_norm1.weight"], P["layer_norm1.bias"])
    S = _ln(X, P["layer_norm2.weight"], P["layer_norm2.bias"])
    w = P["pff_classifier.PWF_Conv0.weight"][0, :, 0]
    z = ((D - S) ** 2) @ w + P["pff_classifier.PWF_Conv0.bias"]
    npm = (x_ids != 0).to(X.dtype)
    return ((z * npm).sum(-1) / (npm.sum(-1) + 1e-15)).unsqueeze(-1)


def forward(m: OracleModel, x: torch.Tensor, random_chrom: Optional[int] = None, train: bool = False,
            seed: int = 0):
    """Classifier.forward (Modules.py:278-318).  x int64 [B, L], 0 = pad.  Returns (logits [B,1], recon [1])."""
    x = x.long()
    B, L = x.shape
    ids = x.reshape(-1)
    E, _ = encoder(m, ids, train=train, seed=seed)
    rl = recon_loss(m, ids, E, random_chrom) if random_chrom is not None else torch.zeros(1, dtype=m.dtype)
    X = mix(m, ids, E).view(B, L, -1)
    dyn2 = attention_block(m, x, X, train=train, seed=seed)
    return score(m, x, dyn2, X), rl


def node_embeddings(m: OracleModel, ids: torch.Tensor):
    """Classifier.get_node_embeddings in eval mode == rows of embeddings.npy (main.py:462-476)."""
    return encoder(m, ids.reshape(-1).long())[0]


def bce_with_logits(logits, y, w):
    """main.py:56: F.binary_cross_entropy_with_logits(pred, y, weight=w) (mean)."""
    return torch.nn.functional.binary_cross_entropy_with_logits(logits, y, weight=w)


def loss_and_grads(m: OracleModel, x, y, w, alpha, beta, random_chrom, train=False, seed=0):
    """One training-step objective (main.py:166) and its gradients wrt every live parameter."""
    names = live_param_names(m)
    m.requires_grad_(names)
    logits, rl = forward(m, x, random_chrom=random_chrom, train=train, seed=seed)
    bce = bce_with_logits(logits, y, w)
    loss = alpha * bce + beta * rl.sum()
    grads = torch.autograd.grad(loss, [m.params[n] for n in names], allow_unused=True)
    g = {n: (torch.zeros_like(m.params[n]) if gi is None else gi) for n, gi in zip(names, grads)}
    return {"logits": logits.detach(), "bce": bce.detach(), "recon": rl.detach(), "loss": loss.detach(), "grads": g}


def adamw_step(p, g, m1, m2, step, lr=1e-3, b1=0.9, b2=0.999, eps=1e-8, wd=0.01):
    """torch.optim.AdamW defaults (main.py:630); step is 1-based.  Returns (p, m1, m2)."""
    p = p * (1 - lr * wd)
    m1 = b1 * m1 + (1 - b1) * g
    m2 = b2 * m2 + (1 - b2) * g * g
    denom = (m2.sqrt() / math.sqrt(1 - b2 ** step)) + eps
    p = p - (lr / (1 - b1 ** step)) * m1 / denom
    return p, m1, m2


# ---------------------------------------------------------------------------------------------
# k = 2 closed form: used to cross-check the table-based pair scorer
# ---------------------------------------------------------------------------------------------
def pair_tables(m: OracleModel):
    """Per-node tables D[n], S[n] (n = 0..N) such that
    logit(i, j) = 0.5 * sum_c w_c[(D[j]-S[i])^2 + (D[i]-S[j])^2] + b   for a width-2 hyperedge (i, j)."""
    P = m.params
    N = int(m.chrom_range[-1][1]) - 1
    ids = torch.arange(0, N + 1)
    E, _ = encoder(m, ids)
    X = mix(m, ids, E)
    a = "encode1.mul_head_attn."
    v = _ln(X, P[a + "layer_norm3.weight"], P[a + "layer_norm3.bias"]) @ P[a + "w_vs.weight"].t()
    dyn = v @ P[a + "fc1.weight"].t() + P[a + "fc1.bias"]           # attention weight is exactly 1 on the other token
    p = "encode1.pff_n1."
    h1 = torch.tanh(dyn @ P[p + "PWF_Conv0.weight"][:, :, 0].t() + P[p + "PWF_Conv0.bias"])
    h2 = h1 @ P[p + "PWF_Conv1.weight"][:, :, 0].t() + P[p + "PWF_Conv1.bias"] + dyn
    dyn2 = _ln(h2, P[p + "layer_norm.weight"], P[p + "layer_norm.bias"])
    D = _ln(dyn2, P["layer_norm1.weight"], P["layer_norm1.bias"])
    S = _ln(X, P["layer_norm2.weight"], P["layer_norm2.bias"])
    return D, S


def pair_logits_closed_form(m: OracleModel, pairs: torch.Tensor):
    D, S = pair_tables(m)
    w = m.params["pff_classifier.PWF_Conv0.weight"][0, :, 0]
    b = m.params["pff_classifier.PWF_Conv0.bias"]
    i, j = pairs[:, 0], pairs[:, 1]
    return (0.5 * ((((D[j] - S[i]) ** 2) @ w) + (((D[i] - S[j]) ** 2) @ w)) + b).unsqueeze(-1)


# ---------------------------------------------------------------------------------------------
# fixture helpers
# ---------------------------------------------------------------------------------------------
def model_from_npz(z, dtype=torch.float32) -> OracleModel:
    params = {}
    for k in z.files:
        if k.startswith("p/"):
            t = torch.from_numpy(z[k])
            params[k[2:]] = t.to(dtype) if t.is_floating_point() else t
    cr = z["chrom_range"]
    feats = [torch.from_numpy(z[f"feat/{c}"]).to(dtype) for c in range(len(cr))]
    inter = torch.from_numpy(z["inter"]).to(dtype) if "inter" in z.files else None
    return OracleModel(params, feats, inter, cr)
