"""Host-side mirror of the reference operator surface (ma-compbio/MATCHA `Code/Modules.py`).

Same class names, constructor signatures, sub-module names and therefore the same ``state_dict`` keys
(including the parameters the reference never trains), so that ``main.py``-style drivers,
``torch.save(model)`` / ``torch.load`` of ``model2load`` and ``load_state_dict`` of a reference
``model.chkpt`` work unchanged.  What differs is what runs: ``Classifier.forward`` /
``get_node_embeddings`` hand the whole computation to the sm_100a kernels behind the C ABI
(``include/matcha_b200.h``) through ``matcha_b200.engine.Engine``.  There is no PyTorch or CPU
implementation of the math in this file and no fallback: without a CUDA device and the built
``libmatcha_b200.so`` the forward raises.

Reference lines mirrored: Modules.py:38-67 (SparseEmbedding), :70-102 (TiedAutoEncoder registration and
init), :125-174 (MultipleEmbedding construction incl. the row-wise z-score of the inter matrix :147-152),
:204-249 (Classifier construction), :252-318 (call signatures and return shapes), :327-352, :385-400,
:463-508, :578-607 (sub-module inventories), :620-681 (DataGenerator).
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn as nn

device = torch.device("cuda" if torch.cuda.is_available() else "cpu")
activation = torch.tanh

_FUSED = ("this sub-module only owns parameters: its arithmetic is fused into the CUDA hyperedge pipeline "
          "(call the enclosing Classifier instead)")


def get_non_pad_mask(seq):
    assert seq.dim() == 2
    return seq.ne(0).type(torch.float).unsqueeze(-1)


class SparseEmbedding(nn.Module):
    """Feature-row table of one chromosome (Modules.py:38-67).  Dense rows live on the device; a scipy
    CSR matrix is kept as CSR (``sparse=True``) and uploaded as CSR by the engine."""

    def __init__(self, embedding_weight, sparse=False):
        super().__init__()
        self.sparse = sparse
        if sparse and hasattr(embedding_weight, "tocsr"):
            self.embedding = embedding_weight.tocsr()
        else:
            self.sparse = False
            dense = embedding_weight.todense() if hasattr(embedding_weight, "todense") else embedding_weight
            self.embedding = torch.from_numpy(np.ascontiguousarray(np.asarray(dense), dtype=np.float32)).to(device)

    @property
    def width(self):
        return int(self.embedding.shape[-1])

    def forward(self, x):  # row lookup used by callers that want raw feature rows
        if self.sparse:
            rows = np.asarray(self.embedding[x.detach().cpu().numpy().reshape(-1), :].todense(), dtype=np.float32)
            return torch.from_numpy(rows).to(device)
        return self.embedding[x, :]


class TiedAutoEncoder(nn.Module):
    """Parameter holder with the reference's registration quirks (Modules.py:83-86: every layer's biases
    are registered under the same two names, so only the last pair survives; names contain a space)."""

    def __init__(self, shape_list, use_bias=True):
        super().__init__()
        self.use_bias = use_bias
        self.weight_list, self.bias_list, self.recon_bias_list = [], [], []
        for i in range(len(shape_list) - 1):
            self.weight_list.append(nn.Parameter(torch.empty(shape_list[i + 1], shape_list[i], device=device)))
            self.bias_list.append(nn.Parameter(torch.empty(shape_list[i + 1], device=device)))
            self.recon_bias_list.append(nn.Parameter(torch.empty(shape_list[i], device=device)))
        self.recon_bias_list = self.recon_bias_list[::-1]
        for i, w in enumerate(self.weight_list):
            self.register_parameter("tied weight_%d" % i, w)
            self.register_parameter("tied bias1", self.bias_list[i])
            self.register_parameter("tied bias2", self.recon_bias_list[i])
        self.reset_parameters()

    def reset_parameters(self):   # same distributions as Modules.py:90-102
        for w in self.weight_list:
            nn.init.kaiming_uniform_(w, a=math.sqrt(5))
        for w, b in zip(self.weight_list, self.bias_list):
            fan_in, _ = nn.init._calculate_fan_in_and_fan_out(w)
            nn.init.uniform_(b, -1 / math.sqrt(fan_in), 1 / math.sqrt(fan_in))
        for w, b in zip(self.weight_list[::-1], self.recon_bias_list):
            _, fan_out = nn.init._calculate_fan_in_and_fan_out(w)
            nn.init.uniform_(b, -1 / math.sqrt(fan_out), 1 / math.sqrt(fan_out))

    def forward(self, *a, **k):
        raise RuntimeError("TiedAutoEncoder: " + _FUSED)


class FeedForward(nn.Module):
    def __init__(self, dims, dropout=None, reshape=False, use_bias=True):
        super().__init__()
        self.w_stack = []
        for i in range(len(dims) - 1):
            self.w_stack.append(nn.Linear(dims[i], dims[i + 1], use_bias))
            self.add_module("FF_Linear%d" % i, self.w_stack[-1])
        self.dropout = nn.Dropout(dropout) if dropout is not None else None
        self.reshape = reshape

    def forward(self, *a, **k):
        raise RuntimeError("FeedForward: " + _FUSED)


class PositionwiseFeedForward(nn.Module):
    def __init__(self, dims, dropout=None, reshape=False, use_bias=True, residual=False, layer_norm=False):
        super().__init__()
        self.w_stack, self.dims = [], dims
        for i in range(len(dims) - 1):
            self.w_stack.append(nn.Conv1d(dims[i], dims[i + 1], 1, bias=use_bias))
            self.add_module("PWF_Conv%d" % i, self.w_stack[-1])
        self.reshape = reshape
        self.layer_norm = nn.LayerNorm(dims[-1])
        self.dropout = nn.Dropout(dropout) if dropout is not None else None
        self.residual = residual
        self.layer_norm_flag = layer_norm

    def forward(self, *a, **k):
        raise RuntimeError("PositionwiseFeedForward: " + _FUSED)


class MultiHeadAttention(nn.Module):
    def __init__(self, n_head, d_model, d_k, d_v, dropout, diag_mask, input_dim):
        super().__init__()
        self.n_head, self.d_k, self.d_v = n_head, d_k, d_v
        self.w_qs = nn.Linear(input_dim, n_head * d_k, bias=False)
        self.w_ks = nn.Linear(input_dim, n_head * d_k, bias=False)
        self.w_vs = nn.Linear(input_dim, n_head * d_v, bias=False)
        nn.init.normal_(self.w_qs.weight, mean=0, std=np.sqrt(2.0 / (d_model + d_k)))   # Modules.py:485-490
        nn.init.normal_(self.w_ks.weight, mean=0, std=np.sqrt(2.0 / (d_model + d_k)))
        nn.init.normal_(self.w_vs.weight, mean=0, std=np.sqrt(2.0 / (d_model + d_v)))
        self.fc1 = nn.Linear(n_head * d_v, d_model)
        self.fc2 = nn.Linear(n_head * d_v, d_model)        # never contributes to the output (Modules.py:573,617)
        self.layer_norm1 = nn.LayerNorm(input_dim)
        self.layer_norm2 = nn.LayerNorm(input_dim)
        self.layer_norm3 = nn.LayerNorm(input_dim)
        self.dropout = nn.Dropout(dropout) if dropout is not None else None
        self.diag_mask_flag = diag_mask
        self.diag_mask = None

    def forward(self, *a, **k):
        raise RuntimeError("MultiHeadAttention: " + _FUSED)


class EncoderLayer(nn.Module):
    def __init__(self, n_head, d_model, d_k, d_v, dropout_mul, dropout_pff, diag_mask, bottle_neck):
        super().__init__()
        self.n_head, self.d_k, self.d_v = n_head, d_k, d_v
        self.mul_head_attn = MultiHeadAttention(n_head, d_model, d_k, d_v, dropout=dropout_mul, diag_mask=diag_mask,
                                                input_dim=bottle_neck)
        self.pff_n1 = PositionwiseFeedForward([d_model, d_model, d_model], dropout=dropout_pff, residual=True,
                                              layer_norm=True)
        self.pff_n2 = PositionwiseFeedForward([bottle_neck, d_model, d_model], dropout=dropout_pff, residual=False,
                                              layer_norm=True)

    def forward(self, *a, **k):
        raise RuntimeError("EncoderLayer: " + _FUSED)


def zscore_positive_rows_host(inter_initial):
    """Modules.py:147-152 on the host: z-score (ddof 0) the positive entries of every row among themselves, NaN -> 0,
    IN PLACE like the reference (which mutates the caller's array).  Used only where no CUDA device exists (building a
    module to inspect its state_dict on a CPU box) and as the checker of the device kernel."""
    import scipy.stats
    for i in range(len(inter_initial)):
        row = inter_initial[i, :]
        pos = row > 0
        if pos.any():
            with np.errstate(invalid="ignore", divide="ignore"):
                inter_initial[i, pos] = scipy.stats.mstats.zscore(row[pos]).astype("float32")
    inter_initial[np.isnan(inter_initial)] = 0.0
    return inter_initial


def zscore_positive_rows(inter_initial):
    """Modules.py:147-152.  With a CUDA device the N rows are z-scored by one kernel launch (csrc/features.cu) and the
    result is written back into the caller's array (the reference mutates it too); the reference's Python loop over N
    rows around scipy costs 14 s at 30k bins."""
    if torch.cuda.is_available() and isinstance(inter_initial, np.ndarray) and inter_initial.dtype == np.float32:
        from .features import zscore_positive_rows_
        t = torch.from_numpy(inter_initial).cuda()
        zscore_positive_rows_(t)
        inter_initial[...] = t.cpu().numpy()
        return inter_initial
    return zscore_positive_rows_host(inter_initial)


class MultipleEmbedding(nn.Module):
    """Per-chromosome feature tables + tied encoder weights + reconstruction heads (Modules.py:125-174)."""

    def __init__(self, embedding_weights, dim, sparse=True, num_list=None, chrom_range=None, inter_initial=None):
        super().__init__()
        self.chrom_range = chrom_range
        self.num_list = torch.tensor([0] + [int(v) for v in num_list]).to(device)
        self.dim = dim
        self.embeddings = [SparseEmbedding(w, sparse) for w in embedding_weights]
        if inter_initial is not None:
            self.inter_initial = SparseEmbedding(zscore_positive_rows(inter_initial), sparse)
        else:
            self.inter_initial = SparseEmbedding(embedding_weights[-1], sparse)
        self.input_size = [e.width for e in self.embeddings]
        self.wstack = [TiedAutoEncoder([self.input_size[i], self.dim, self.dim], use_bias=False).to(device)
                       for i in range(len(self.embeddings))]
        self.next_w = FeedForward([self.dim, self.dim]).to(device)          # registered, never used (Modules.py:165)
        self.recon = [FeedForward([self.dim, int(v[1] - v[0])]).to(device) for v in self.chrom_range]
        for i, w in enumerate(self.wstack):
            self.add_module("Embedding_Linear%d" % i, w)
            self.add_module("Embedding_recon%d" % i, self.recon[i])
        self.dropout = nn.Dropout(0.2)

    def forward(self, x):
        raise RuntimeError("MultipleEmbedding: " + _FUSED + "; use Classifier.get_node_embeddings")


class Classifier(nn.Module):
    """Hyper-SAGNN hyperedge classifier (Modules.py:204-318) executed by the CUDA engine."""

    def __init__(self, n_head, d_model, d_k, d_v, node_embedding, diag_mask, bottle_neck, attribute_dict=None, **args):
        super().__init__()
        self.pff_classifier = PositionwiseFeedForward([d_model, 1], reshape=True, use_bias=True)
        self.node_embedding = node_embedding
        self.encode1 = EncoderLayer(n_head, d_model, d_k, d_v, dropout_mul=0.3, dropout_pff=0.4, diag_mask=diag_mask,
                                    bottle_neck=bottle_neck)
        self.encode2 = EncoderLayer(n_head, d_model, d_k, d_v, dropout_mul=0.3, dropout_pff=0.4, diag_mask=diag_mask,
                                    bottle_neck=bottle_neck)      # constructed, never called (Modules.py:272)
        self.diag_mask_flag = diag_mask
        self.layer_norm1 = nn.LayerNorm(d_model)
        self.layer_norm2 = nn.LayerNorm(d_model)
        self.next_w = FeedForward([bottle_neck, bottle_neck]).to(device)
        if attribute_dict is not None:
            table = torch.from_numpy(np.asarray(attribute_dict, dtype=np.float32)).to(device)
            self.attribute_dict_embedding = nn.Embedding(len(table), 1, padding_idx=0)
            self.attribute_dict_embedding.weight = nn.Parameter(table)
            self.attribute_dict_embedding.weight.requires_grad = False
            self.attribute_nn = nn.Linear(table.shape[-1], bottle_neck)
            self.attribute_dict = self.attribute_dict_embedding         # second registration, as Modules.py:249

    # -- engine plumbing -------------------------------------------------------------------
    def _engine(self):
        eng = self.__dict__.get("_matcha_engine")
        if eng is None:
            from .engine import Engine
            eng = Engine(self)
            self.__dict__["_matcha_engine"] = eng
        return eng

    def __getstate__(self):       # torch.save(model): device caches are rebuilt lazily after loading
        state = self.__dict__.copy()
        state.pop("_matcha_engine", None)
        return state

    def _draw_recon_chrom(self):
        # the reference draws one chromosome per forward from numpy's global stream (Modules.py:192)
        return int(np.random.choice(np.arange(len(self.node_embedding.chrom_range)), 1)[0])

    # -- reference API ---------------------------------------------------------------------
    def get_node_embeddings(self, x, return_recon=False):
        """x [b, L] int64 -> [b, L, d] (eval-mode encoder output; Modules.py:252-259)."""
        sz_b, len_seq = x.shape
        out = self._engine().node_embeddings(x.reshape(-1), training=self.training)
        out = out.view(sz_b, len_seq, out.shape[-1])
        if return_recon:
            return out, torch.zeros(1, device=out.device)
        return out

    def forward(self, x, mask=None, get_outlier=None, return_recon=False):
        """x int64 [B, L] (0 = pad) -> raw logits [B, 1] (and the reconstruction loss [1])."""
        x = x.long()
        rchrom = self._draw_recon_chrom()
        logits, recon = self._engine().forward(x, training=self.training, random_chrom=rchrom if return_recon else -1)
        if return_recon:
            return logits, recon
        return logits


class DataGenerator:
    """Per-size pools of positive hyperedges, `num_batch_per_iter * batch_size` per size per call
    (Modules.py:620-681).  numpy >= 1.24 refuses the ragged arrays the reference builds, so hyperedges
    are held zero-padded to `max_size` columns (int64)."""

    def __init__(self, edges, edge_weight, batch_size, num_batch_per_iter, min_size=2, max_size=2, flag=False, rng=None):
        # rng: a np.random.RandomState shared by construction across data-parallel ranks (same seed on every rank), so
        # `rank::world` slices partition ONE global batch; None keeps the reference's global np.random (Modules.py:651)
        self.rng = rng
        edges = pad_edges(edges, max_size)
        edge_weight = np.asarray(edge_weight)
        sizes = (edges != 0).sum(1)
        self.batch_size, self.num_batch_per_iter = batch_size, num_batch_per_iter
        self.min_size, self.max_size, self.flag = min_size, max_size, flag
        self.edges = [np.zeros((0, max_size), dtype=np.int64) for _ in range(max_size + 1)]
        self.edge_weight = [np.zeros((0,), dtype=edge_weight.dtype) for _ in range(max_size + 1)]
        need = num_batch_per_iter * batch_size
        for k in range(min_size, max_size + 1):
            e, w = edges[sizes == k], edge_weight[sizes == k]
            if len(e) == 0:
                continue
            while len(e) <= need:
                e, w = np.concatenate([e, e]), np.concatenate([w, w])
            self.edges[k], self.edge_weight[k] = e, w
            self.shuffle(k)
        self.pointer = np.zeros(len(self.edges), dtype="int")

    def shuffle(self, i):
        index = (self.rng or np.random).permutation(len(self.edges[i]))
        self.edges[i], self.edge_weight[i] = self.edges[i][index], self.edge_weight[i][index]

    def next_iter(self):
        need = self.num_batch_per_iter * self.batch_size
        out_e, out_w = [], []
        for k in range(self.min_size, self.max_size + 1):
            if len(self.edges[k]) == 0:
                continue
            start = self.pointer[k]
            self.pointer[k] += need
            if self.pointer[k] <= len(self.edges[k]):
                out_e.append(self.edges[k][start:self.pointer[k]])
                out_w.append(self.edge_weight[k][start:self.pointer[k]])
            else:
                e, w = self.edges[k][start:], self.edge_weight[k][start:]
                self.shuffle(k)
                left = need - len(e)
                self.pointer[k] = left
                out_e.append(np.concatenate([e, self.edges[k][:left]]))
                out_w.append(np.concatenate([w, self.edge_weight[k][:left]]))
        return np.concatenate(out_e), np.concatenate(out_w)


def pad_edges(edges, width=None):
    """list of hyperedges (ragged or rectangular) -> int64 [n, width], zero padded, rows sorted ascending."""
    if isinstance(edges, np.ndarray) and edges.dtype != object and edges.ndim == 2:
        e = edges.astype(np.int64)
        if width is not None and e.shape[1] < width:
            e = np.concatenate([e, np.zeros((len(e), width - e.shape[1]), dtype=np.int64)], 1)
        return e
    rows = [np.asarray(r, dtype=np.int64).reshape(-1) for r in edges]
    width = width or max((len(r) for r in rows), default=0)
    out = np.zeros((len(rows), width), dtype=np.int64)
    for i, r in enumerate(rows):
        out[i, :len(r)] = r
    return out
