"""Binding between a ``Classifier`` module tree and the CUDA engine behind the C ABI.

The engine owns three flat fp32 device buffers -- parameters, gradients, derived (folded) weights --
and re-points every LIVE ``nn.Parameter`` of the module at a view of the flat parameter buffer, so
``state_dict`` / ``load_state_dict`` / ``torch.save`` / torch optimizers keep working while the kernels
see one contiguous allocation (which is also the NCCL all-reduce buffer for data-parallel training).
PyTorch is used here for device memory, streams and autograd plumbing only.
"""
from __future__ import annotations

import ctypes as C
import weakref

import numpy as np
import torch

from . import _lib
from ._lib import MatchaError, ModelDesc, check, ptr, stream_ptr

_ALIGN = 64           # element alignment of every tensor inside the flat buffers
_EVAL_CHUNK = 32768   # hyperedges per eval launch (bounds the workspace)
_M64 = (1 << 64) - 1


def _require_cuda():
    if not torch.cuda.is_available():
        raise MatchaError("matcha_b200 needs a CUDA device (sm_100a): the hot path has no CPU fallback")


class Engine:
    def __init__(self, model):
        self._model = weakref.ref(model)
        self.lib = _lib.load()
        self.bound = False
        self.tape_id = 0
        self.seed_base = int(torch.initial_seed()) & _M64
        self._ws = None
        self._prepared_version = None

    # ------------------------------------------------------------------------------------
    # parameter inventory
    # ------------------------------------------------------------------------------------
    def _inventory(self):
        """[(field, Parameter)] for the always-live tensors and per-chromosome [(w0, w1, rw, rb)]."""
        m = self._model()
        a, p = m.encode1.mul_head_attn, m.encode1.pff_n1
        always = [
            ("off_attr_w", m.attribute_nn.weight), ("off_attr_b", m.attribute_nn.bias),
            ("off_next_w", m.next_w.FF_Linear0.weight), ("off_next_b", m.next_w.FF_Linear0.bias),
            ("off_lnq_g", a.layer_norm1.weight), ("off_lnq_b", a.layer_norm1.bias),
            ("off_lnk_g", a.layer_norm2.weight), ("off_lnk_b", a.layer_norm2.bias),
            ("off_lnv_g", a.layer_norm3.weight), ("off_lnv_b", a.layer_norm3.bias),
            ("off_wq", a.w_qs.weight), ("off_wk", a.w_ks.weight), ("off_wv", a.w_vs.weight),
            ("off_fc1_w", a.fc1.weight), ("off_fc1_b", a.fc1.bias),
            ("off_pff_w0", p.PWF_Conv0.weight), ("off_pff_b0", p.PWF_Conv0.bias),
            ("off_pff_w1", p.PWF_Conv1.weight), ("off_pff_b1", p.PWF_Conv1.bias),
            ("off_pff_g", p.layer_norm.weight), ("off_pff_b", p.layer_norm.bias),
            ("off_ln1_g", m.layer_norm1.weight), ("off_ln1_b", m.layer_norm1.bias),
            ("off_ln2_g", m.layer_norm2.weight), ("off_ln2_b", m.layer_norm2.bias),
            ("off_cls_w", m.pff_classifier.PWF_Conv0.weight), ("off_cls_b", m.pff_classifier.PWF_Conv0.bias),
        ]
        ne = m.node_embedding
        per_chrom = []
        for c in range(len(ne.chrom_range)):
            enc = getattr(ne, "Embedding_Linear%d" % c)
            rec = getattr(ne, "Embedding_recon%d" % c)
            per_chrom.append((enc._parameters["tied weight_0"], enc._parameters["tied weight_1"],
                              rec.FF_Linear0.weight, rec.FF_Linear0.bias))
        return always, per_chrom

    def live_parameters(self):
        always, per_chrom = self._inventory()
        return [p for _, p in always] + [p for grp in per_chrom for p in grp]

    # ------------------------------------------------------------------------------------
    # binding
    # ------------------------------------------------------------------------------------
    def _bind(self):
        _require_cuda()
        m = self._model()
        dev = torch.device("cuda", torch.cuda.current_device())
        ne = m.node_embedding
        cr = np.asarray(ne.chrom_range, dtype=np.int64)
        C_ = len(cr)
        if C_ > _lib.MAX_CHROM:
            raise MatchaError(f"{C_} chromosomes exceed MATCHA_MAX_CHROM={_lib.MAX_CHROM}")
        always, per_chrom = self._inventory()
        d = m.next_w.FF_Linear0.weight.shape[0]

        # layout: [always-live tensors | per-chromosome tensors], each aligned to _ALIGN elements
        off, layout = 0, []
        def place(p):
            nonlocal off
            o = off
            off += (p.numel() + _ALIGN - 1) // _ALIGN * _ALIGN
            return o
        for _, p in always:
            layout.append((p, place(p)))
        self.n_always = off
        self.segments = []            # (begin, end, flag index)
        for c, grp in enumerate(per_chrom):
            for j, p in enumerate(grp):
                o = place(p)
                layout.append((p, o))
                self.segments.append((o, o + p.numel(), c if j < 2 else C_ + c))
        self.n_flat = off
        flat = torch.zeros(self.n_flat, dtype=torch.float32, device=dev)
        for p, o in layout:
            view = flat[o:o + p.numel()].view(p.shape)
            view.copy_(p.data.to(dev, torch.float32))
            p.data = view
        self.flat, self.layout = flat, layout
        self.gflat = torch.zeros(self.n_flat + 2 * C_, dtype=torch.float32, device=dev)   # tail: activity flags (DP)
        self.active = torch.zeros(2 * C_, dtype=torch.int32, device=dev)

        desc = ModelDesc()
        desc.d, desc.n_head, desc.n_chrom = d, m.encode1.mul_head_attn.n_head, C_
        desc.n_nodes = int(cr[:, 1].max() - 1)
        desc.params, desc.grads = flat.data_ptr(), self.gflat.data_ptr()
        pos = {id(p): o for p, o in layout}
        for name, p in always:
            setattr(desc, name, pos[id(p)])
        self._keep = []
        for c, grp in enumerate(per_chrom):
            desc.chrom_start[c], desc.chrom_end[c] = int(cr[c, 0]), int(cr[c, 1])
            desc.off_w0[c], desc.off_w1[c], desc.off_rw[c], desc.off_rb[c] = (pos[id(p)] for p in grp)
            emb = ne.embeddings[c]
            n_c = int(cr[c, 1] - cr[c, 0])
            if getattr(emb, "sparse", False):
                # CSR feature rows (Modules.py:58-65, sparse=True): uploaded as CSR, consumed by the SpMM encoder kernels
                csr = emb.embedding.tocsr()
                if csr.shape[0] != n_c:
                    raise MatchaError(f"chromosome {c}: CSR feature table {csr.shape} does not match {n_c} bins")
                indptr = torch.from_numpy(np.ascontiguousarray(csr.indptr, dtype=np.int64)).to(dev)
                indices = torch.from_numpy(np.ascontiguousarray(csr.indices, dtype=np.int32)).to(dev)
                values = torch.from_numpy(np.ascontiguousarray(csr.data, dtype=np.float32)).to(dev)
                self._keep += [indptr, indices, values]
                desc.feat[c], desc.feat_ld[c] = None, 0
                desc.feat_indptr[c], desc.feat_indices[c], desc.feat_values[c] = indptr.data_ptr(), indices.data_ptr(), values.data_ptr()
                continue
            f = emb.embedding.to(dev, torch.float32)
            if f.shape != (n_c, n_c) and f.shape[0] != n_c:
                raise MatchaError(f"chromosome {c}: feature table {tuple(f.shape)} does not match {n_c} bins")
            ld = (f.shape[1] + 3) // 4 * 4        # 16-byte aligned rows for 128-bit loads
            if ld != f.shape[1] or not f.is_contiguous():
                fp = torch.zeros(f.shape[0], ld, dtype=torch.float32, device=dev)
                fp[:, :f.shape[1]] = f
                f = fp
            self._keep.append(f)
            desc.feat[c], desc.feat_ld[c] = f.data_ptr(), ld
        attr = m.attribute_dict_embedding.weight.data.to(dev, torch.float32).contiguous()
        self._keep.append(attr)
        desc.attr_dim, desc.attr_table = attr.shape[1], attr.data_ptr()
        inter = getattr(ne, "inter_initial", None)
        self.inter_problem = None            # why no reconstruction target is bound (reported when a caller asks for one)
        if inter is None:
            self.inter_problem = "node_embedding.inter_initial is None"
        elif getattr(inter, "sparse", False):
            self.inter_problem = "a scipy-sparse inter_initial is not supported (pass the dense [N, N] array of main.py:569)"
        elif tuple(inter.embedding.shape) != (desc.n_nodes, desc.n_nodes):
            self.inter_problem = f"inter_initial has shape {tuple(inter.embedding.shape)}, expected [{desc.n_nodes}, {desc.n_nodes}]"
        else:
            it = inter.embedding.to(dev, torch.float32).contiguous()
            self._keep.append(it)
            desc.inter, desc.inter_ld = it.data_ptr(), it.shape[1]
        n_der = self.lib.matcha_derived_elems(C.byref(desc))
        self.derived = torch.zeros(n_der, dtype=torch.float32, device=dev)
        self.derived_grad = torch.zeros(n_der, dtype=torch.float32, device=dev)
        desc.derived, desc.derived_grad = self.derived.data_ptr(), self.derived_grad.data_ptr()
        desc.p_feature = float(ne.dropout.p)
        desc.p_attn = float(m.encode1.mul_head_attn.dropout.p) if m.encode1.mul_head_attn.dropout is not None else 0.0
        desc.p_pff = float(m.encode1.pff_n1.dropout.p) if m.encode1.pff_n1.dropout is not None else 0.0
        self.desc, self.dev, self.d, self.C = desc, dev, d, C_
        self.chrom_range = cr
        self.bound = True

    def ensure_bound(self):
        if not self.bound:
            self._bind()
            return
        base = self.flat.data_ptr()
        for p, o in self.layout:
            if p.data_ptr() != base + 4 * o:      # .to() / .cuda() / external re-assignment replaced the storage
                self._bind()
                return
        ne = self._model().node_embedding         # dropout probabilities may be edited by callers (tests do)
        self.desc.p_feature = float(ne.dropout.p)
        mha, pff = self._model().encode1.mul_head_attn, self._model().encode1.pff_n1
        self.desc.p_attn = float(mha.dropout.p) if mha.dropout is not None else 0.0
        self.desc.p_pff = float(pff.dropout.p) if pff.dropout is not None else 0.0

    def _workspace(self, B, L, training):
        need = self.lib.matcha_workspace_bytes(C.byref(self.desc), B, L, 1 if training else 0)
        if need < 0:
            raise MatchaError("matcha_workspace_bytes failed")
        if self._ws is None or self._ws.numel() < need:
            self._ws = None
            self._ws = torch.empty(int(need * 1.25) + 1024, dtype=torch.uint8, device=self.dev)
        return self._ws

    def prepare(self):
        check(self.lib.matcha_prepare(C.byref(self.desc), stream_ptr()), "matcha_prepare")

    def next_seed(self):
        self.tape_id += 1
        return (self.seed_base * 0x9E3779B97F4A7C15 + self.tape_id * 0xD1B54A32D192ED03) & _M64

    # ------------------------------------------------------------------------------------
    # raw passes (no autograd)
    # ------------------------------------------------------------------------------------
    def require_inter(self):
        """Modules.py:192-199 needs the z-scored inter-chromosomal table; without one the loss would silently be 0."""
        if not self.desc.inter:
            raise MatchaError("reconstruction loss requested but no inter-chromosomal target is bound: " + str(self.inter_problem))

    def run_forward(self, x, training, seed, random_chrom, logits=None, recon=None):
        B, L = x.shape
        ws = self._workspace(B, L, training)
        if logits is None:
            logits = torch.empty(B, dtype=torch.float32, device=self.dev)
        if recon is None:
            recon = torch.zeros(1, dtype=torch.float32, device=self.dev)
        check(self.lib.matcha_forward(C.byref(self.desc), ptr(x), B, L, 1 if training else 0, seed, random_chrom,
                                      ptr(logits), ptr(recon), ptr(ws), ws.numel(), stream_ptr()), "matcha_forward")
        return logits, recon

    def run_backward(self, x, seed, random_chrom, dlogit, beta, want_active=True):
        B, L = x.shape
        ws = self._workspace(B, L, True)
        check(self.lib.matcha_backward(C.byref(self.desc), ptr(x), B, L, seed, random_chrom, ptr(dlogit), float(beta),
                                       ptr(self.active) if want_active else 0, ptr(ws), ws.numel(), stream_ptr()),
              "matcha_backward")

    # ------------------------------------------------------------------------------------
    # Classifier.forward / get_node_embeddings
    # ------------------------------------------------------------------------------------
    def forward(self, x, training, random_chrom):
        self.ensure_bound()
        if x.dim() != 2:
            raise MatchaError("x must be [B, L]")
        x = x.to(self.dev, torch.int64).contiguous()
        B, L = x.shape
        if B > 0:
            # nn.Embedding in the reference raises on ids outside [0, N] (Modules.py:243-249); the kernels index the
            # attribute / feature tables with them, so the public entry point checks (one small reduction + sync)
            lo, hi = torch.aminmax(x)
            lo, hi = int(lo), int(hi)
            if lo < 0 or hi > self.desc.n_nodes:
                raise MatchaError(f"node ids must lie in [0, {self.desc.n_nodes}] (0 = padding): got [{lo}, {hi}]")
        if B == 0:      # empty batch: nothing to launch (an empty tensor has no device pointer to hand to the library)
            return (torch.empty(0, 1, dtype=torch.float32, device=self.dev),
                    torch.zeros(1, dtype=torch.float32, device=self.dev))
        if random_chrom >= 0 and self.C > 1:
            self.require_inter()              # return_recon=True with no usable target: fail, do not report 0 (Modules.py:192-199)
        self.prepare()
        if training and torch.is_grad_enabled():
            anchor = self._model().layer_norm1.weight
            logits, recon = _HyperedgeFn.apply(self, x, random_chrom, anchor)
            return logits.view(B, 1), recon
        seed = self.next_seed()
        if training or B <= _EVAL_CHUNK:
            logits, recon = self.run_forward(x, training, seed, random_chrom)
            return logits.view(B, 1), recon
        logits = torch.empty(B, dtype=torch.float32, device=self.dev)
        recon = torch.zeros(1, dtype=torch.float32, device=self.dev)
        for s in range(0, B, _EVAL_CHUNK):     # eval: chunk to bound the workspace (recon is not defined chunk-wise)
            e = min(B, s + _EVAL_CHUNK)
            self.run_forward(x[s:e], False, seed, -1, logits=logits[s:e])
        return logits.view(B, 1), recon

    def node_embeddings(self, ids, training=False):
        self.ensure_bound()
        ids = ids.to(self.dev, torch.int64).contiguous().view(-1)
        T = ids.numel()
        out = torch.empty(T, self.d, dtype=torch.float32, device=self.dev)
        if T == 0:
            return out
        self.prepare()
        if training and self.desc.p_feature > 0:
            raise MatchaError("train-mode get_node_embeddings (feature dropout) is only defined inside forward()")
        ws = self._workspace(T, 1, False)
        check(self.lib.matcha_node_embeddings(C.byref(self.desc), ptr(ids), T, ptr(out), ptr(ws), ws.numel(),
                                              stream_ptr()), "matcha_node_embeddings")
        return out

    def pair_tables(self):
        """Per-node tables (D, S) of the k = 2 closed form, each [N + 1, d]."""
        self.ensure_bound()
        self.prepare()
        T = self.desc.n_nodes + 1
        D = torch.empty(T, self.d, dtype=torch.float32, device=self.dev)
        S = torch.empty(T, self.d, dtype=torch.float32, device=self.dev)
        ws = self._workspace(T, 1, False)
        check(self.lib.matcha_pair_tables(C.byref(self.desc), ptr(D), ptr(S), ptr(ws), ws.numel(), stream_ptr()),
              "matcha_pair_tables")
        return D, S

    # ------------------------------------------------------------------------------------
    # gradient views
    # ------------------------------------------------------------------------------------
    def grad_view(self, p, o):
        return self.gflat[o:o + p.numel()].view(p.shape)

    def publish_grads(self, active_host):
        """Give every parameter that received a gradient a `.grad` view of the flat gradient buffer; the
        others keep `.grad = None`, which is what torch's autograd leaves behind in the reference."""
        always_n = len(self._inventory()[0])
        for i, (p, o) in enumerate(self.layout):
            if i < always_n:
                p.grad = self.grad_view(p, o)
            else:
                j = i - always_n
                c, which = divmod(j, 4)
                flag = active_host[c] if which < 2 else active_host[self.C + c]
                p.grad = self.grad_view(p, o) if flag else None


class _HyperedgeFn(torch.autograd.Function):
    """Train-mode Classifier.forward under torch autograd: lets the reference's own training loop
    (loss_func(pred, y, weight=w); loss.backward(); opt.step() -- main.py:164-183) drive the CUDA backward."""

    @staticmethod
    def forward(ctx, engine, x, random_chrom, anchor):
        seed = engine.next_seed()
        logits, recon = engine.run_forward(x, True, seed, random_chrom)
        ctx.engine, ctx.x, ctx.seed, ctx.rchrom, ctx.tape = engine, x, seed, random_chrom, engine.tape_id
        return logits, recon

    @staticmethod
    def backward(ctx, dlogits, drecon):
        eng = ctx.engine
        if eng.tape_id != ctx.tape:
            raise MatchaError("backward() called after another training forward overwrote the activation tape")
        beta = float(drecon.reshape(-1)[0].item()) if drecon is not None else 0.0
        if dlogits is None:
            dlogits = torch.zeros(ctx.x.shape[0], dtype=torch.float32, device=eng.dev)
        dlogits = dlogits.reshape(-1).contiguous().float()
        live = [p for p, _ in eng.layout]
        if all(p.grad is None for p in live):
            eng.gflat.zero_()
        eng.run_backward(ctx.x, ctx.seed, ctx.rchrom if beta != 0.0 else -1, dlogits, beta)
        eng.publish_grads(eng.active.cpu().numpy())
        return None, None, None, None


class FlatAdamW:
    """torch.optim.AdamW(lr=1e-3) semantics (main.py:630) as one fused kernel over the flat buffers."""

    def __init__(self, engine: Engine, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01):
        engine.ensure_bound()
        self.e = engine
        self.lr, self.betas, self.eps, self.wd = lr, betas, eps, weight_decay
        self.m1 = torch.zeros_like(engine.flat)
        self.m2 = torch.zeros_like(engine.flat)
        seg = engine.segments
        dev = engine.dev
        self.seg_begin = torch.tensor([s[0] for s in seg], dtype=torch.int64, device=dev)
        self.seg_end = torch.tensor([s[1] for s in seg], dtype=torch.int64, device=dev)
        self.seg_flag = torch.tensor([s[2] for s in seg], dtype=torch.int32, device=dev)
        self.seg_step = torch.zeros(len(seg), dtype=torch.int32, device=dev)
        self.t = 0

    def step(self, grad_scale=1.0):
        e = self.e
        self.t += 1
        check(e.lib.matcha_adamw(ptr(e.flat), ptr(e.gflat), ptr(self.m1), ptr(self.m2), e.n_always, self.t,
                                 len(e.segments), ptr(self.seg_begin), ptr(self.seg_end), ptr(self.seg_flag),
                                 ptr(self.seg_step), ptr(e.active), self.lr, self.betas[0], self.betas[1], self.eps,
                                 self.wd, float(grad_scale), stream_ptr()), "matcha_adamw")
