"""Denoise post-processing on the GPU (reference: denoise_contact.py:31-61 proba2matrix, :160-192 per-chromosome tail).

The all-pairs scorer leaves the sigmoid scores of one chromosome packed in HBM (generate_pair_wise order).  The reference
copies them to the host, loops over 3e8 pairs in Python to fetch the observed contacts (:160), scatters both into dense
matrices with np.add.at, makes six full-matrix numpy passes and fits / applies a sklearn QuantileTransformer.  Here the
scores never leave the device: `csrc/denoise.cu` builds both symmetric matrices, the three sqrt-coverage normalisations,
max(p * obs, p) and the gap masking in four streaming kernels, and the uniform quantile map runs element-wise on the
device against a table fitted the way sklearn fits it (a 10 000-element random subsample: 40 KB to the host and back).
Only the finished matrix / pixel values are copied out, because writing them is the script's output.
"""
from __future__ import annotations

import numpy as np
import torch

from ._lib import MatchaError, check, load, ptr, stream_ptr


class QuantileUniform:
    """sklearn.preprocessing.QuantileTransformer(n_quantiles, output_distribution="uniform") for ONE column that lives on
    the device (denoise_contact.py:107,189).  fit = sklearn's `_dense_fit`: draw `subsample` rows without replacement,
    np.nanpercentile at linspace(0, 1, n_quantiles); transform = sklearn's `_transform_col` (mean of the ascending and the
    mirrored np.interp, bounds snapped to 0 / 1) as a CUDA kernel.

    Up to 2**24 values the subsample indices are drawn exactly as sklearn.utils.resample does (arange + RandomState.shuffle),
    so a seeded RandomState reproduces sklearn's table bit for bit; beyond that (cfg4: 6.2e8 values, where that shuffle
    alone costs the host ~10 s and 5 GB) indices come from a direct draw of distinct positions -- the same distribution."""

    def __init__(self, n_quantiles=1000, subsample=10_000, random_state=None):
        self.n_quantiles, self.subsample = int(n_quantiles), int(subsample)
        self.rng = random_state if isinstance(random_state, np.random.RandomState) else np.random.RandomState(random_state)
        self.quantiles_ = self.references_ = None

    def _subsample_indices(self, n):
        if n <= (1 << 24):
            idx = np.arange(n)
            self.rng.shuffle(idx)
            return idx[:self.subsample]
        got = np.unique(self.rng.randint(0, n, size=int(self.subsample * 1.25), dtype=np.int64))
        while len(got) < self.subsample:
            got = np.unique(np.concatenate([got, self.rng.randint(0, n, size=self.subsample, dtype=np.int64)]))
        return self.rng.permutation(got)[:self.subsample]

    def fit(self, x: torch.Tensor):
        lib = load()
        flat = x.reshape(-1)
        n = flat.numel()
        nq = max(1, min(self.n_quantiles, n))
        self.references_ = np.linspace(0, 1, nq, endpoint=True)
        if self.subsample < n:
            idx = torch.from_numpy(np.ascontiguousarray(self._subsample_indices(n), dtype=np.int64)).to(flat.device)
            sub = torch.empty(idx.numel(), dtype=torch.float32, device=flat.device)
            check(lib.matcha_gather_f32(ptr(flat), ptr(idx), idx.numel(), ptr(sub), stream_ptr()), "matcha_gather_f32")
            col = sub.cpu().numpy()
        else:
            col = flat.cpu().numpy()
        self.quantiles_ = np.nanpercentile(col.reshape(-1, 1), self.references_ * 100, axis=0)[:, 0].astype(np.float64)
        return self

    def transform_(self, x: torch.Tensor):
        """In place on a contiguous fp32 CUDA tensor."""
        if self.quantiles_ is None:
            raise MatchaError("QuantileUniform.transform_ before fit")
        if not (x.is_cuda and x.dtype == torch.float32 and x.is_contiguous()):
            raise MatchaError("QuantileUniform works in place on contiguous fp32 CUDA tensors")
        q = torch.from_numpy(self.quantiles_).to(x.device)
        r = torch.from_numpy(np.ascontiguousarray(self.references_, dtype=np.float64)).to(x.device)
        check(load().matcha_quantile_uniform(ptr(x), x.numel(), ptr(q), ptr(r), len(self.quantiles_), stream_ptr()),
              "matcha_quantile_uniform")
        return x

    def fit_transform_(self, x):
        return self.fit(x).transform_(x)


def denoise_matrix(proba: torch.Tensor, origin_block: torch.Tensor, n: int, min_dis: int = 0, transformer: QuantileUniform | None = None,
                   want_pixels: bool = False):
    """proba: CUDA fp32 [pair_count(n, min_dis)] sigmoid scores of one chromosome in generate_pair_wise order;
    origin_block: CUDA fp32 view [n, >= n] of the observed contacts of the same bins (any row stride).
    Returns the reference's `my` [n, n] on the device (after the quantile map when a transformer is given) and, if
    want_pixels, its values at the scored pairs in the same packed order (the `balanced` column, denoise_contact.py:205)."""
    lib = load()
    n, min_dis = int(n), int(min_dis)
    full = n - min_dis
    if full < 1:
        raise MatchaError("min_distance leaves no pairs")
    if not (proba.is_cuda and proba.dtype == torch.float32 and proba.is_contiguous() and proba.numel() == full * (full + 1) // 2):
        raise MatchaError("proba must be a contiguous fp32 CUDA vector of pair_count(n, min_dis) scores")
    if not (origin_block.is_cuda and origin_block.dtype == torch.float32 and origin_block.stride(1) == 1
            and origin_block.shape[0] >= n and origin_block.shape[1] >= n):
        raise MatchaError("origin_block must be an fp32 CUDA matrix view with unit column stride")
    my = torch.empty(n, n, dtype=torch.float32, device=proba.device)
    nbytes = int(lib.matcha_denoise_workspace_bytes(n))
    ws = torch.empty(nbytes, dtype=torch.uint8, device=proba.device)
    check(lib.matcha_denoise_matrix(ptr(proba), ptr(origin_block), origin_block.stride(0), n, min_dis, ptr(my), ptr(ws), nbytes,
                                    stream_ptr()), "matcha_denoise_matrix")
    del ws
    if transformer is not None:
        transformer.fit_transform_(my)
    if not want_pixels:
        return my
    pix = torch.empty(proba.numel(), dtype=torch.float32, device=proba.device)
    check(lib.matcha_pair_gather(ptr(my), n, min_dis, ptr(pix), stream_ptr()), "matcha_pair_gather")
    return my, pix
