"""CPU ORACLE (TEST INFRASTRUCTURE -- never imported by the product path) of the denoise post-processing: the numpy tail
of the reference's ``denoise_contact.py`` (proba2matrix :31-61 for pairs, sqrt-coverage normalisation and combination with
the observed map :162-189), in the reference's order and dtypes, pinned against the unmodified script
(``oracle/make_denoise_golden.py`` -> ``tests/golden/denoise_small.npz``, ``tests/test_cpu_host.py``).  The product path is
the CUDA implementation in ``matcha_b200/denoise.py`` / ``csrc/denoise.cu``, checked against this file and the golden.
"""
from __future__ import annotations

import numpy as np


def fill_symmetric(n, ii, jj, vals):
    """proba2matrix (denoise_contact.py:31-61) for pairs: m[i, j] += v; m = m + m.T (the diagonal doubles)."""
    m = np.zeros((n, n), dtype="float32")
    np.add.at(m, (ii, jj), vals)
    return m + m.T


def sqrt_coverage_normalise(m):
    """denoise_contact.py:163-166 (and :171-174, :178-181): divide by the square roots of the row and column means."""
    c1 = np.sqrt(np.mean(m, axis=-1, keepdims=True))
    c2 = np.sqrt(np.mean(m, axis=0, keepdims=True))
    m = m / (c1 + 1e-15)
    return m / (c2 + 1e-15)


def denoise_matrix(n, ii, jj, proba, weight, transformer=None):
    """One chromosome: ii, jj 0-based bin indices of the scored pairs (generate_pair_wise order), proba their sigmoid
    scores, weight the observed contacts of the same pairs.  Returns the reference's ``my`` [n, n] (denoise_contact.py:189);
    with ``transformer=None`` the matrix before the final quantile transform (:185-186)."""
    my_proba = sqrt_coverage_normalise(fill_symmetric(n, ii, jj, proba))              # :162-166
    origin_part = fill_symmetric(n, ii, jj, weight)                                   # :168
    gap1, gap2 = origin_part.sum(-1) == 0, origin_part.sum(0) == 0                   # :169-170
    origin_part = sqrt_coverage_normalise(origin_part)                                # :171-174
    my = sqrt_coverage_normalise(np.maximum(my_proba * origin_part, my_proba))        # :177-181
    my[gap1, :] = 0.0                                                                 # :184-185
    my[:, gap2] = 0.0
    if transformer is None:
        return my
    return transformer.fit_transform(my.reshape((-1, 1))).reshape((n, -1))            # :189
