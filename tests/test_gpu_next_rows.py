"""GPU parity tests of the SURVEY section 8(f) "next" rows: device metrics (f3: utils.py:32-72), denoise post-processing
(f2: denoise_contact.py:31-61,160-192), feature construction (f4: main.py:569-577, Modules.py:147-152,
process.py:90-105,107-176).  Each CUDA path is checked against the reference's own library call (sklearn / numpy / scipy
on the same inputs), the CPU oracle, or golden matrices produced by executing the unmodified reference script."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ------------------------------------------------------------------------------------------
# f3: AUROC / AUPR / accuracy per hyperedge size
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,ties", [(37, False), (5000, False), (40000, True), (300_000, True), (2049, True)])
def test_device_metrics_match_sklearn(n, ties):
    from sklearn.metrics import average_precision_score, roc_auc_score
    from matcha_b200.metrics import binary_metrics
    rng = np.random.default_rng(n)
    y = (rng.random(n) < 0.3).astype(np.float32)
    s = (rng.normal(0, 1, n) + 1.2 * y).astype(np.float32)
    s = 1.0 / (1.0 + np.exp(-s))
    if ties:                                           # heavy ties: saturated sigmoids, a quantised block, signed zeros
        s = np.round(s, 2).astype(np.float32)
        s[: n // 10] = 1.0
        s[n // 10: n // 8] = 0.0
        s[n // 8: n // 7] = -0.0
    size = rng.integers(2, 6, n).astype(np.int64)
    if n > 1000:
        y[size == 4] = 1.0                             # one-label slice: sklearn raises there, the reference swallows it
    got = binary_metrics(torch.from_numpy(y).cuda(), torch.from_numpy(s).cuda(), torch.from_numpy(size).cuda(), max_size=5)
    assert got["all"][3] == n
    np.testing.assert_allclose(got["all"][0], roc_auc_score(y, s), rtol=0, atol=1e-12)
    np.testing.assert_allclose(got["all"][1], average_precision_score(y, s), rtol=0, atol=1e-12)
    np.testing.assert_allclose(got["all"][2], float(((s >= 0.5) == (y >= 0.5)).mean()), rtol=0, atol=1e-12)
    for k in np.unique(size):
        m = size == k
        row = got[int(k)]
        assert row[3] == m.sum()
        np.testing.assert_allclose(row[2], float(((s[m] >= 0.5) == (y[m] >= 0.5)).mean()), atol=1e-12)
        if y[m].min() == y[m].max():
            assert np.isnan(row[0]) and np.isnan(row[1])
        else:
            np.testing.assert_allclose(row[0], roc_auc_score(y[m], s[m]), rtol=0, atol=1e-12)
            np.testing.assert_allclose(row[1], average_precision_score(y[m], s[m]), rtol=0, atol=1e-12)


def test_metric_strings_have_the_reference_format():
    """utils.py:38-52,57-72: 'all 0.912 2 0.901 ...' and '2 0.910 3 0.880 '; main.py:313 parses split(' ')[-2]."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "scripts"))
    from utils import accuracy, roc_auc_cuda
    rng = np.random.default_rng(0)
    n = 4000
    y = torch.from_numpy((rng.random(n) < 0.25).astype(np.float32)).cuda().view(-1, 1)
    p = torch.from_numpy(rng.random(n).astype(np.float32)).cuda().view(-1, 1)
    size = torch.from_numpy(rng.integers(2, 6, n)).cuda()
    roc, aupr = roc_auc_cuda(y, p, size, None)
    acc = accuracy(p, y, size)
    toks = aupr.split(" ")
    assert toks[0] == "all" and toks[2::2] == ["2", "3", "4", "5"] and toks[-2] == "5"
    assert all(0.0 <= float(v) <= 1.0 for v in toks[1::2]) and len(roc.split(" ")) == 10
    assert acc.endswith(" ") and acc.split(" ")[0::2][:4] == ["2", "3", "4", "5"]
    assert accuracy(p, y).strip().replace(".", "").isdigit()


# ------------------------------------------------------------------------------------------
# f2: denoise post-processing
# ------------------------------------------------------------------------------------------
def test_denoise_postprocessing_on_device_matches_the_reference_script():
    """csrc/denoise.cu against the matrices produced by EXECUTING the reference's loop (tests/golden/denoise_small.npz):
    the matrix before the quantile map at 2e-6 (fp32, other summation order), the final `my` and the `balanced` pixels.
    The uniform quantile map is a rank transform whose slope is 1 / (1000 * gap between neighbouring quantiles), so a 1e-7
    input difference can move single elements by ~1e-4: final values are held to 2e-3 absolute, the map itself is held to
    1e-6 against sklearn on identical input below."""
    from oracle.denoise_oracle import denoise_matrix as denoise_host
    from matcha_b200.denoise import QuantileUniform, denoise_matrix
    from matcha_b200.scorer import pair_index_to_ij
    g = np.load(os.path.join(ROOT, "tests", "golden", "denoise_small.npz"))
    cr, origin = g["chrom_range"], g["origin"]
    origin_dev = torch.from_numpy(origin).cuda()
    for ci in range(2):
        min_dis = int(g[f"min_dis/{ci}"])
        for c, (lo, hi) in enumerate(cr):
            lo, hi = int(lo), int(hi)
            n = hi - lo
            logits = g[f"logits/{ci}/{c}"]
            proba = torch.sigmoid(torch.from_numpy(logits)).cuda()
            block = origin_dev[lo - 1:hi - 1, lo - 1:hi - 1]                  # strided view into intra_adj
            ii, jj = pair_index_to_ij(np.arange(len(logits)), lo, hi, min_dis)
            weight = origin[ii - 1, jj - 1]
            pre_host = denoise_host(n, ii - lo, jj - lo, proba.cpu().numpy(), weight, None)
            pre = denoise_matrix(proba, block, n, min_dis, None)
            np.testing.assert_allclose(pre.cpu().numpy(), pre_host, rtol=2e-6, atol=1e-7)
            torch.testing.assert_close(pre, pre.t().contiguous(), rtol=5e-7, atol=0)   # (x / c_i) / c_j vs (x / c_j) / c_i: fp32 rounding only
            my, pix = denoise_matrix(proba, block, n, min_dis, QuantileUniform(1000), want_pixels=True)
            np.testing.assert_allclose(my.cpu().numpy(), g[f"my/{ci}/{c}"], rtol=0, atol=2e-3)
            np.testing.assert_allclose(pix.cpu().numpy(), g[f"balanced/{ci}/{c}"], rtol=0, atol=2e-3)
            np.testing.assert_array_equal(pix.cpu().numpy(), my.cpu().numpy()[ii - lo, jj - lo])


@pytest.mark.parametrize("n,seed", [(3000, 0), (250_000, 1), (1_000_003, 2)])
def test_quantile_uniform_matches_sklearn(n, seed):
    """Same fp32 column, same seeded RandomState: the fitted table is sklearn's bit for bit (same subsample indices, same
    np.nanpercentile) and the device transform equals sklearn's _transform_col (many repeated quantiles: 40 % zeros)."""
    from sklearn.preprocessing import QuantileTransformer
    from matcha_b200.denoise import QuantileUniform
    rng = np.random.default_rng(seed)
    x = rng.gamma(0.7, 1.0, n).astype(np.float32)
    x[rng.random(n) < 0.4] = 0.0
    x[:5] = [np.float32(x.max() * 2), np.float32(0.0), np.float32(x.max() * 2), np.nan, np.float32(1e-30)]
    ref = QuantileTransformer(n_quantiles=1000, output_distribution="uniform", random_state=np.random.RandomState(seed))
    want = ref.fit_transform(x.copy().reshape(-1, 1)).reshape(-1)
    qt = QuantileUniform(1000, random_state=np.random.RandomState(seed))
    xd = torch.from_numpy(x.copy()).cuda()
    qt.fit_transform_(xd)
    np.testing.assert_array_equal(qt.quantiles_, ref.quantiles_[:, 0])
    got = xd.cpu().numpy()
    assert np.isnan(got[3]) and np.isnan(want[3])
    np.testing.assert_allclose(np.delete(got, 3), np.delete(want, 3), rtol=0, atol=1e-6)


def test_denoise_cfg4_width_properties():
    """configs[3] width (24,897 bins, 3.1e8 packed scores, 2.5 GB matrices): symmetric output, gap rows / columns zero,
    pixels are the matrix at the pairs, and a 64-row sample equals the host restatement of denoise_contact.py on the same
    rows (row means need the whole matrix, so the host side gets them from a float64 reduction of the device matrices)."""
    from matcha_b200.denoise import denoise_matrix
    n, md = 24897, 0
    total = n * (n + 1) // 2
    g = torch.Generator(device="cuda").manual_seed(3)
    proba = torch.rand(total, device="cuda", generator=g)
    origin = (torch.rand(n, n, device="cuda", generator=g) < 0.02).float() * torch.randint(1, 6, (n, n), device="cuda", generator=g).float()
    dead = torch.tensor([5, 777, 20000], device="cuda")
    origin[dead, :] = 0
    origin[:, dead] = 0
    my, pix = denoise_matrix(proba, origin, n, md, None, want_pixels=True)
    assert bool(torch.isfinite(my).all())
    assert float(my[dead].abs().max()) == 0.0 and float(my[:, dead].abs().max()) == 0.0
    blk = my[:4096, :4096]
    torch.testing.assert_close(blk, blk.t().contiguous(), rtol=5e-7, atol=0)
    # pixels: row 0 is the first n entries, the last entry is (n - 1, n - 1)
    assert bool((pix[:n] == my[0]).all()) and float(pix[-1]) == float(my[-1, -1])
    # host restatement on sampled rows
    up = torch.triu(origin)
    W = up + up.t()                                                      # m + m.T (:46): the diagonal doubles
    rows = torch.tensor([0, 1, 5, 1234, 20001, n - 1], device="cuda")
    idx = torch.arange(n, device="cuda")

    def packed(i, j):                                                    # pair index of (min, max)
        a, b = torch.minimum(i, j), torch.maximum(i, j)
        return a * n - a * (a - 1) // 2 + (b - a)
    P_rows = proba[packed(rows.view(-1, 1).expand(-1, n), idx.view(1, -1).expand(len(rows), -1))]
    P_rows[torch.arange(len(rows)), rows] *= 2
    # coverage vectors of the full matrices from float64 reductions
    rsP = torch.zeros(n, dtype=torch.float64, device="cuda")
    for s in range(0, n, 2048):
        e = min(n, s + 2048)
        r = torch.arange(s, e, device="cuda")
        blkP = proba[packed(r.view(-1, 1).expand(-1, n), idx.view(1, -1).expand(e - s, -1))]
        blkP[torch.arange(e - s), r] *= 2
        rsP[s:e] = blkP.double().sum(1)
    cP = torch.sqrt((rsP / n).float()) + 1e-15
    rsW = W.double().sum(1)
    cW = torch.sqrt((rsW / n).float()) + 1e-15
    p = P_rows / cP[rows].view(-1, 1) / cP.view(1, -1)
    o = W[rows] / cW[rows].view(-1, 1) / cW.view(1, -1)
    my0_rows = torch.maximum(p * o, p)
    # third coverage needs all rows of my0: recompute blockwise
    rsY = torch.zeros(n, dtype=torch.float64, device="cuda")
    for s in range(0, n, 2048):
        e = min(n, s + 2048)
        r = torch.arange(s, e, device="cuda")
        blkP = proba[packed(r.view(-1, 1).expand(-1, n), idx.view(1, -1).expand(e - s, -1))]
        blkP[torch.arange(e - s), r] *= 2
        pp = blkP / cP[s:e].view(-1, 1) / cP.view(1, -1)
        oo = W[s:e] / cW[s:e].view(-1, 1) / cW.view(1, -1)
        rsY[s:e] = torch.maximum(pp * oo, pp).double().sum(1)
    cY = torch.sqrt((rsY / n).float()) + 1e-15
    want = my0_rows / cY[rows].view(-1, 1) / cY.view(1, -1)
    gap = rsW == 0
    want[gap[rows]] = 0
    want[:, gap] = 0
    np.testing.assert_allclose(my[rows].cpu().numpy(), want.cpu().numpy(), rtol=3e-6, atol=1e-9)


# ------------------------------------------------------------------------------------------
# f4: feature construction
# ------------------------------------------------------------------------------------------
def test_adjacency_builders_match_the_reference_functions_bit_exact():
    """csrc/features.cu against arrays written by the UNMODIFIED process.py functions (edgelist2adj :90-105,
    parse_cool_contact :107-172) executed by oracle/make_features_golden.py: float64, bit for bit (every cell receives
    exactly-representable sums: integer counts, or one pixel weight per direction -- twice on the diagonal)."""
    from matcha_b200.features import adjacency_from_clusters, adjacency_from_pixels
    g = np.load(os.path.join(ROOT, "tests", "golden", "features_small.npz"))
    N = int(g["chrom_range"][-1, 1]) - 1
    adj = adjacency_from_clusters(g["members"], g["offsets"], N)
    np.testing.assert_array_equal(adj.cpu().numpy(), g["edge_adj"])
    intra, inter = adjacency_from_pixels(g["bin1"], g["bin2"], g["count"], g["cool2node"], g["node2chrom"], N)
    np.testing.assert_array_equal(intra.cpu().numpy(), g["intra"])
    np.testing.assert_array_equal(inter.cpu().numpy(), g["inter"])
    assert float(inter.sum()) > 0 and float(intra.diagonal().sum()) > 0


def test_adjacency_builders_at_cfg2_size_match_the_oracle():
    from oracle import features_oracle as FO
    from matcha_b200.features import adjacency_from_clusters, adjacency_from_pixels
    from matcha_b200.synthetic import chrom_bins, CONFIGS
    nums, cr = chrom_bins(CONFIGS["cfg2"][0], CONFIGS["cfg2"][1])
    N = int(sum(nums))
    rng = np.random.default_rng(1)
    clusters = [np.unique(rng.integers(1, N + 1, int(min(25, 2 + rng.geometric(0.3))))) for _ in range(3000)]
    members = np.concatenate(clusters).astype(np.int64)
    offsets = np.concatenate([[0], np.cumsum([len(c) for c in clusters])]).astype(np.int64)
    np.testing.assert_array_equal(adjacency_from_clusters(members, offsets, N).cpu().numpy(), FO.edgelist2adj(clusters, N))
    n_pix = 200_000
    key = np.unique(rng.integers(0, N, n_pix) * N + rng.integers(0, N, n_pix))
    b1, b2 = np.minimum(key // N, key % N), np.maximum(key // N, key % N)
    key = np.unique(b1 * N + b2)
    b1, b2 = key // N, key % N
    cnt = rng.gamma(2.0, 1.0, len(b1))
    cnt[::37] = np.nan
    cool2node = np.arange(1, N + 1, dtype=np.int64)
    cool2node[100:130] = 0                                          # bins of a chromosome outside chrom_list
    n2c = np.zeros(N + 1, dtype=np.int32)
    for c, (s, e) in enumerate(cr):
        n2c[int(s):int(e)] = c
    intra, inter = adjacency_from_pixels(b1, b2, cnt, cool2node, n2c, N)
    c2n = {i: int(v) for i, v in enumerate(cool2node) if v > 0}
    wi, we = FO.pixels2adj(b1, b2, cnt, c2n, {i: int(n2c[i]) for i in range(1, N + 1)}, N)
    np.testing.assert_array_equal(intra.cpu().numpy(), wi)
    np.testing.assert_array_equal(inter.cpu().numpy(), we)


@pytest.mark.parametrize("n", [2, 65, 250, 1111])
def test_corrcoef_matches_numpy(n):
    """main.py:572-577: np.corrcoef of the chromosome block (float64 inside numpy), NaN -> 0, float32."""
    from matcha_b200.features import corrcoef_features
    rng = np.random.default_rng(n)
    N = n + 7
    adj = (rng.poisson(1.5, (N, N)) * (rng.random((N, N)) < 0.3)).astype(np.float32)
    adj = adj + adj.T
    if n > 10:
        adj[5 + 3, :] = 0                                               # an unmappable bin: zero variance -> NaN -> 0
        adj[:, 5 + 3] = 0
        adj[5 + 9, 5:5 + n] = 2.0                                       # constant row inside the block
    cr = np.asarray([[6, 6 + n]])
    with np.errstate(invalid="ignore", divide="ignore"):
        want = np.corrcoef(adj[5:5 + n, 5:5 + n]).astype(np.float32)
    want[np.isnan(want)] = 0.0
    got = corrcoef_features(adj, cr)[0].cpu().numpy()
    np.testing.assert_allclose(got, want, rtol=0, atol=2e-7)


def test_zscore_positive_rows_matches_scipy_and_feeds_the_model(golden):
    """Modules.py:147-152 on the device against the host restatement (itself pinned to the reference's own table by
    tests/test_cpu_host.py::test_zscore_restatement_matches_reference), incl. all-zero rows, single-positive rows, rows of
    equal positives (0 / 0 -> NaN -> 0) and NaN entries."""
    from matcha_b200.features import zscore_positive_rows_
    from matcha_b200.hyper_sagnn import zscore_positive_rows_host
    rng = np.random.default_rng(0)
    M = (rng.gamma(1.0, 2.0, (300, 1000)) * (rng.random((300, 1000)) < 0.2)).astype(np.float32)
    M[3] = 0
    M[4] = 0; M[4, 17] = 2.5
    M[5] = 0; M[5, 10:20] = 1.25
    M[6, 5] = np.nan
    M[7] = -M[7]
    want = zscore_positive_rows_host(M.copy())
    got = zscore_positive_rows_(torch.from_numpy(M.copy()).cuda()).cpu().numpy()
    np.testing.assert_allclose(got, want, rtol=2e-6, atol=2e-6)
    assert got[3].max() == 0 and got[4, 17] == 0 and (got[5] == 0).all() and got[6, 5] == 0
    # strided view: a chromosome block of a larger matrix
    big = torch.from_numpy(M.copy()).cuda()
    zscore_positive_rows_(big[10:50, 100:400])
    want_blk = zscore_positive_rows_host(M[10:50, 100:400].copy())
    np.testing.assert_allclose(big[10:50, 100:400].cpu().numpy(), want_blk, rtol=2e-6, atol=2e-6)
    np.testing.assert_array_equal(big[:10].cpu().numpy()[~np.isnan(M[:10])], M[:10][~np.isnan(M[:10])])
