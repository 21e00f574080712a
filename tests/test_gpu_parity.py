"""GPU parity tests (run on the B200 box: pytest -m gpu).  Every test calls the CUDA path through the
C ABI (ctypes) and checks it against (a) golden vectors produced by the unmodified reference and/or
(b) the CPU oracle on the same seeded inputs.  Tolerances: fp32, rtol 1e-4 on scores/embeddings
(BASELINE.json north_star); sampled / indexed work bit-exact."""
import ctypes as C
import io
import json

import numpy as np
import pytest
import torch

from helpers import model_from_golden, oracle_from_model, step_seed, tc_gemm_rel_err
from oracle import hypersagnn_oracle as O
from oracle import sampler_oracle as SO

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def model(golden):
    return model_from_golden(golden)


def _lib():
    from matcha_b200 import _lib
    return _lib


# ------------------------------------------------------------------------------------------
# building block: fp32 contraction kernel
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("form,M,N,K", [(0, 130, 1536, 64), (0, 37, 250, 24), (0, 64, 64, 250), (1, 300, 64, 1536),
                                        (1, 65, 64, 133), (2, 1536, 64, 1000), (2, 64, 250, 777), (2, 133, 64, 4096)])
def test_gemm_simt_matches_torch(form, M, N, K):
    L = _lib()
    lib = L.load()
    g = torch.Generator(device="cuda").manual_seed(form * 1000 + M)
    if form == 0:
        A, B = torch.randn(M, K, device="cuda", generator=g), torch.randn(N, K, device="cuda", generator=g)
        ref = A.double() @ B.double().t()
    elif form == 1:
        A, B = torch.randn(M, K, device="cuda", generator=g), torch.randn(K, N, device="cuda", generator=g)
        ref = A.double() @ B.double()
    else:
        A, B = torch.randn(K, M, device="cuda", generator=g), torch.randn(K, N, device="cuda", generator=g)
        ref = A.double().t() @ B.double()
    bias = torch.randn(N, device="cuda", generator=g) if form == 0 else None
    if bias is not None:
        ref = ref + bias.double()
    Cm = torch.zeros(M, N, device="cuda")
    L.check(lib.matcha_gemm(form, 0, A.data_ptr(), B.data_ptr(), Cm.data_ptr(), L.ptr(bias), M, N, K, A.stride(0),
                            B.stride(0), N, 0, 0, L.stream_ptr()), "matcha_gemm")
    torch.cuda.synchronize()
    err = (Cm.double() - ref).abs().max().item() / ref.abs().max().item()
    assert err < 2e-6, err


@pytest.mark.parametrize("form,M,N,K", [(0, 130, 1536, 64), (0, 4099, 512, 64), (1, 300, 64, 1536), (1, 5000, 64, 128),
                                        (2, 1536, 64, 3000), (2, 512, 64, 20000)])
def test_gemm_tcgen05_matches_fp64(form, M, N, K):
    """tcgen05 path (bf16 hi/lo split x3, fp32 TMEM accumulator): ~2^-16 relative error per product."""
    assert tc_gemm_rel_err(form, M, N, K, impl=1, bias=(form == 0)) < 3e-5


@pytest.mark.parametrize("form,M,N,K", [(0, 4099, 3072, 128), (0, 1500, 250, 2491), (0, 1024, 64, 70), (1, 5000, 128, 3072),
                                        (1, 1100, 2491, 128), (1, 2048, 72, 64), (2, 3072, 128, 5000), (2, 128, 2491, 1500),
                                        (2, 200, 70, 1024)])
def test_general_tcgen05_gemm_matches_fp64(form, M, N, K):
    """General tcgen05 kernel (csrc/gemm_tcg.cu): the embed_dim-128 shapes of cfg5 and ragged M / N / K edges in all three
    forms, bias epilogue in form 0, split-K atomics in form 2."""
    assert tc_gemm_rel_err(form, M, N, K, impl=2, bias=(form == 0), v2=False) < 3e-5


def test_pipeline_agrees_between_simt_and_tcgen05(golden, model):
    L = _lib()
    lib = L.load()
    model.eval()
    x = torch.from_numpy(golden["x/L5"]).cuda().repeat(40, 1)        # 640 hyperedges -> 3200 tokens (>= 2048 for the TN path)
    try:
        lib.matcha_set_gemm_impl(0)
        with torch.no_grad():
            a = model(x)
        lib.matcha_set_gemm_impl(1)
        with torch.no_grad():
            b = model(x)
    finally:
        lib.matcha_set_gemm_impl(1)
    np.testing.assert_allclose(a.cpu().numpy(), b.cpu().numpy(), rtol=1e-4, atol=5e-5)
    np.testing.assert_allclose(b.cpu().numpy()[:16], golden["logits_eval/L5"], rtol=1e-4, atol=5e-5)


# ------------------------------------------------------------------------------------------
# Classifier.forward / get_node_embeddings vs the reference's own outputs
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("L", [2, 3, 4, 5])
def test_eval_logits_match_reference_golden(golden, model, L):
    model.eval()
    x = torch.from_numpy(golden[f"x/L{L}"]).cuda()
    with torch.no_grad():
        got = model(x)
    assert got.shape == (x.shape[0], 1)          # raw logits [B, 1] like Modules.py:309-318
    np.testing.assert_allclose(got.cpu().numpy(), golden[f"logits_eval/L{L}"], rtol=1e-4, atol=5e-5)


def test_recon_loss_matches_reference_golden(golden, model, monkeypatch):
    model.eval()
    for L in (2, 5):
        x = torch.from_numpy(golden[f"x/L{L}"]).cuda()
        for r in range(len(golden["nums"])):
            monkeypatch.setattr(np.random, "choice", lambda a, size=None, r=r: np.asarray([r]))
            with torch.no_grad():
                logits, rl = model(x, return_recon=True)
            np.testing.assert_allclose(rl.cpu().numpy(), golden[f"recon_eval/L{L}/r{r}"], rtol=1e-4)
            np.testing.assert_allclose(logits.cpu().numpy(), golden[f"logits_eval/L{L}"], rtol=1e-4, atol=5e-5)


def test_embeddings_match_reference_golden(golden, model):
    model.eval()
    N = golden["embeddings"].shape[0]
    ids = torch.arange(1, N + 1).view(-1, 1).cuda()
    with torch.no_grad():
        e = model.get_node_embeddings(ids)
    assert e.shape == (N, 1, 64)
    np.testing.assert_allclose(e[:, 0, :].cpu().numpy(), golden["embeddings"], rtol=1e-4, atol=2e-6)
    with torch.no_grad():       # id 0 -> zero row (Modules.py:178)
        z = model.get_node_embeddings(torch.zeros(3, 1, dtype=torch.long).cuda())
    assert float(z.abs().max()) == 0.0


def test_known_answer_facts(golden, model):
    model.eval()
    pw = golden["kat/pad_width"]
    with torch.no_grad():
        for i, x in enumerate([[3, 47, 90], [3, 47, 90, 0], [3, 47, 90, 0, 0]]):
            assert abs(model(torch.tensor([x]).cuda()).item() - pw[i]) < 1e-4      # padded width changes the score
        assert abs(model(torch.tensor([[90, 3, 47]]).cuda()).item() - golden["kat/permuted"][0]) < 1e-4


@pytest.mark.parametrize("L", [3, 5])
def test_gradients_match_reference_golden(golden, L, monkeypatch):
    """Train mode, dropout p = 0, loss = 1.0 * bce + 0.5 * recon, driven exactly like main.py:164-183
    (torch's BCE + loss.backward() through our autograd Function)."""
    model = model_from_golden(golden)
    model.train()
    for mod in model.modules():
        if isinstance(mod, torch.nn.Dropout):
            mod.p = 0.0
    r = int(golden[f"train_rchrom/L{L}"][0])
    monkeypatch.setattr(np.random, "choice", lambda a, size=None: np.asarray([r]))
    x = torch.from_numpy(golden[f"x/L{L}"]).cuda()
    y, w = torch.from_numpy(golden[f"y/L{L}"]).cuda(), torch.from_numpy(golden[f"w/L{L}"]).cuda()
    model.zero_grad(set_to_none=True)
    pred, rl = model(x, return_recon=True)
    bce = torch.nn.functional.binary_cross_entropy_with_logits(pred, y, weight=w)
    (bce * 1.0 + rl * 0.5).backward()
    np.testing.assert_allclose(pred.detach().cpu().numpy(), golden[f"train_logits/L{L}"], rtol=1e-4, atol=5e-5)
    np.testing.assert_allclose(rl.detach().cpu().numpy(), golden[f"train_recon/L{L}"], rtol=1e-4)
    meta = json.loads(str(golden["meta"]))
    live = set(meta["live_keys_by_L"][str(L)])
    got_live = set()
    for k, p in model.named_parameters():
        if p.grad is not None:
            got_live.add(k)
            ref = golden[f"grad/L{L}/{k}"]
            scale = float(np.abs(ref).max())
            err = float(np.abs(p.grad.cpu().numpy() - ref).max())
            assert err <= 3e-4 * scale + 2e-7, (k, err, scale)
    assert got_live == live          # same parameters get a gradient as in the reference (others stay None)


@pytest.mark.parametrize("L", [2, 4, 5])
def test_train_step_with_dropout_matches_oracle(golden, L, monkeypatch):
    """Dropout ON: the oracle regenerates the kernel's masks from the shared counter RNG, so logits and
    gradients must still agree."""
    model = model_from_golden(golden)
    model.train()
    eng = model._engine()
    r = L % 3
    monkeypatch.setattr(np.random, "choice", lambda a, size=None: np.asarray([r]))
    x = torch.from_numpy(golden[f"x/L{L}"]).cuda()
    y, w = torch.from_numpy(golden[f"y/L{L}"]).cuda(), torch.from_numpy(golden[f"w/L{L}"]).cuda()
    pred, rl = model(x, return_recon=True)
    (torch.nn.functional.binary_cross_entropy_with_logits(pred, y, weight=w) + 0.25 * rl.sum()).backward()
    om = oracle_from_model(model).to(torch.float64)
    out = O.loss_and_grads(om, x.cpu(), y.cpu().double(), w.cpu().double(), 1.0, 0.25, random_chrom=r, train=True,
                           seed=step_seed(eng))
    np.testing.assert_allclose(pred.detach().cpu().numpy(), out["logits"].numpy(), rtol=2e-4, atol=1e-4)
    np.testing.assert_allclose(rl.detach().cpu().numpy(), out["recon"].numpy(), rtol=1e-4)
    sd = dict(model.named_parameters())
    for k, gref in out["grads"].items():
        p = sd[k]
        gref = gref.numpy()
        if p.grad is None:
            assert np.abs(gref).max() == 0.0, k
            continue
        scale = float(np.abs(gref).max())
        err = float(np.abs(p.grad.cpu().numpy() - gref).max())
        assert err <= 5e-4 * scale + 2e-7, (k, err, scale)


# ------------------------------------------------------------------------------------------
# all-pairs scorer (k = 2 closed form) and tuple scorer
# ------------------------------------------------------------------------------------------
def test_pair_scorer_matches_reference_and_generic_path(golden, model):
    from matcha_b200.scorer import PairScorer, pair_count, pair_index_to_ij
    model.eval()
    ps = PairScorer(model)
    N = int(golden["chrom_range"][-1][1]) - 1
    # whole node range as one "chromosome": pair order of generate_pair_wise (denoise_contact.py:67-74)
    for md in (0, 2):
        total = pair_count(1, N + 1, md)
        ref_pairs = np.asarray([(i, j) for i in range(1, N + 1) for j in range(i + md, N + 1)], dtype=np.int64)
        assert total == len(ref_pairs)
        got = ps.score_range(1, N + 1, md).cpu().numpy()
        ii, jj = pair_index_to_ij(np.arange(total), 1, N + 1, md)
        assert (ii == ref_pairs[:, 0]).all() and (jj == ref_pairs[:, 1]).all()
        sel = np.arange(0, total, 7)
        keep = ref_pairs[sel][:, 0] != ref_pairs[sel][:, 1]        # (i, i) is not a 2-token hyperedge for the generic path
        with torch.no_grad():
            generic = model(torch.from_numpy(ref_pairs[sel][keep]).cuda()).cpu().numpy().reshape(-1)
        np.testing.assert_allclose(got[sel][keep], generic, rtol=1e-4, atol=5e-5)
        # sub-range (sharding) gives the same values
        b, e = total // 3, total // 3 + 1000
        np.testing.assert_array_equal(ps.score_range(1, N + 1, md, b, e).cpu().numpy(), got[b:e])
    # golden pairs scored by the unmodified reference
    pairs = golden["kat/pairs"]
    got = ps.score_range(1, N + 1, 0).cpu().numpy()
    n = N
    idx = np.asarray([(i - 1) * n - (i - 1) * (i - 2) // 2 + (j - i) for i, j in pairs])
    np.testing.assert_allclose(got[idx], golden["kat/pair_logits"].reshape(-1), rtol=1e-4, atol=5e-5)


def test_score_tuples_pads_per_batch(golden, model):
    from matcha_b200.scorer import score_tuples
    samples = [[3, 47], [5, 9, 90], [1, 2], [7, 50, 60, 99]]
    outs = score_tuples(model, samples, batch_size=2)
    model.eval()
    with torch.no_grad():
        a = model(torch.tensor([[3, 47, 0], [5, 9, 90]]).cuda())
        b = model(torch.tensor([[1, 2, 0, 0], [7, 50, 60, 99]]).cuda())
    np.testing.assert_allclose(outs[0][1].cpu().numpy(), a.cpu().numpy(), rtol=1e-6)
    np.testing.assert_allclose(outs[1][1].cpu().numpy(), b.cpu().numpy(), rtol=1e-6)


# ------------------------------------------------------------------------------------------
# hash set + sampler: bit-exact vs the exact-set oracle
# ------------------------------------------------------------------------------------------
def _toy_kmers(rng, cr, n, L=5):
    rows = np.zeros((n, L), dtype=np.int64)
    for i in range(n):
        k = int(rng.integers(2, L + 1))
        c = int(rng.integers(0, len(cr)))
        if rng.random() < 0.9:
            ids = rng.choice(np.arange(cr[c][0], cr[c][1]), size=k, replace=False)
        else:
            ids = rng.choice(np.arange(1, cr[-1][1]), size=k, replace=False)
        rows[i, :k] = np.sort(ids)
    return np.unique(rows, axis=0)


def test_hashset_membership_exact(golden):
    from matcha_b200.sampler import KmerHashSet
    rng = np.random.default_rng(3)
    cr = golden["chrom_range"]
    kmers = _toy_kmers(rng, cr, 20000)
    hs = KmerHashSet(len(kmers), width=5).insert(kmers)
    assert not hs.overflowed()
    s = SO.build_set(kmers)
    probe = np.concatenate([kmers[::3], _toy_kmers(rng, cr, 5000)])
    got = hs.contains(probe).cpu().numpy()
    want = np.asarray([SO.kmer_key(r) in s for r in probe])
    assert (got == want).all()
    assert got[: len(kmers[::3])].all()


@pytest.mark.parametrize("min_dis", [0, 1])
def test_negative_sampler_bit_exact_vs_oracle(golden, min_dis):
    from matcha_b200.sampler import KmerHashSet, NegativeSampler
    rng = np.random.default_rng(11)
    cr = golden["chrom_range"]
    kmers = _toy_kmers(rng, cr, 30000)
    hs = KmerHashSet(len(kmers), width=5).insert(kmers)
    pos = kmers[rng.choice(len(kmers), 96, replace=False)]
    smp = NegativeSampler(hs, cr, min_dis=min_dis, neg_num=3, seed=2, max_rounds=64)
    rounds = torch.zeros(96 * 3, dtype=torch.int32, device="cuda")
    neg, valid = smp.sample(torch.from_numpy(pos).cuda(), rounds=rounds, step=5)
    want_neg, want_valid, want_rounds = SO.sample_negatives(pos, SO.build_set(kmers), cr, 3, min_dis, seed=2, step=5)
    assert (neg.cpu().numpy() == want_neg).all()
    assert (valid.cpu().numpy() == want_valid).all()
    assert (rounds.cpu().numpy() == want_rounds).all()
    # domain properties: sorted, unique, never a positive, same chromosomes as the source positive
    n, v = neg.cpu().numpy(), valid.cpu().numpy()
    s = SO.build_set(kmers)
    assert v.mean() > 0.8
    for g, row in enumerate(n):
        if not v[g]:          # every same-chromosome corruption was a positive: the row is the positive, flagged
            assert (row == pos[g // 3]).all()
            continue
        live = row[row != 0]
        assert (np.diff(live) > min_dis).all()
        assert SO.kmer_key(row) not in s


# ------------------------------------------------------------------------------------------
# optimizer, persistence, fused trainer
# ------------------------------------------------------------------------------------------
def test_flat_adamw_matches_torch(golden):
    from matcha_b200.engine import FlatAdamW
    model = model_from_golden(golden)
    eng = model._engine()
    eng.ensure_bound()
    live = eng.live_parameters()
    ref_params = [p.detach().clone().requires_grad_(True) for p in live]
    topt = torch.optim.AdamW(ref_params, lr=1e-3)
    opt = FlatAdamW(eng)
    g = torch.Generator(device="cuda").manual_seed(0)
    for it in range(3):
        eng.gflat.normal_(generator=g)
        eng.active.fill_(1)
        if it == 1:
            eng.active[0] = 0                 # chromosome 0 absent: its encoder weights must not move
        n_always = len(eng.layout) - len(eng.segments)
        for i, ((p, o), rp) in enumerate(zip(eng.layout, ref_params)):
            rp.grad = eng.grad_view(p, o).clone()
            if it == 1 and i >= n_always and eng.segments[i - n_always][2] == 0:
                rp.grad = None                # what autograd leaves when the chromosome is absent
        topt.step()
        opt.step()
    for (p, o), rp in zip(eng.layout, ref_params):
        np.testing.assert_allclose(p.detach().cpu().numpy(), rp.detach().cpu().numpy(), rtol=2e-5, atol=2e-7)


def test_whole_module_pickle_roundtrip(golden, model):
    model.eval()
    x = torch.from_numpy(golden["x/L4"]).cuda()
    with torch.no_grad():
        a = model(x)
    buf = io.BytesIO()
    torch.save(model, buf)                    # main.py:322,685 (model2load)
    buf.seek(0)
    m2 = torch.load(buf, weights_only=False)  # denoise_contact.py:99 (needs weights_only=False on torch >= 2.6)
    m2.eval()
    with torch.no_grad():
        b = m2(x)
    assert torch.equal(a, b)
    assert str(m2.layer_norm1.weight.device).startswith("cuda")      # denoise_contact.py:101-104 reads this


def test_fused_trainer_step_matches_oracle(golden):
    """One Trainer.step (sampler -> fwd -> bce -> bwd -> AdamW) against the oracle chained the same way."""
    from matcha_b200.sampler import KmerHashSet, NegativeSampler
    from matcha_b200.trainer import Trainer
    rng = np.random.default_rng(21)
    cr = golden["chrom_range"]
    kmers = _toy_kmers(rng, cr, 20000)
    hs = KmerHashSet(len(kmers), width=5).insert(kmers)
    model = model_from_golden(golden)
    om = oracle_from_model(model).to(torch.float64)
    smp = NegativeSampler(hs, cr, neg_num=3, seed=4)
    tr = Trainer(model, smp, alpha=1.0, beta=0.3, seed=9)
    pos = kmers[rng.choice(len(kmers), 32, replace=False)]
    pw = rng.uniform(0.5, 3.0, 32).astype(np.float32)
    before = {k: v.detach().cpu().double().clone() for k, v in model.named_parameters()}
    tr.step(torch.from_numpy(pos).cuda(), torch.from_numpy(pw).cuda())
    torch.cuda.synchronize()
    eng = tr.e
    neg, valid, _ = SO.sample_negatives(pos, SO.build_set(kmers), cr, 3, 0, seed=4, step=0)
    x = torch.from_numpy(np.concatenate([pos, neg]))
    y = torch.cat([torch.ones(32, 1), torch.zeros(96, 1)]).double()
    w = torch.cat([torch.from_numpy(pw).view(-1, 1), torch.from_numpy(valid.astype(np.float32)).view(-1, 1)]).double()
    rchrom = int(np.random.RandomState(9).randint(0, len(cr)))
    out = O.loss_and_grads(om, x, y, w, 1.0, 0.3, random_chrom=rchrom, train=True, seed=step_seed(eng))
    losses = tr.loss_out.cpu().numpy()
    np.testing.assert_allclose(losses[0], out["bce"].item(), rtol=2e-4)
    np.testing.assert_allclose(losses[1], out["recon"].item(), rtol=2e-4)
    np.testing.assert_allclose(losses[2], out["loss"].item(), rtol=2e-4)
    after = dict(model.named_parameters())
    for k, gref in out["grads"].items():
        if float(gref.abs().max()) == 0.0:
            continue
        p0 = before[k]
        want, _, _ = O.adamw_step(p0, gref, torch.zeros_like(p0), torch.zeros_like(p0), 1)
        got = after[k].detach().cpu().double()
        # the first AdamW step moves a coordinate by ~lr * g / (|g| + eps): compare the update where the
        # gradient is well above its own rounding error (near-zero gradients make the sign ill-conditioned)
        mask = gref.abs() > 1e-3 * gref.abs().max()
        upd_err = float((((got - p0) - (want - p0)).abs() * mask).max())
        assert upd_err < 2e-5, (k, upd_err)


def test_tile_path_gradients_match_simt_path(golden, monkeypatch):
    """Training step on 1665 tokens (not a multiple of the 128-token tile): the tcgen05 tile pipeline (pre-split
    operands, bulk copies, TMEM accumulators, ones-column bias gradient) vs the all-SIMT fp32 pipeline."""
    L = _lib()
    lib = L.load()
    monkeypatch.setattr(np.random, "choice", lambda a, size=None: np.asarray([1]))
    rng = np.random.default_rng(0)
    N = int(golden["chrom_range"][-1][1]) - 1
    x = np.zeros((333, 5), dtype=np.int64)
    for b in range(333):
        k = int(rng.integers(2, 6))
        x[b, :k] = np.sort(rng.choice(np.arange(1, N + 1), size=k, replace=False))
    x = torch.from_numpy(x).cuda()
    y = torch.from_numpy((rng.random((333, 1)) < 0.3).astype("float32")).cuda()
    w = torch.from_numpy(rng.uniform(0.5, 3, (333, 1)).astype("float32")).cuda()
    grads = {}
    try:
        for impl in (0, 1):
            lib.matcha_set_gemm_impl(impl)
            model = model_from_golden(golden)
            model.train()
            eng = model._engine()
            eng.seed_base = 77
            pred, rl = model(x, return_recon=True)
            (torch.nn.functional.binary_cross_entropy_with_logits(pred, y, weight=w) + 0.1 * rl.sum()).backward()
            grads[impl] = ({k: p.grad.detach().cpu().numpy().copy() for k, p in model.named_parameters() if p.grad is not None},
                           pred.detach().cpu().numpy())
    finally:
        lib.matcha_set_gemm_impl(1)
    np.testing.assert_allclose(grads[1][1], grads[0][1], rtol=2e-4, atol=1e-4)
    assert grads[0][0].keys() == grads[1][0].keys()
    for k, g0 in grads[0][0].items():
        g1 = grads[1][0][k]
        scale = float(np.abs(g0).max())
        assert float(np.abs(g1 - g0).max()) <= 5e-4 * scale + 1e-7, (k, float(np.abs(g1 - g0).max()), scale)


# ------------------------------------------------------------------------------------------
# fused hyperedge-tile kernels (attn_fused.cu) vs the decomposed pipeline and the reference golden
# ------------------------------------------------------------------------------------------
def _random_hyperedges(golden, B, L, seed, kmin=2):
    rng = np.random.default_rng(seed)
    N = int(golden["chrom_range"][-1][1]) - 1
    x = np.zeros((B, L), dtype=np.int64)
    for b in range(B):
        k = int(rng.integers(kmin, L + 1))
        x[b, :k] = np.sort(rng.choice(np.arange(1, N + 1), size=k, replace=False))
    return x


@pytest.mark.parametrize("L,B", [(2, 700), (3, 411), (4, 777), (5, 333), (5, 4099), (6, 300)])
def test_fused_attention_forward_matches_decomposed(golden, model, L, B):
    """Eval logits through the fused tcgen05 attention kernel (QKG in TMEM, shuffle softmax) vs the decomposed
    pipeline on the same inputs; B chosen so the last hyperedge-aligned tile is ragged."""
    lib = _lib().load()
    model.eval()
    x = torch.from_numpy(_random_hyperedges(golden, B, L, seed=L * 1000 + B)).cuda()
    if f"x/L{L}" in golden.files:
        x[:16] = torch.from_numpy(golden[f"x/L{L}"]).cuda()
    try:
        lib.matcha_set_fused(0)
        with torch.no_grad():
            a = model(x).cpu().numpy()
        lib.matcha_set_fused(1)
        with torch.no_grad():
            b = model(x).cpu().numpy()
    finally:
        lib.matcha_set_fused(1)
    np.testing.assert_allclose(b, a, rtol=1e-4, atol=5e-5)
    if f"logits_eval/L{L}" in golden.files:
        np.testing.assert_allclose(b[:16], golden[f"logits_eval/L{L}"], rtol=1e-4, atol=5e-5)


@pytest.mark.parametrize("L,B", [(5, 333), (4, 300), (3, 411), (2, 640), (5, 2100)])
def test_fused_attention_training_matches_decomposed(golden, monkeypatch, L, B):
    """Training step (dropout ON, same counter-RNG masks) through the fused forward + fused backward kernels vs the
    decomposed pipeline: logits, recon loss and every gradient."""
    lib = _lib().load()
    monkeypatch.setattr(np.random, "choice", lambda a, size=None: np.asarray([1]))
    rng = np.random.default_rng(B)
    x = torch.from_numpy(_random_hyperedges(golden, B, L, seed=7 * L + B)).cuda()
    y = torch.from_numpy((rng.random((B, 1)) < 0.3).astype("float32")).cuda()
    w = torch.from_numpy(rng.uniform(0.5, 3, (B, 1)).astype("float32")).cuda()
    res = {}
    try:
        for fused in (0, 1):
            lib.matcha_set_fused(fused)
            model = model_from_golden(golden)
            model.train()
            eng = model._engine()
            eng.seed_base = 91
            pred, rl = model(x, return_recon=True)
            (torch.nn.functional.binary_cross_entropy_with_logits(pred, y, weight=w) + 0.1 * rl.sum()).backward()
            res[fused] = ({k: p.grad.detach().cpu().numpy().copy() for k, p in model.named_parameters() if p.grad is not None},
                          pred.detach().cpu().numpy(), float(rl.sum()))
    finally:
        lib.matcha_set_fused(1)
    np.testing.assert_allclose(res[1][1], res[0][1], rtol=1e-4, atol=5e-5)
    assert abs(res[1][2] - res[0][2]) <= 1e-5 * abs(res[0][2]) + 1e-6
    assert res[0][0].keys() == res[1][0].keys()
    for k, g0 in res[0][0].items():
        g1 = res[1][0][k]
        scale = float(np.abs(g0).max())
        assert float(np.abs(g1 - g0).max()) <= 5e-4 * scale + 1e-7, (k, float(np.abs(g1 - g0).max()), scale)


@pytest.mark.parametrize("L,B", [(5, 333), (4, 300), (3, 411), (2, 640), (5, 2100)])
def test_chain_kernels_match_simt_layers(golden, monkeypatch, L, B):
    """The tcgen05 row-chain kernels (attribute mix + next_w + LayerNorm, pff_n1 + scorer and their backward
    counterparts) vs the SIMT fp32 layers, inside the fused pipeline: eval logits, train logits, every gradient."""
    lib = _lib().load()
    monkeypatch.setattr(np.random, "choice", lambda a, size=None: np.asarray([1]))
    rng = np.random.default_rng(B + 1)
    x = torch.from_numpy(_random_hyperedges(golden, B, L, seed=11 * L + B)).cuda()
    y = torch.from_numpy((rng.random((B, 1)) < 0.3).astype("float32")).cuda()
    w = torch.from_numpy(rng.uniform(0.5, 3, (B, 1)).astype("float32")).cuda()
    res = {}
    try:
        for chain in (0, 1):
            lib.matcha_set_chain(chain)
            model = model_from_golden(golden)
            model.eval()
            with torch.no_grad():
                ev = model(x).cpu().numpy()
            model.train()
            eng = model._engine()
            eng.seed_base = 123
            eng.tape_id = 0
            pred, rl = model(x, return_recon=True)
            (torch.nn.functional.binary_cross_entropy_with_logits(pred, y, weight=w) + 0.1 * rl.sum()).backward()
            res[chain] = ({k: p.grad.detach().cpu().numpy().copy() for k, p in model.named_parameters() if p.grad is not None},
                          pred.detach().cpu().numpy(), ev)
    finally:
        lib.matcha_set_chain(1)
    np.testing.assert_allclose(res[1][2], res[0][2], rtol=1e-4, atol=5e-5)
    np.testing.assert_allclose(res[1][1], res[0][1], rtol=1e-4, atol=5e-5)
    assert res[0][0].keys() == res[1][0].keys()
    for k, g0 in res[0][0].items():
        g1 = res[1][0][k]
        scale = float(np.abs(g0).max())
        assert float(np.abs(g1 - g0).max()) <= 5e-4 * scale + 1e-7, (k, float(np.abs(g1 - g0).max()), scale)


@pytest.mark.parametrize("n,min_dis", [(1000, 0), (777, 3), (300, 130), (2500, 1)])
def test_pair_scorer_tensor_core_matches_simt(n, min_dis):
    """tcgen05 all-pairs scorer (u_i + u_j - PA_i.PB_j + b) vs the fp32 FMA form on random tables, incl. a ragged last
    block, min_distance inside and beyond one block, and an interior sub-range (the multi-GPU shard)."""
    L = _lib()
    lib = L.load()
    g = torch.Generator(device="cuda").manual_seed(n)
    lo = 5
    D = torch.randn(lo + n, 64, device="cuda", generator=g)
    S = torch.randn(lo + n, 64, device="cuda", generator=g)
    w = (torch.rand(64, device="cuda", generator=g) - 0.5) * 0.25
    b = torch.full((1,), 0.1, device="cuda")
    total = int(lib.matcha_pair_count(lo, lo + n, min_dis))
    nbytes = int(lib.matcha_pair_tc_workspace_bytes(lo, lo + n))
    ws = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    L.check(lib.matcha_pair_tc_prepare(L.ptr(D), L.ptr(S), L.ptr(w), L.ptr(b), 64, lo, lo + n, L.ptr(ws), nbytes, L.stream_ptr()), "prep")
    for (pb, pe) in [(0, total), (total // 3, 2 * total // 3 + 1)]:
        for sig in (0, 1):
            ref = torch.full((pe - pb,), -7.0, device="cuda")
            out = torch.full((pe - pb,), -9.0, device="cuda")
            L.check(lib.matcha_pair_score_range(L.ptr(D), L.ptr(S), L.ptr(w), L.ptr(b), 64, lo, lo + n, min_dis, pb, pe, sig,
                                                L.ptr(ref), L.stream_ptr()), "simt")
            L.check(lib.matcha_pair_tc_score_range(L.ptr(ws), lo, lo + n, min_dis, pb, pe, sig, L.ptr(out),
                                                   L.stream_ptr()), "tc")
            torch.cuda.synchronize()
            np.testing.assert_allclose(out.cpu().numpy(), ref.cpu().numpy(), rtol=1e-4, atol=5e-5)


def test_csr_encoder_matches_dense_rows(golden, monkeypatch):
    """SparseEmbedding(sparse=True) (Modules.py:58-65): the CSR SpMM encoder kernels (forward, weight gradient, feature
    dropout on the nonzeros) vs the dense-row path on the SAME thresholded feature matrices: embeddings, eval logits, and
    every gradient of a training step with dropout on."""
    import scipy.sparse as sp
    monkeypatch.setattr(np.random, "choice", lambda a, size=None: np.asarray([1]))
    nchrom = len(golden["nums"])
    dense = []
    for c in range(nchrom):
        f = np.array(golden[f"feat/{c}"], dtype=np.float32)
        f[np.abs(f) < 0.35] = 0.0                      # genuinely sparse rows (some of them empty)
        dense.append(f)
    B, L = 400, 5
    rng = np.random.default_rng(3)
    x = torch.from_numpy(_random_hyperedges(golden, B, L, seed=99)).cuda()
    y = torch.from_numpy((rng.random((B, 1)) < 0.3).astype("float32")).cuda()
    w = torch.from_numpy(rng.uniform(0.5, 3, (B, 1)).astype("float32")).cuda()
    res = {}
    for kind in ("dense", "csr"):
        feats = dense if kind == "dense" else [sp.csr_matrix(f) for f in dense]
        model = model_from_golden(golden, feats=feats, sparse=(kind == "csr"))
        model.eval()
        N = int(golden["chrom_range"][-1][1]) - 1
        with torch.no_grad():
            emb = model.get_node_embeddings(torch.arange(1, N + 1, device="cuda").view(-1, 1)).cpu().numpy()
            ev = model(x).cpu().numpy()
        model.train()
        eng = model._engine()
        eng.seed_base = 55
        eng.tape_id = 0
        pred, rl = model(x, return_recon=True)
        (torch.nn.functional.binary_cross_entropy_with_logits(pred, y, weight=w) + 0.1 * rl.sum()).backward()
        res[kind] = (emb, ev, pred.detach().cpu().numpy(),
                     {k: p.grad.detach().cpu().numpy().copy() for k, p in model.named_parameters() if p.grad is not None})
    np.testing.assert_allclose(res["csr"][0], res["dense"][0], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(res["csr"][1], res["dense"][1], rtol=1e-4, atol=5e-5)
    np.testing.assert_allclose(res["csr"][2], res["dense"][2], rtol=1e-4, atol=5e-5)
    assert res["csr"][3].keys() == res["dense"][3].keys()
    for k, g0 in res["dense"][3].items():
        g1 = res["csr"][3][k]
        scale = float(np.abs(g0).max())
        assert float(np.abs(g1 - g0).max()) <= 5e-4 * scale + 1e-7, (k, float(np.abs(g1 - g0).max()), scale)


# ------------------------------------------------------------------------------------------
# embed_dim 128 (BASELINE.json configs[4]): fp32 SIMT path, same bar as embed_dim 64
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("L", [3, 5])
def test_d128_eval_matches_reference_golden(golden128, L, monkeypatch):
    model = model_from_golden(golden128, d=128)
    model.eval()
    x = torch.from_numpy(golden128[f"x/L{L}"]).cuda()
    with torch.no_grad():
        got = model(x)
    np.testing.assert_allclose(got.cpu().numpy(), golden128[f"logits_eval/L{L}"], rtol=1e-4, atol=5e-5)
    for r in range(len(golden128["nums"])):
        monkeypatch.setattr(np.random, "choice", lambda a, size=None, r=r: np.asarray([r]))
        with torch.no_grad():
            _, rl = model(x, return_recon=True)
        np.testing.assert_allclose(rl.cpu().numpy(), golden128[f"recon_eval/L{L}/r{r}"], rtol=1e-4)
    N = golden128["embeddings"].shape[0]
    with torch.no_grad():
        e = model.get_node_embeddings(torch.arange(1, N + 1).view(-1, 1).cuda())
    assert e.shape == (N, 1, 128)
    np.testing.assert_allclose(e[:, 0, :].cpu().numpy(), golden128["embeddings"], rtol=1e-4, atol=2e-6)


def test_d128_gradients_match_reference_golden(golden128, monkeypatch):
    L = 5
    model = model_from_golden(golden128, d=128)
    model.train()
    for mod in model.modules():
        if isinstance(mod, torch.nn.Dropout):
            mod.p = 0.0
    r = int(golden128[f"train_rchrom/L{L}"][0])
    monkeypatch.setattr(np.random, "choice", lambda a, size=None: np.asarray([r]))
    x = torch.from_numpy(golden128[f"x/L{L}"]).cuda()
    y, w = torch.from_numpy(golden128[f"y/L{L}"]).cuda(), torch.from_numpy(golden128[f"w/L{L}"]).cuda()
    model.zero_grad(set_to_none=True)
    pred, rl = model(x, return_recon=True)
    bce = torch.nn.functional.binary_cross_entropy_with_logits(pred, y, weight=w)
    (bce * 1.0 + rl * 0.5).backward()
    np.testing.assert_allclose(pred.detach().cpu().numpy(), golden128[f"train_logits/L{L}"], rtol=1e-4, atol=5e-5)
    np.testing.assert_allclose(rl.detach().cpu().numpy(), golden128[f"train_recon/L{L}"], rtol=1e-4)
    meta = json.loads(str(golden128["meta"]))
    live = set(meta["live_keys_by_L"][str(L)])
    got_live = set()
    for k, p in model.named_parameters():
        if p.grad is None:
            continue
        got_live.add(k)
        g = p.grad.cpu().numpy()
        if f"grad/L{L}/{k}" in golden128.files:
            ref = golden128[f"grad/L{L}/{k}"]
            scale = float(np.abs(ref).max())
            err = float(np.abs(g - ref).max())
            assert err <= 3e-4 * scale + 2e-7, (k, err, scale)
        else:       # large tensors: Frobenius norm only in the fixture
            ref = float(golden128[f"gradnorm/L{L}/{k}"][0])
            assert abs(float(np.linalg.norm(g.astype(np.float64))) - ref) <= 3e-4 * ref + 1e-9, k
    assert got_live == live


@pytest.mark.parametrize("L,B", [(3, 16), (5, 600)])
def test_d128_train_step_with_dropout_matches_oracle(golden128, L, B, monkeypatch):
    """Dropout ON at embed_dim 128, every gradient element against the oracle's fp64 autograd (B = 600 exercises the
    multi-block grids)."""
    model = model_from_golden(golden128, d=128)
    model.train()
    eng = model._engine()
    r = L % 3
    monkeypatch.setattr(np.random, "choice", lambda a, size=None: np.asarray([r]))
    rng = np.random.default_rng(11)
    N = int(golden128["embeddings"].shape[0])
    xs = np.zeros((B, L), dtype=np.int64)
    for b in range(B):
        k = L if b % 2 == 0 else int(rng.integers(2, L + 1))
        xs[b, :k] = np.sort(rng.choice(np.arange(1, N + 1), size=k, replace=False))
    x = torch.from_numpy(xs).cuda()
    y = torch.from_numpy((rng.random((B, 1)) < 0.4).astype("float32")).cuda()
    w = torch.from_numpy(rng.uniform(0.5, 3.0, size=(B, 1)).astype("float32")).cuda()
    pred, rl = model(x, return_recon=True)
    (torch.nn.functional.binary_cross_entropy_with_logits(pred, y, weight=w) + 0.25 * rl.sum()).backward()
    om = oracle_from_model(model).to(torch.float64)
    out = O.loss_and_grads(om, x.cpu(), y.cpu().double(), w.cpu().double(), 1.0, 0.25, random_chrom=r, train=True,
                           seed=step_seed(eng))
    np.testing.assert_allclose(pred.detach().cpu().numpy(), out["logits"].numpy(), rtol=2e-4, atol=1e-4)
    np.testing.assert_allclose(rl.detach().cpu().numpy(), out["recon"].numpy(), rtol=1e-4)
    sd = dict(model.named_parameters())
    for k, gref in out["grads"].items():
        p = sd[k]
        gref = gref.numpy()
        if p.grad is None:
            assert np.abs(gref).max() == 0.0, k
            continue
        scale = float(np.abs(gref).max())
        err = float(np.abs(p.grad.cpu().numpy() - gref).max())
        assert err <= 5e-4 * scale + 2e-7, (k, err, scale)


def test_d128_general_tcgen05_step_matches_simt(golden128, monkeypatch):
    """embed_dim 128 at 10,240 tokens: one training step (dropout ON, reconstruction head on) with every contraction on the
    general tcgen05 kernel against the same step on the fp32 SIMT kernel -- logits, recon loss, every gradient."""
    lib = _lib().load()
    L, B = 5, 2048
    rng = np.random.default_rng(5)
    N = int(golden128["embeddings"].shape[0])
    xs = np.zeros((B, L), dtype=np.int64)
    for b in range(B):
        k = L if b % 2 == 0 else int(rng.integers(2, L + 1))
        xs[b, :k] = np.sort(rng.choice(np.arange(1, N + 1), size=k, replace=False))
    x = torch.from_numpy(xs).cuda()
    y = torch.from_numpy((rng.random((B, 1)) < 0.4).astype("float32")).cuda()
    w = torch.from_numpy(rng.uniform(0.5, 3.0, size=(B, 1)).astype("float32")).cuda()
    monkeypatch.setattr(np.random, "choice", lambda a, size=None: np.asarray([1]))
    res = {}
    try:
        for on in (0, 1):
            lib.matcha_set_gemm_tcg(on)
            model = model_from_golden(golden128, d=128)
            model.train()
            model._engine().seed_base = 23
            pred, rl = model(x, return_recon=True)
            (torch.nn.functional.binary_cross_entropy_with_logits(pred, y, weight=w) + 0.25 * rl.sum()).backward()
            res[on] = (pred.detach().cpu().numpy(), float(rl.detach().sum()),
                       {k: p.grad.detach().cpu().numpy().copy() for k, p in model.named_parameters() if p.grad is not None})
    finally:
        lib.matcha_set_gemm_tcg(1)
    np.testing.assert_allclose(res[1][0], res[0][0], rtol=2e-4, atol=1e-4)
    assert abs(res[1][1] - res[0][1]) <= 1e-4 * abs(res[0][1])
    assert res[0][2].keys() == res[1][2].keys()
    for k, g0 in res[0][2].items():
        scale = float(np.abs(g0).max())
        err = float(np.abs(res[1][2][k] - g0).max())
        assert err <= 3e-4 * scale + 2e-7, (k, err, scale)


# ------------------------------------------------------------------------------------------
# fused tensor-core reconstruction head vs the SIMT launches (chromosomes wider than one 128-column block)
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("L,B,rchrom", [(5, 700, 0), (3, 450, 1), (4, 1031, 0)])
def test_recon_head_tensor_core_matches_simt(monkeypatch, L, B, rchrom):
    from matcha_b200.synthetic import build_model, make_dataset
    lib = _lib().load()
    ds = make_dataset("cfg1", kmers_per_size=3000, seed=3)            # chr1 + chr2 at 1 Mb: 250 and 244 bins -> two column blocks
    N = int(ds["chrom_range"][-1][1]) - 1
    monkeypatch.setattr(np.random, "choice", lambda a, size=None: np.asarray([rchrom]))
    rng = np.random.default_rng(100 + B)
    xs = np.zeros((B, L), dtype=np.int64)
    for b in range(B):
        k = L if b % 3 else int(rng.integers(2, L + 1))
        xs[b, :k] = np.sort(rng.choice(np.arange(1, N + 1), size=k, replace=False))
    x = torch.from_numpy(xs).cuda()
    y = torch.from_numpy((rng.random((B, 1)) < 0.3).astype("float32")).cuda()
    w = torch.from_numpy(rng.uniform(0.5, 3, (B, 1)).astype("float32")).cuda()
    res = {}
    try:
        for tc in (0, 1):
            lib.matcha_set_recon_tc(tc)
            model = build_model(ds, seed=1)
            model.train()
            model._engine().seed_base = 17
            pred, rl = model(x, return_recon=True)
            (torch.nn.functional.binary_cross_entropy_with_logits(pred, y, weight=w) + 0.7 * rl.sum()).backward()
            res[tc] = ({k: p.grad.detach().cpu().numpy().copy() for k, p in model.named_parameters() if p.grad is not None},
                       pred.detach().cpu().numpy(), float(rl.sum()))
            model.eval()
            with torch.no_grad():
                res[tc] += (float(model(x, return_recon=True)[1].sum()),)
    finally:
        lib.matcha_set_recon_tc(1)
    assert res[0][2] > 0
    assert abs(res[1][2] - res[0][2]) <= 2e-5 * abs(res[0][2])          # training-mode recon loss
    assert abs(res[1][3] - res[0][3]) <= 2e-5 * abs(res[0][3])          # eval-mode recon loss
    np.testing.assert_allclose(res[1][1], res[0][1], rtol=1e-5, atol=1e-6)
    assert res[0][0].keys() == res[1][0].keys()
    for k, g0 in res[0][0].items():
        g1 = res[1][0][k]
        scale = float(np.abs(g0).max())
        assert float(np.abs(g1 - g0).max()) <= 3e-4 * scale + 1e-7, (k, float(np.abs(g1 - g0).max()), scale)


# ------------------------------------------------------------------------------------------
# tensor-core node encoder forward vs the grouped SIMT launches (chromosomes of 250 / 244 bins: 4 weight chunks, ragged)
# and, for the backward kernel's column groups, chromosomes wider than the 384 bins one TMEM-resident group holds
# (chr21 / chr22 at 50 kb: 936 / 1018 bins = 15 / 16 chunks = 3 groups each, the last one ragged)
# ------------------------------------------------------------------------------------------
WIDE = (["chr21", "chr22"], 50_000, 64)


@pytest.mark.parametrize("cfg,L,B", [("cfg1", 5, 900), (WIDE, 5, 700), (WIDE, 3, 1100)])
def test_pipelined_kernels_match_unit_kernels(monkeypatch, cfg, L, B):
    """The pipelined reconstruction head / encoder forward / encoder backward (warp-specialised, cp.async / bulk-copy rings)
    against the unit kernels they replace: one training step with dropout on -- same arithmetic, different schedule, so
    logits, recon loss and every gradient agree to the order of the atomic accumulations."""
    from matcha_b200.synthetic import build_model, make_dataset
    lib = _lib().load()
    ds = make_dataset(cfg, kmers_per_size=3000, seed=3)
    N = int(ds["chrom_range"][-1][1]) - 1
    monkeypatch.setattr(np.random, "choice", lambda a, size=None: np.asarray([0]))
    rng = np.random.default_rng(300 + B)
    xs = np.zeros((B, L), dtype=np.int64)
    for b in range(B):
        k = L if b % 3 else int(rng.integers(2, L + 1))
        xs[b, :k] = np.sort(rng.choice(np.arange(1, N + 1), size=k, replace=False))
    x = torch.from_numpy(xs).cuda()
    y = torch.from_numpy((rng.random((B, 1)) < 0.3).astype("float32")).cuda()
    w = torch.from_numpy(rng.uniform(0.5, 3, (B, 1)).astype("float32")).cuda()
    res = {}
    try:
        for pipe in (0, 1):
            lib.matcha_set_recon_pipe(pipe)
            lib.matcha_set_enc_pipe(pipe, pipe)
            model = build_model(ds, seed=1)
            model._engine().seed_base = 29
            model.train()
            pred, rl = model(x, return_recon=True)
            (torch.nn.functional.binary_cross_entropy_with_logits(pred, y, weight=w) + 0.7 * rl.sum()).backward()
            res[pipe] = ({k: p.grad.detach().cpu().numpy().copy() for k, p in model.named_parameters() if p.grad is not None},
                         pred.detach().cpu().numpy(), float(rl.detach().sum()))
    finally:
        lib.matcha_set_recon_pipe(1)
        lib.matcha_set_enc_pipe(2, 1)
    np.testing.assert_allclose(res[1][1], res[0][1], rtol=2e-5, atol=2e-6)
    assert abs(res[1][2] - res[0][2]) <= 2e-5 * abs(res[0][2])
    assert res[0][0].keys() == res[1][0].keys()
    for k, g0 in res[0][0].items():
        scale = float(np.abs(g0).max())
        err = float(np.abs(res[1][0][k] - g0).max())
        assert err <= 5e-5 * scale + 1e-7, (k, err, scale)


@pytest.mark.parametrize("cfg,L,B,train", [("cfg1", 5, 700, True), ("cfg1", 3, 450, False), ("cfg1", 4, 1031, True),
                                           (WIDE, 5, 700, True), (WIDE, 3, 900, True)])
def test_encoder_tensor_core_matches_simt(monkeypatch, cfg, L, B, train):
    from matcha_b200.synthetic import build_model, make_dataset
    lib = _lib().load()
    ds = make_dataset(cfg, kmers_per_size=3000, seed=3)
    N = int(ds["chrom_range"][-1][1]) - 1
    monkeypatch.setattr(np.random, "choice", lambda a, size=None: np.asarray([1]))
    rng = np.random.default_rng(200 + B)
    xs = np.zeros((B, L), dtype=np.int64)
    for b in range(B):
        k = L if b % 3 else int(rng.integers(2, L + 1))
        xs[b, :k] = np.sort(rng.choice(np.arange(1, N + 1), size=k, replace=False))
    x = torch.from_numpy(xs).cuda()
    y = torch.from_numpy((rng.random((B, 1)) < 0.3).astype("float32")).cuda()
    w = torch.from_numpy(rng.uniform(0.5, 3, (B, 1)).astype("float32")).cuda()
    res = {}
    try:
        for tc in (0, 1):
            lib.matcha_set_enc_tc(tc)
            model = build_model(ds, seed=1)
            model._engine().seed_base = 23
            if train:
                model.train()
                pred, rl = model(x, return_recon=True)
                (torch.nn.functional.binary_cross_entropy_with_logits(pred, y, weight=w) + 0.7 * rl.sum()).backward()
                grads = {k: p.grad.detach().cpu().numpy().copy() for k, p in model.named_parameters() if p.grad is not None}
            else:
                model.eval()
                with torch.no_grad():
                    pred, rl = model(x, return_recon=True)
                grads = {}
            with torch.no_grad():
                model.eval()
                emb = model.get_node_embeddings(torch.arange(0, 3 * N + 3).remainder(N + 1).view(-1, 1).cuda())[:, 0, :].cpu().numpy()
            res[tc] = (grads, pred.detach().cpu().numpy(), float(rl.sum()), emb)
    finally:
        lib.matcha_set_enc_tc(1)
    np.testing.assert_allclose(res[1][3], res[0][3], rtol=1e-4, atol=2e-6)          # embeddings.npy rows (eval, ids incl. 0)
    assert float(np.abs(res[1][3][0]).max()) == 0.0                                  # id 0 -> zero row
    np.testing.assert_allclose(res[1][1], res[0][1], rtol=1e-4, atol=5e-5)
    assert abs(res[1][2] - res[0][2]) <= 1e-4 * abs(res[0][2])
    assert res[0][0].keys() == res[1][0].keys()
    for k, g0 in res[0][0].items():
        g1 = res[1][0][k]
        scale = float(np.abs(g0).max())
        assert float(np.abs(g1 - g0).max()) <= 5e-4 * scale + 1e-7, (k, float(np.abs(g1 - g0).max()), scale)


def test_wide_chromosome_train_step_matches_oracle(monkeypatch):
    """Chromosomes of ~1000 bins (the cfg3 / cfg5 regime: several column groups in the encoder backward, a 16-chunk
    contraction in the forward, a ~1000-column reconstruction head): one training step with dropout ON against the
    oracle's fp64 autograd on the kernels' own dropout masks."""
    from matcha_b200.synthetic import build_model, make_dataset
    ds = make_dataset(WIDE, kmers_per_size=3000, seed=5)
    N = int(ds["chrom_range"][-1][1]) - 1
    monkeypatch.setattr(np.random, "choice", lambda a, size=None: np.asarray([0]))
    rng = np.random.default_rng(77)
    B, L = 320, 5
    xs = np.zeros((B, L), dtype=np.int64)
    for b in range(B):
        k = L if b % 2 == 0 else int(rng.integers(2, L + 1))
        xs[b, :k] = np.sort(rng.choice(np.arange(1, N + 1), size=k, replace=False))
    x = torch.from_numpy(xs).cuda()
    y = torch.from_numpy((rng.random((B, 1)) < 0.4).astype("float32")).cuda()
    w = torch.from_numpy(rng.uniform(0.5, 3.0, size=(B, 1)).astype("float32")).cuda()
    model = build_model(ds, seed=1)
    model.train()
    eng = model._engine()
    eng.ensure_bound()
    pred, recon = model(x, return_recon=True)
    (torch.nn.functional.binary_cross_entropy_with_logits(pred, y, weight=w) + 0.5 * recon.sum()).backward()
    om = oracle_from_model(model)
    out = O.loss_and_grads(om.to(torch.float64), x.cpu(), y.cpu().double(), w.cpu().double(), 1.0, 0.5, random_chrom=0,
                           train=True, seed=step_seed(eng))
    np.testing.assert_allclose(pred.detach().cpu().numpy(), out["logits"].numpy(), rtol=1e-4, atol=2e-4)
    assert abs(float(recon.sum()) - float(out["recon"])) <= 1e-4 * abs(float(out["recon"]))
    n_checked = 0
    for k, p in model.named_parameters():
        if p.grad is None:
            continue
        ref = out["grads"][k].numpy()
        scale = max(1e-12, float(np.abs(ref).max()))
        assert float(np.abs(p.grad.cpu().numpy() - ref).max()) <= 2e-3 * scale, (k, scale)
        n_checked += 1
    assert n_checked >= 30


def test_d128_pair_scorer_falls_back_to_generic_path(golden128):
    """embed_dim 128: all-pairs scoring goes through Classifier.forward on device-generated width-2 tuples."""
    from matcha_b200.scorer import PairScorer, pair_count, pair_index_to_ij
    model = model_from_golden(golden128, d=128)
    model.eval()
    lo, hi = (int(v) for v in golden128["chrom_range"][1])
    sc = PairScorer(model)
    total = pair_count(lo, hi, 1)
    got = sc.score_range(lo, hi, min_dis=1, p_begin=5, p_end=total - 3).cpu().numpy()
    i, j = pair_index_to_ij(np.arange(5, total - 3), lo, hi, 1)
    with torch.no_grad():
        want = model(torch.from_numpy(np.stack([i, j], 1)).cuda()).view(-1).cpu().numpy()
    np.testing.assert_array_equal(got, want)


# ------------------------------------------------------------------------------------------
# BASELINE.json's FULL sizes, through size-independent properties (the oracle does not finish these in seconds)
# ------------------------------------------------------------------------------------------
def test_cfg4_full_size_all_pairs_properties():
    """configs[3]: chr1 at 10 kb, 24,897 bins, n(n+1)/2 = 3.1e8 pairs (denoise_contact.py:67-88 with min_distance 0).
    (i) the 8 contiguous rank shards of `score_chromosome` concatenate bit-exactly to the single-GPU pass (no-communication
    sharding); (ii) 200k sampled pairs equal the fp64 closed form of SURVEY 8a evaluated on the host from the
    same tables; (iii) pair order is generate_pair_wise's: the first n outputs are row i = lo; (iv) sigmoid(logit) = prob."""
    L = _lib()
    lib = L.load()
    n, lo = 24897, 1
    g = torch.Generator(device="cuda").manual_seed(4)
    D = torch.randn(lo + n, 64, device="cuda", generator=g) * 0.5
    S = torch.randn(lo + n, 64, device="cuda", generator=g) * 0.5
    w = (torch.rand(64, device="cuda", generator=g) - 0.3) * 0.1
    b = torch.full((1,), -0.2, device="cuda")
    total = int(lib.matcha_pair_count(lo, lo + n, 0))
    assert total == n * (n + 1) // 2 == 309942753
    nbytes = int(lib.matcha_pair_tc_workspace_bytes(lo, lo + n))
    ws = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    L.check(lib.matcha_pair_tc_prepare(L.ptr(D), L.ptr(S), L.ptr(w), L.ptr(b), 64, lo, lo + n, L.ptr(ws), nbytes, L.stream_ptr()), "prep")
    full = torch.empty(total, dtype=torch.float32, device="cuda")
    L.check(lib.matcha_pair_tc_score_range(L.ptr(ws), lo, lo + n, 0, 0, total, 0, L.ptr(full), L.stream_ptr()), "full")
    world = 8
    for rank in range(world):
        pb, pe = total * rank // world, total * (rank + 1) // world
        part = torch.full((pe - pb,), float("nan"), device="cuda")
        L.check(lib.matcha_pair_tc_score_range(L.ptr(ws), lo, lo + n, 0, pb, pe, 0, L.ptr(part), L.stream_ptr()), "shard")
        assert torch.equal(part, full[pb:pe]), rank
        del part
    assert bool(torch.isfinite(full).all())
    # sampled pairs vs the closed form in fp64: logit(i, j) = 1/2 [sum_c w_c (D_jc - S_ic)^2 + sum_c w_c (D_ic - S_jc)^2] + b
    from matcha_b200.scorer import pair_index_to_ij
    rng = np.random.default_rng(0)
    p = np.unique(np.concatenate([rng.integers(0, total, 200_000), [0, n - 1, n, total - 1]]))
    i, j = pair_index_to_ij(p, lo, lo + n, 0)
    assert i[0] == lo and j[0] == lo and i[-1] == lo + n - 1 and j[-1] == lo + n - 1
    assert (i[p < n] == lo).all()                                     # row i = lo comes first, j ascending
    Dh, Sh, wh = D.double().cpu().numpy(), S.double().cpu().numpy(), w.double().cpu().numpy()
    want = 0.5 * (((Dh[j] - Sh[i]) ** 2) @ wh + ((Dh[i] - Sh[j]) ** 2) @ wh) + float(b)
    got = full[torch.from_numpy(p).cuda()].cpu().numpy()
    np.testing.assert_allclose(got, want, rtol=1e-4, atol=1e-4)
    prob = torch.empty(1 << 20, dtype=torch.float32, device="cuda")
    L.check(lib.matcha_pair_tc_score_range(L.ptr(ws), lo, lo + n, 0, 12345, 12345 + (1 << 20), 1, L.ptr(prob), L.stream_ptr()), "sig")
    np.testing.assert_allclose(prob.cpu().numpy(), torch.sigmoid(full[12345:12345 + (1 << 20)]).cpu().numpy(), rtol=2e-6, atol=1e-7)


def test_cfg3_full_size_model_properties(monkeypatch):
    """configs[2] bins: whole genome at 100 kb, 30,344 bins, chromosomes of up to 2,491 bins (39 encoder weight chunks, 7
    column groups in the encoder backward, 20 column blocks in the reconstruction head).  Properties that do not need the
    oracle: the tensor-core path agrees with the independent fp32 SIMT path on one training step (logits, recon loss, every
    gradient) and on eval logits; hyperedges are order-invariant; the k = 2 closed form equals the generic forward on pairs
    across the genome; id 0 embeds to zero."""
    from matcha_b200.scorer import PairScorer
    from matcha_b200.synthetic import build_model, make_dataset
    lib = _lib().load()
    ds = make_dataset("cfg3", kmers_per_size=20000, seed=11)
    N = ds["N"]
    assert N == 30344 and max(ds["nums"]) == 2491
    monkeypatch.setattr(np.random, "choice", lambda a, size=None: np.asarray([0]))     # recon on chr1 (2,491 columns)
    rng = np.random.default_rng(5)
    B, L = 600, 5
    pos = ds["positives"]
    xs = pos[rng.choice(len(pos), B, replace=False)].copy()
    x = torch.from_numpy(xs).cuda()
    y = torch.from_numpy((rng.random((B, 1)) < 0.4).astype("float32")).cuda()
    w = torch.from_numpy(rng.uniform(0.5, 3.0, size=(B, 1)).astype("float32")).cuda()
    model = build_model(ds, seed=1)
    res = {}
    try:
        for impl in (0, 1):
            lib.matcha_set_gemm_impl(impl)
            model.zero_grad(set_to_none=True)
            model.train()
            model._engine().seed_base, model._engine().tape_id = 23, 0
            pred, rl = model(x, return_recon=True)
            (torch.nn.functional.binary_cross_entropy_with_logits(pred, y, weight=w) + 0.5 * rl.sum()).backward()
            grads = {k: p.grad.detach().cpu().numpy().copy() for k, p in model.named_parameters() if p.grad is not None}
            model.eval()
            with torch.no_grad():
                ev = model(x).cpu().numpy()
            res[impl] = (pred.detach().cpu().numpy(), float(rl.detach().sum()), grads, ev)
    finally:
        lib.matcha_set_gemm_impl(1)
    np.testing.assert_allclose(res[1][0], res[0][0], rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(res[1][3], res[0][3], rtol=1e-4, atol=1e-4)
    assert abs(res[1][1] - res[0][1]) <= 1e-4 * abs(res[0][1])
    assert res[0][2].keys() == res[1][2].keys() and len(res[1][2]) >= 30
    for k, g0 in res[0][2].items():
        scale = float(np.abs(g0).max())
        assert float(np.abs(res[1][2][k] - g0).max()) <= 1e-3 * scale + 1e-7, (k, scale)
    # order invariance within a hyperedge (full-width rows: no padding moves)
    model.eval()
    full_rows = xs[(xs != 0).all(1)][:200]
    with torch.no_grad():
        a = model(torch.from_numpy(full_rows).cuda()).cpu().numpy()
        bperm = model(torch.from_numpy(np.ascontiguousarray(full_rows[:, ::-1])).cuda()).cpu().numpy()
    np.testing.assert_allclose(bperm, a, rtol=1e-4, atol=1e-4)
    # k = 2 closed form vs the generic forward, pairs inside chr1 (ids 1..2491) and inside the last chromosome
    sc = PairScorer(model)
    for (lo, hi) in (tuple(int(v) for v in ds["chrom_range"][0]), tuple(int(v) for v in ds["chrom_range"][-1])):
        total = int(lib.matcha_pair_count(lo, hi, 0))
        pb = total // 2
        got = sc.score_range(lo, hi, 0, pb, pb + 4096).cpu().numpy()
        from matcha_b200.scorer import pair_index_to_ij
        i, j = pair_index_to_ij(np.arange(pb, pb + 4096), lo, hi, 0)
        with torch.no_grad():
            want = model(torch.from_numpy(np.stack([i, j], 1)).cuda()).view(-1).cpu().numpy()
        np.testing.assert_allclose(got, want, rtol=1e-4, atol=1e-4)
    with torch.no_grad():
        emb = model.get_node_embeddings(torch.tensor([[0], [1], [N]]).cuda())[:, 0, :].cpu().numpy()
    assert float(np.abs(emb[0]).max()) == 0.0 and float(np.abs(emb[1]).max()) > 0 and float(np.abs(emb[2]).max()) > 0


def test_trainer_prefetch_and_host_loop_match_plain_steps(golden):
    """The next-batch prefetch of Trainer.step (negatives sampled on a side stream under the previous step) and the
    host-fed loop (pinned positives, prefetched H2D, asynchronous loss read-back) train exactly like plain step() calls:
    same sampler streams, same dropout seeds -> same per-step losses and the same weights after 6 steps."""
    from matcha_b200.sampler import KmerHashSet, NegativeSampler
    from matcha_b200.trainer import Trainer
    rng = np.random.default_rng(33)
    cr = golden["chrom_range"]
    kmers = _toy_kmers(rng, cr, 20000)
    hs = KmerHashSet(len(kmers), width=5).insert(kmers)
    P, steps = 512, 6
    pos_host = torch.from_numpy(kmers[rng.choice(len(kmers), P * steps, replace=False)]).pin_memory()
    w_host = torch.from_numpy(rng.uniform(0.5, 3.0, P * steps).astype(np.float32)).pin_memory()
    pos_dev, w_dev = pos_host.cuda(), w_host.cuda()
    results = {}
    for mode in ("plain", "prefetch", "host", "miss"):
        model = model_from_golden(golden)
        tr = Trainer(model, NegativeSampler(hs, cr, neg_num=3, seed=4), alpha=1.0, beta=0.3, seed=9)
        losses = []
        if mode == "host":
            out, h2d, d2h = tr.run_host_batches(pos_host, w_host, P, steps)
            assert h2d == P * 5 * 8 + P * 4 and d2h == 12
            losses = [out[i].numpy().copy() for i in range(steps)]
        else:
            for i in range(steps):
                cur = (pos_dev[i * P:(i + 1) * P], w_dev[i * P:(i + 1) * P])
                nxt = (pos_dev[(i + 1) * P:(i + 2) * P], w_dev[(i + 1) * P:(i + 2) * P]) if i + 1 < steps else (None, None)
                if mode == "plain":
                    tr.step(*cur)
                elif mode == "prefetch":
                    tr.step(*cur, *nxt)
                else:     # announce a batch, then pass a different tensor object holding the same rows: falls back, still correct
                    tr.step(cur[0].clone(), cur[1].clone(), *nxt)
                losses.append(tr.loss_out.cpu().numpy().copy())
        torch.cuda.synchronize()
        results[mode] = (np.stack(losses), tr.e.flat.detach().cpu().numpy().copy())
    for mode in ("prefetch", "host"):
        np.testing.assert_allclose(results[mode][0], results["plain"][0], rtol=2e-5, atol=1e-6)
        np.testing.assert_allclose(results[mode][1], results["plain"][1], rtol=1e-4, atol=2e-6)
    # a missed prefetch consumes extra sampler steps (different negatives) but must stay finite and trained
    assert np.isfinite(results["miss"][0]).all() and np.isfinite(results["miss"][1]).all()
