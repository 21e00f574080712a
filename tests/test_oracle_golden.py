"""The oracle restatement vs golden vectors produced by the UNMODIFIED reference
(oracle/make_golden.py).  CPU only."""
import json

import numpy as np
import pytest
import torch

from oracle import hypersagnn_oracle as O


@pytest.fixture(scope="module")
def model(golden):
    return O.model_from_npz(golden)


@pytest.mark.parametrize("L", [2, 3, 4, 5])
def test_eval_logits_match_reference(golden, model, L):
    x = torch.from_numpy(golden[f"x/L{L}"])
    got, _ = O.forward(model, x)
    np.testing.assert_allclose(got.numpy(), golden[f"logits_eval/L{L}"], rtol=1e-4, atol=2e-5)


@pytest.mark.parametrize("L", [2, 5])
def test_recon_loss_matches_reference(golden, model, L):
    x = torch.from_numpy(golden[f"x/L{L}"])
    for r in range(len(golden["nums"])):
        _, rl = O.forward(model, x, random_chrom=r)
        np.testing.assert_allclose(rl.numpy(), golden[f"recon_eval/L{L}/r{r}"], rtol=1e-5)


def test_embeddings_match_reference(golden, model):
    N = golden["embeddings"].shape[0]
    got = O.node_embeddings(model, torch.arange(1, N + 1))
    np.testing.assert_allclose(got.numpy(), golden["embeddings"], rtol=1e-4, atol=1e-6)


@pytest.mark.parametrize("L", [3, 5])
def test_gradients_match_reference(golden, L):
    m = O.model_from_npz(golden, dtype=torch.float64)
    x = torch.from_numpy(golden[f"x/L{L}"])
    y = torch.from_numpy(golden[f"y/L{L}"]).double()
    w = torch.from_numpy(golden[f"w/L{L}"]).double()
    r = int(golden[f"train_rchrom/L{L}"][0])
    m.p_feature = m.p_attn = m.p_pff = 0.0
    out = O.loss_and_grads(m, x, y, w, alpha=1.0, beta=0.5, random_chrom=r, train=True)
    np.testing.assert_allclose(out["logits"].numpy(), golden[f"train_logits/L{L}"], rtol=1e-4, atol=2e-5)
    np.testing.assert_allclose(out["bce"].numpy(), golden[f"train_bce/L{L}"][0], rtol=1e-5)
    np.testing.assert_allclose(out["recon"].numpy(), golden[f"train_recon/L{L}"], rtol=1e-5)
    meta = json.loads(str(golden["meta"]))
    live_ref = set(meta["live_keys_by_L"][str(L)])
    for k, g in out["grads"].items():
        if k in live_ref:
            ref = golden[f"grad/L{L}/{k}"]
            scale = float(np.abs(ref).max())
            # atol 1e-7: the key-side LayerNorm bias has an exactly-zero true gradient (softmax is
            # invariant to a per-query constant); the fp32 reference holds ~1e-8 rounding noise there
            assert np.abs(g.numpy() - ref).max() <= 2e-4 * scale + 1e-7, k
        else:   # recon heads of chromosomes that were not drawn: no gradient in the reference
            assert float(g.abs().max()) == 0.0, k
    # every key the reference gives a gradient to is one the oracle calls live
    assert live_ref <= set(out["grads"].keys())


def test_known_answer_facts(golden, model):
    """pad-width dependence, permutation invariance, k=2 closed form."""
    pw = golden["kat/pad_width"]
    for i, x in enumerate([[3, 47, 90], [3, 47, 90, 0], [3, 47, 90, 0, 0]]):
        got, _ = O.forward(model, torch.tensor([x]))
        assert abs(got.item() - pw[i]) < 5e-5
    assert abs(pw[0] - pw[1]) > 1e-3 and abs(pw[1] - pw[2]) > 1e-4       # padding changes the score
    got, _ = O.forward(model, torch.tensor([[90, 3, 47]]))
    assert abs(got.item() - golden["kat/permuted"][0]) < 5e-5
    assert abs(golden["kat/permuted"][0] - pw[0]) < 5e-6
    pairs = torch.from_numpy(golden["kat/pairs"])
    cf = O.pair_logits_closed_form(model, pairs)
    np.testing.assert_allclose(cf.numpy(), golden["kat/pair_logits"], rtol=1e-4, atol=2e-5)


def test_state_dict_inventory(golden):
    meta = json.loads(str(golden["meta"]))
    keys = meta["state_dict_keys"]
    assert "node_embedding.Embedding_Linear0.tied weight_0" in keys       # names contain a space
    assert keys["encode1.mul_head_attn.w_qs.weight"] == [512, 64]
    assert "encode2.pff_n2.PWF_Conv1.weight" in meta["no_grad_keys"]
    assert "encode1.mul_head_attn.fc2.weight" in meta["no_grad_keys"]


def test_counter_rng_vector_matches_scalar():
    seed, site = 1234, O.SITE_ATTN
    keep = O.dropout_keep_mask(seed, site, np.asarray([0, 7, 123456]), 10, 0.3)
    key = O.site_key(seed, site)
    for ri, row in enumerate([0, 7, 123456]):
        for j in range(10):
            r = O.splitmix64((key + (row << 24) + (j >> 2)) & ((1 << 64) - 1))
            v = (r >> (16 * (j & 3))) & 0xFFFF
            assert bool(keep[ri, j]) == (v >= O.dropout_threshold16(0.3))
    big = O.dropout_keep_mask(7, O.SITE_PFF, np.arange(4000), 64, 0.4)
    assert abs(big.mean() - 0.6) < 0.01


# ---- embed_dim 128 (BASELINE.json configs[4]) --------------------------------------------------------------
@pytest.mark.parametrize("L", [3, 5])
def test_d128_eval_logits_and_recon_match_reference(golden128, L):
    m = O.model_from_npz(golden128)
    x = torch.from_numpy(golden128[f"x/L{L}"])
    got, _ = O.forward(m, x)
    np.testing.assert_allclose(got.numpy(), golden128[f"logits_eval/L{L}"], rtol=1e-4, atol=2e-5)
    for r in range(len(golden128["nums"])):
        _, rl = O.forward(m, x, random_chrom=r)
        np.testing.assert_allclose(rl.numpy(), golden128[f"recon_eval/L{L}/r{r}"], rtol=1e-5)
    N = golden128["embeddings"].shape[0]
    emb = O.node_embeddings(m, torch.arange(1, N + 1))
    assert emb.shape[1] == 128
    np.testing.assert_allclose(emb.numpy(), golden128["embeddings"], rtol=1e-4, atol=1e-6)


def test_d128_gradients_match_reference(golden128):
    L = 5
    m = O.model_from_npz(golden128, dtype=torch.float64)
    x = torch.from_numpy(golden128[f"x/L{L}"])
    y = torch.from_numpy(golden128[f"y/L{L}"]).double()
    w = torch.from_numpy(golden128[f"w/L{L}"]).double()
    r = int(golden128[f"train_rchrom/L{L}"][0])
    m.p_feature = m.p_attn = m.p_pff = 0.0
    out = O.loss_and_grads(m, x, y, w, alpha=1.0, beta=0.5, random_chrom=r, train=True)
    np.testing.assert_allclose(out["logits"].numpy(), golden128[f"train_logits/L{L}"], rtol=1e-4, atol=2e-5)
    meta = json.loads(str(golden128["meta"]))
    live_ref = set(meta["live_keys_by_L"][str(L)])
    checked = 0
    for k, g in out["grads"].items():
        if k not in live_ref:
            assert float(g.abs().max()) == 0.0, k
        elif f"grad/L{L}/{k}" in golden128.files:        # full gradient kept for the tensors up to 20 000 elements
            ref = golden128[f"grad/L{L}/{k}"]
            assert np.abs(g.numpy() - ref).max() <= 2e-4 * float(np.abs(ref).max()) + 1e-7, k
            checked += 1
        else:                                            # Frobenius norm for the large ones
            ref = float(golden128[f"gradnorm/L{L}/{k}"][0])
            assert abs(float(g.double().norm()) - ref) <= 2e-4 * ref + 1e-9, k
            checked += 1
    assert checked == len(live_ref)


# ------------------------------------------------------------------------------------------
# k-mer enumeration / counting oracle vs the unmodified reference's build_dict
# ------------------------------------------------------------------------------------------
def test_kmer_oracle_matches_reference_build_dict():
    import os

    import numpy as np
    from conftest import GOLDEN
    from oracle import kmer_oracle as KO
    g = np.load(os.path.join(GOLDEN, "kmer_small.npz"))
    members, offsets = g["members"], g["offsets"]
    clusters = [members[offsets[i]:offsets[i + 1]] for i in range(len(offsets) - 1)]
    assert len(g["cases"]) == 7
    for ci, (k, min_dis, max_size, min_freq) in enumerate(g["cases"]):
        rows, freq = KO.count_kmers(clusters, int(k), int(min_dis), int(max_size), int(min_freq))
        assert rows.shape == g[f"rows/{ci}"].shape and (rows == g[f"rows/{ci}"]).all(), (k, min_dis)
        assert (freq == g[f"freq/{ci}"]).all()
        # domain facts: rows ascending with every gap > min_distance; frequencies respect the cutoff
        assert (np.diff(rows, axis=1) > min_dis).all() and (freq >= min_freq).all()
    # enumeration work the kernel has to cover
    sizes = np.diff(offsets)
    assert KO.n_subsets(sizes, 3, 25).sum() == sum(len(list(__import__("itertools").combinations(range(int(n)), 3)))
                                                   for n in sizes if 3 <= n <= 25)
