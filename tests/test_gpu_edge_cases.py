"""Edge cases of the hot path on the GPU (empty and ragged inputs, extreme widths, bad ids, a full positive set, empty
pair ranges) -- each against the oracle where a value is defined, or against the reference's error behaviour."""
import numpy as np
import pytest
import torch

from helpers import model_from_golden, oracle_from_model
from oracle import hypersagnn_oracle as O
from oracle import sampler_oracle as SO

pytestmark = pytest.mark.gpu


@pytest.fixture()
def model(golden):
    return model_from_golden(golden)


def test_empty_batch(model):
    model.eval()
    with torch.no_grad():
        out = model(torch.zeros(0, 5, dtype=torch.long).cuda())
    assert tuple(out.shape) == (0, 1)
    model.train()
    pred, recon = model(torch.zeros(0, 4, dtype=torch.long).cuda(), return_recon=True)
    assert tuple(pred.shape) == (0, 1) and recon.numel() == 1 and float(recon.sum()) == 0.0
    emb = model.get_node_embeddings(torch.zeros(0, 1, dtype=torch.long).cuda())
    assert emb.shape[0] == 0


@pytest.mark.parametrize("B", [7, 420])          # below / above the 1024-token switch to the tensor-core kernels
def test_all_pad_and_single_node_rows_match_oracle(golden, model, B):
    """Rows that are entirely padding (masked mean over zero tokens: 0 / 1e-15 = 0, Modules.py:309-311), rows with one or
    two real nodes, and full rows, in one batch: pads are live keys, so every row still runs the whole block."""
    N = int(golden["chrom_range"][-1][1]) - 1
    rng = np.random.default_rng(B)
    x = np.zeros((B, 5), dtype=np.int64)
    for b in range(B):
        k = b % 6                                 # 0 (all pad), 1, ..., 5 real nodes
        if k:
            x[b, :k] = np.sort(rng.choice(np.arange(1, N + 1), size=k, replace=False))
    model.eval()
    with torch.no_grad():
        got = model(torch.from_numpy(x).cuda()).cpu().numpy()
    want, _ = O.forward(oracle_from_model(model), torch.from_numpy(x))
    np.testing.assert_allclose(got, want.numpy(), rtol=1e-4, atol=5e-5)
    assert (got[0::6] == 0).all()                 # all-pad rows score exactly 0


@pytest.mark.parametrize("L,B", [(2, 60), (2, 700), (6, 40), (6, 300)])
def test_extreme_widths_match_oracle(golden, model, L, B):
    """Narrowest (pairs) and widest supported padded width, on both kernel families."""
    N = int(golden["chrom_range"][-1][1]) - 1
    rng = np.random.default_rng(10 * L + B)
    x = np.zeros((B, L), dtype=np.int64)
    for b in range(B):
        k = int(rng.integers(2, L + 1))
        x[b, :k] = np.sort(rng.choice(np.arange(1, N + 1), size=k, replace=False))
    model.eval()
    with torch.no_grad():
        got = model(torch.from_numpy(x).cuda()).cpu().numpy()
    want, _ = O.forward(oracle_from_model(model), torch.from_numpy(x))
    np.testing.assert_allclose(got, want.numpy(), rtol=1e-4, atol=5e-5)


def test_bad_inputs_raise(golden, model):
    from matcha_b200 import MatchaError
    N = int(golden["chrom_range"][-1][1]) - 1
    model.eval()
    with torch.no_grad():
        with pytest.raises(MatchaError):          # nn.Embedding would raise IndexError in the reference
            model(torch.tensor([[1, N + 1, 0]]).cuda())
        with pytest.raises(MatchaError):
            model(torch.tensor([[-1, 2, 3]]).cuda())
        with pytest.raises(MatchaError):          # padded widths outside 2..6 are rejected, not silently mis-scored
            model(torch.ones(3, 7, dtype=torch.long).cuda())
        with pytest.raises(MatchaError):
            model(torch.ones(3, dtype=torch.long).cuda())
        ok = model(torch.tensor([[1, N, 0]]).cuda())          # the largest id is valid
    assert bool(torch.isfinite(ok).all())


def test_hashset_duplicates_capacity_and_empty(golden):
    from matcha_b200 import MatchaError
    from matcha_b200.sampler import KmerHashSet
    rows = np.asarray([[1, 2, 0, 0, 0], [1, 2, 3, 0, 0], [4, 9, 11, 12, 13]], dtype=np.int64)
    hs = KmerHashSet(16, width=5)
    assert not hs.contains(rows).any()                        # empty table
    hs.insert(rows).insert(rows)                              # duplicates are idempotent for membership
    assert hs.contains(rows).all() and not hs.overflowed()
    probe = np.asarray([[1, 2, 4, 0, 0], [2, 1, 0, 0, 0], [1, 0, 0, 0, 0]], dtype=np.int64)      # near misses, prefix
    assert not hs.contains(probe).any()
    assert hs.contains(np.asarray([[1, 2]], dtype=np.int64)).all()                                # narrower rows are zero padded
    with pytest.raises(MatchaError):
        hs.insert(np.arange(1, 1 + 5 * 4096, dtype=np.int64).reshape(-1, 5))                      # over capacity
    with pytest.raises(MatchaError):
        hs.contains(np.ones((1, 6), dtype=np.int64))                                              # wider than the table


def test_sampler_full_positive_set_is_flagged_bit_exact():
    """Every same-chromosome corruption is a positive: the rounds run out, the row comes back as the positive itself with
    valid = 0 (the reference would loop forever, main.py:399-428) -- identical to the oracle, rounds included."""
    from matcha_b200.sampler import KmerHashSet, NegativeSampler
    cr = np.asarray([[1, 4], [4, 9]])
    full = np.asarray([[1, 2, 0], [1, 3, 0], [2, 3, 0]], dtype=np.int64)
    extra = np.asarray([[4, 5, 6], [5, 7, 8]], dtype=np.int64)
    kmers = np.concatenate([full, extra])
    hs = KmerHashSet(64, width=3).insert(kmers)
    pos = np.asarray([[1, 2, 0], [4, 5, 6], [2, 3, 0]], dtype=np.int64)
    smp = NegativeSampler(hs, cr, min_dis=0, neg_num=2, seed=5, max_rounds=8)
    rounds = torch.zeros(6, dtype=torch.int32, device="cuda")
    neg, valid = smp.sample(torch.from_numpy(pos).cuda(), rounds=rounds, step=3)
    want_neg, want_valid, want_rounds = SO.sample_negatives(pos, SO.build_set(kmers), cr, 2, 0, seed=5, step=3, max_rounds=8)
    assert (neg.cpu().numpy() == want_neg).all() and (valid.cpu().numpy() == want_valid).all()
    assert (rounds.cpu().numpy() == want_rounds).all()
    v = valid.cpu().numpy()
    assert (v[[0, 1, 4, 5]] == 0).all() and (neg.cpu().numpy()[0] == pos[0]).all()      # chromosome 1 is saturated
    assert v[2:4].all()                                                                # chromosome 2 still has free triplets


def test_pair_range_edges(golden, model):
    from matcha_b200 import MatchaError
    from matcha_b200.scorer import PairScorer, pair_count
    model.eval()
    lo, hi = (int(v) for v in golden["chrom_range"][0])
    sc = PairScorer(model)
    assert pair_count(lo, hi, hi - lo) == 0                    # min_distance beyond the chromosome: no pairs
    assert sc.score_range(lo, hi, min_dis=hi - lo).numel() == 0
    total = pair_count(lo, hi, 0)
    assert sc.score_range(lo, hi, 0, 17, 17).numel() == 0      # empty shard
    last = sc.score_range(lo, hi, 0, total - 1, total).cpu().numpy()
    with torch.no_grad():
        want = model(torch.tensor([[hi - 1, hi - 1]]).cuda()).view(-1).cpu().numpy()      # the last pair is (n, n): j = i is included
    np.testing.assert_allclose(last, want, rtol=1e-4, atol=5e-5)
    with pytest.raises(MatchaError):
        sc.score_range(lo, hi, 0, 0, total + 1)


def test_score_tuples_ragged_batches(golden, model):
    """predict_multiway.py:74-87 on tuples of mixed size: each batch is padded to ITS longest tuple, and the score of a
    tuple depends on that width -- so the result must equal Classifier.forward on exactly that padding."""
    from matcha_b200.scorer import score_tuples
    N = int(golden["chrom_range"][-1][1]) - 1
    rng = np.random.default_rng(8)
    samples = []
    for i in range(25):
        k = int(rng.integers(2, 4)) if i < 10 else int(rng.integers(2, 6))      # first batch: width 3, later ones: width 5
        samples.append(sorted(rng.choice(np.arange(1, N + 1), size=k, replace=False).tolist()))
    outs = score_tuples(model, samples, batch_size=10)
    assert [j for j, _ in outs] == [0, 1, 2]
    model.eval()
    for j, o in outs:
        chunk = samples[j * 10:(j + 1) * 10]
        L = max(len(s) for s in chunk)
        x = np.zeros((len(chunk), L), dtype=np.int64)
        for i, s in enumerate(chunk):
            x[i, :len(s)] = s
        want, _ = O.forward(oracle_from_model(model), torch.from_numpy(x))
        np.testing.assert_allclose(o.cpu().numpy(), want.numpy(), rtol=1e-4, atol=5e-5)


# ------------------------------------------------------------------------------------------
# k-mer enumeration + counting (generate_kmers.py) -- bit-exact vs the reference's own output and vs the oracle
# ------------------------------------------------------------------------------------------
def test_kmer_counting_matches_reference_output_bit_exact():
    import os
    from conftest import GOLDEN
    from matcha_b200.kmers import count_kmers
    g = np.load(os.path.join(GOLDEN, "kmer_small.npz"))
    for ci, (k, min_dis, max_size, min_freq) in enumerate(g["cases"]):
        rows, freq = count_kmers(g["members"], g["offsets"], int(k), int(min_dis), int(max_size), int(min_freq))
        want_rows, want_freq = g[f"rows/{ci}"], g[f"freq/{ci}"]
        assert rows.shape == want_rows.shape, (ci, rows.shape, want_rows.shape)
        assert (rows == want_rows).all() and (freq == want_freq).all(), ci


def test_kmer_counting_matches_oracle_and_edges():
    from matcha_b200 import MatchaError
    from matcha_b200.kmers import clusters_to_csr, count_kmers
    from oracle import kmer_oracle as KO
    rng = np.random.default_rng(12)
    clusters = []
    for _ in range(3000):
        size = int(min(40, 1 + rng.geometric(0.2)))                      # sizes 2..40: some above max_cluster_size, some of one node
        a = int(rng.integers(1, 5000))
        clusters.append(sorted(set(int(np.clip(a + rng.integers(-15, 16), 1, 60000)) for _ in range(size))))
    members, offsets = clusters_to_csr(clusters)
    for (k, min_dis, max_size, min_freq) in [(2, 0, 25, 2), (3, 1, 25, 2), (5, 0, 14, 1), (6, 0, 10, 1)]:
        rows, freq = count_kmers(members, offsets, k, min_dis, max_size, min_freq)
        want_rows, want_freq = KO.count_kmers(clusters, k, min_dis, max_size, min_freq)
        assert rows.shape == want_rows.shape and (rows == want_rows).all() and (freq == want_freq).all(), (k, min_dis)
    # nothing eligible / nothing frequent enough -> empty results of the right shape
    r, f = count_kmers(*clusters_to_csr([[1, 2], [3, 4]]), 3)
    assert r.shape == (0, 3) and f.shape == (0,)
    r, f = count_kmers(*clusters_to_csr([[1, 2, 3], [4, 5, 6]]), 3, min_freq_cutoff=2)
    assert r.shape == (0, 3)
    r, f = count_kmers(*clusters_to_csr([[1, 2, 3], [1, 2, 3], [1, 2, 9]]), 2, min_freq_cutoff=2)
    assert r.tolist() == [[1, 2], [1, 3], [2, 3]] and f.tolist() == [3, 2, 2]
    with pytest.raises(MatchaError):                                     # a table too small for the distinct k-mers reports it
        count_kmers(members, offsets, 3, capacity=64)
    with pytest.raises(MatchaError):
        count_kmers(*clusters_to_csr([[1, 2, 1 << 21]]), 2, min_freq_cutoff=1)
