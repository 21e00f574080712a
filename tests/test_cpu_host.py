"""CPU-only tests: the C ABI loads and exports what the header declares, host-side logic, the sampler
oracle's distribution against the reference's own generate_negative, and the 2-rank (gloo) data-parallel
plumbing.  No kernel is launched here."""
import ctypes
import json
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ------------------------------------------------------------------------------------------
# C ABI
# ------------------------------------------------------------------------------------------
def test_library_builds_and_exports_every_declared_symbol():
    from matcha_b200 import _lib
    lib = _lib.load()
    header = open(os.path.join(ROOT, "include", "matcha_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(matcha_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.matcha_version() >= 100


def test_model_desc_layout_matches_the_c_struct(tmp_path):
    from matcha_b200 import _lib
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "matcha_b200.h"\nint main(){printf("%zu %zu %zu %zu",'
                   'sizeof(matcha_model_desc), offsetof(matcha_model_desc, chrom_start), offsetof(matcha_model_desc, feat),'
                   'offsetof(matcha_model_desc, p_feature));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    size, o_cs, o_feat, o_p = (int(v) for v in subprocess.check_output([str(exe)]).split())
    D = _lib.ModelDesc
    assert ctypes.sizeof(D) == size
    assert D.chrom_start.offset == o_cs and D.feat.offset == o_feat and D.p_feature.offset == o_p


def test_pair_count_and_index_inverse():
    from matcha_b200.scorer import pair_count, pair_index_to_ij
    for lo, hi, md in [(1, 11, 0), (5, 40, 0), (5, 40, 3), (1, 4, 5), (7, 8, 0)]:
        ref = [(i, j) for i in range(lo, hi) for j in range(i + md, hi)]      # denoise_contact.py:67-74
        assert pair_count(lo, hi, md) == len(ref)
        if ref:
            i, j = pair_index_to_ij(np.arange(len(ref)), lo, hi, md)
            assert list(zip(i.tolist(), j.tolist())) == ref
    n = 24897                                                                  # chr1 @ 10 kb (cfg4)
    assert pair_count(1, n + 1, 0) == n * (n + 1) // 2 == 309942753


def test_product_path_fails_loudly_without_cuda(golden):
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from helpers import model_from_golden
    from matcha_b200 import MatchaError
    m = model_from_golden(golden)
    with pytest.raises(MatchaError):
        m(torch.tensor([[1, 2, 3]]))
    with pytest.raises(RuntimeError):
        m.encode1.mul_head_attn(None, None, None, None)      # sub-modules own parameters only


# ------------------------------------------------------------------------------------------
# host mirror of the reference operator surface
# ------------------------------------------------------------------------------------------
def test_state_dict_keys_and_shapes_equal_the_reference(golden):
    from helpers import model_from_golden
    m = model_from_golden(golden)
    meta = json.loads(str(golden["meta"]))
    ours = {k: list(v.shape) for k, v in m.state_dict().items()}
    assert ours == meta["state_dict_keys"]
    assert sum(1 for k in ours if " " in k) > 0             # 'tied weight_0' etc. keep their space


def test_zscore_restatement_matches_reference(golden):
    from matcha_b200.hyper_sagnn import zscore_positive_rows
    got = zscore_positive_rows(golden["inter_raw"].copy())
    np.testing.assert_allclose(got, golden["inter"], rtol=1e-6, atol=1e-7)


def test_data_generator_per_size_pools():
    from matcha_b200.hyper_sagnn import DataGenerator, pad_edges
    rng = np.random.default_rng(0)
    edges = [sorted(rng.choice(np.arange(1, 100), size=int(k), replace=False)) for k in rng.integers(2, 6, 500)]
    w = rng.random(500).astype(np.float32)
    np.random.seed(0)
    dg = DataGenerator(edges, w, batch_size=8, num_batch_per_iter=10, min_size=2, max_size=5)
    e, ww = dg.next_iter()
    assert e.shape == (4 * 80, 5) and ww.shape == (4 * 80,)
    sizes = (e != 0).sum(1)
    assert [(sizes == k).sum() for k in (2, 3, 4, 5)] == [80] * 4         # 80 per size per call (Modules.py:658-666)
    padded = pad_edges(edges, 5)
    lookup = {tuple(r): wi for r, wi in zip(padded, w)}
    for r, wi in zip(e, ww):
        assert lookup[tuple(r)] == wi                                       # weights stay attached to their edge
    for _ in range(30):                                                     # wrap-around reshuffles, never fails
        e, ww = dg.next_iter()
        assert e.shape == (320, 5)


def test_synthetic_dataset_shapes_follow_the_reference_binning():
    from matcha_b200.synthetic import chrom_bins, make_dataset, CONFIGS
    nums, cr = chrom_bins(*CONFIGS["cfg1"][:2])
    assert nums == [250, 244] and cr.tolist() == [[1, 251], [251, 495]]     # N = 494
    nums2, _ = chrom_bins(*CONFIGS["cfg2"][:2])
    assert sum(nums2) == 3067 and len(nums2) == 23
    assert sum(chrom_bins(*CONFIGS["cfg4"][:2])[0]) == 24897
    ds = make_dataset("cfg1", kmers_per_size=3000, seed=0)
    assert ds["N"] == 494 and ds["attr"].shape == (495, 3)
    for k, rows in ds["kmers"].items():
        assert rows.shape[1] == k and (np.diff(rows, axis=1) > 0).all() and rows.min() >= 1 and rows.max() <= 494
    assert ds["positives"].shape[1] == 5 and len(ds["positives"]) < len(ds["dict"])
    assert abs(ds["pos_weight"].mean() - 3.0) < 1e-4                        # weight /= mean; *= neg_num (main.py:594-595)
    assert [f.shape for f in ds["features"]] == [(250, 250), (244, 244)]


# ------------------------------------------------------------------------------------------
# sampler oracle: exact-set semantics and agreement in distribution with the reference's sampler
# ------------------------------------------------------------------------------------------
def _toy():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    rng = np.random.default_rng(0)
    nums = [120, 90, 150]
    starts = np.concatenate([[0], np.cumsum(nums)])
    cr = np.stack([starts[:-1] + 1, starts[1:] + 1], 1).astype(np.int64)
    rows = []
    while len(rows) < 6000:
        c = int(rng.integers(0, 3))
        a = int(rng.integers(cr[c, 0], cr[c, 1]))
        ids = {a}
        while len(ids) < 4:
            ids.add(int(np.clip(a + rng.integers(-12, 13), cr[c, 0], cr[c, 1] - 1)))
        rows.append(sorted(ids))
    return cr, np.unique(np.asarray(rows, dtype=np.int64), axis=0)


def test_sampler_oracle_properties_and_reference_distribution():
    from oracle import sampler_oracle as SO
    cr, kmers = _toy()
    s = SO.build_set(kmers)
    pos = np.concatenate([kmers[:2880], np.zeros((2880, 1), dtype=np.int64)], 1)      # width 5, k = 4
    neg, valid, rounds = SO.sample_negatives(pos, s, cr, neg_num=3, min_dis=0, seed=1, step=0)
    assert valid.all() and (rounds >= 1).all()
    chrom = lambda v: int(np.searchsorted(cr[:, 1], v, side="right"))
    hist = np.zeros(5)
    for g, row in enumerate(neg):
        live, p = row[:4], pos[g // 3][:4]
        assert row[4] == 0 and (np.diff(live) > 0).all() and tuple(live) not in s
        assert sorted(chrom(v) for v in live) == sorted(chrom(v) for v in p)            # same-chromosome replacement
        hist[4 - len(set(live.tolist()) & set(p.tolist()))] += 1
    hist /= hist.sum()
    ref = json.load(open(os.path.join(ROOT, "tests", "golden", "sampler_stats.json")))["sizes"]["4"]
    ref_hist = np.asarray(ref["changed_hist"]) / ref["n"]
    assert ref["in_positive_set"] == 0 and ref["same_chrom_multiset"] == ref["n"]
    # both follow Binomial(4, 1/2) | > 0  = [4, 6, 4, 1] / 15 up to collisions; 8640 draws -> sigma ~ 0.005
    assert np.abs(hist - ref_hist).max() < 0.025
    assert np.abs(hist[1:] - np.asarray([4, 6, 4, 1]) / 15).max() < 0.03
    # mean candidate rounds per negative in the range the survey measured for real data (~1.1 - 1.3)
    assert 1.0 <= rounds.mean() < 1.6


def test_sampler_oracle_is_deterministic_and_counter_based():
    from oracle import sampler_oracle as SO
    cr, kmers = _toy()
    s = SO.build_set(kmers)
    pos = np.concatenate([kmers[:20], np.zeros((20, 1), dtype=np.int64)], 1)
    a = SO.sample_negatives(pos, s, cr, seed=3, step=7)
    b = SO.sample_negatives(pos, s, cr, seed=3, step=7)
    c = SO.sample_negatives(pos, s, cr, seed=3, step=8)
    assert (a[0] == b[0]).all() and not (a[0] == c[0]).all()
    # exhausting the rounds falls back to the positive with valid = 0
    tiny_cr = np.asarray([[1, 4]])
    full = {(1, 2), (1, 3), (2, 3)}
    neg, valid, rounds = SO.sample_negatives(np.asarray([[1, 2]]), full, tiny_cr, neg_num=2, max_rounds=8)
    assert (valid == 0).all() and (neg == np.asarray([[1, 2]])).all() and (rounds == 8).all()


# ------------------------------------------------------------------------------------------
# data-parallel plumbing on 2 gloo ranks
# ------------------------------------------------------------------------------------------
def _dp_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    from matcha_b200.parallel import allreduce_grads_and_flags, init_from_env, range_shard, shard_rows
    r, w, _ = init_from_env("gloo")
    n_flat, n_flag = 1000, 6
    g = torch.Generator().manual_seed(100 + r)
    gflat = torch.zeros(n_flat + n_flag)
    gflat[:n_flat] = torch.randn(n_flat, generator=g)
    active = torch.tensor([1, 0, r, 0, 1 - r, 0], dtype=torch.int32)
    mine = gflat[:n_flat].clone()
    scale = allreduce_grads_and_flags(gflat, n_flat, active, w)
    rows = list(range(10))[shard_rows(10, r, w)]
    q.put((r, mine, gflat[:n_flat].clone(), active.clone(), scale, rows, range_shard(101, r, w)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_allreduce_of_gradients_and_flags():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29000 + os.getpid() % 2000
    procs = [ctx.Process(target=_dp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(2)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    total = res[0][1] + res[1][1]
    for r in range(2):
        assert torch.allclose(res[r][2], total)                       # both ranks hold the summed gradient
        assert res[r][3].tolist() == [1, 0, 1, 0, 1, 0]               # flags OR-reduced in the same collective
        assert res[r][4] == 0.5
    assert res[0][5] == [0, 2, 4, 6, 8] and res[1][5] == [1, 3, 5, 7, 9]
    assert res[0][6] == (0, 50) and res[1][6] == (50, 101)


def test_bench_steps_are_rank_symmetric():
    """Every rank must execute the same training steps: a step holds an NCCL all-reduce when world > 1, so a step that only
    rank 0 runs deadlocks the job (regression guard for bench.py's clock-sampling warm-up)."""
    import re
    src = open(os.path.join(ROOT, "bench.py")).read()
    body = src[src.index("def main():"):]
    lines = body.split("\n")
    for i, ln in enumerate(lines):
        if re.match(r"\s*if rank == 0:", ln):
            indent = len(ln) - len(ln.lstrip())
            j = i + 1
            while j < len(lines) and (not lines[j].strip() or len(lines[j]) - len(lines[j].lstrip()) > indent):
                assert not re.search(r"\b(run|timed|trainer\.step)\(", lines[j]), f"rank-0-only step at bench.py main() line {j}: {lines[j].strip()}"
                j += 1


def test_bench_reference_arm_prints_one_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs beside ours): exactly one stdout line, valid JSON, the
    contract's keys, our arm's metric / unit / config, no GPU needed."""
    import json
    import subprocess
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                          "--kmers-per-size", "20000"], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [ln for ln in res.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, res.stdout[:500]
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "hyperedges_per_s_trained" and d["unit"] == "hyperedges/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["steps"] == 1 and d["warmup"] == 1
    assert d["config"]["workload"].startswith("cfg2") and d["gpu_launches"] == 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_denoise_postprocessing_matches_the_reference_script():
    """oracle.denoise_oracle (the numpy restatement of denoise_contact.py:31-61,160-189) against matrices and pixel values produced
    by EXECUTING the reference's own loop (oracle/make_denoise_golden.py): same pair order (generate_pair_wise), same
    0-based global bin ids, same denoised matrix `my` and `balanced` values, min_distance 0 and 2, chromosomes with gaps."""
    from sklearn.preprocessing import QuantileTransformer
    from oracle.denoise_oracle import denoise_matrix
    from matcha_b200.scorer import pair_index_to_ij
    g = np.load(os.path.join(ROOT, "tests", "golden", "denoise_small.npz"))
    cr, origin = g["chrom_range"], g["origin"]
    for ci in range(2):
        min_dis = int(g[f"min_dis/{ci}"])
        for c, (lo, hi) in enumerate(cr):
            lo, hi = int(lo), int(hi)
            n = hi - lo
            logits = g[f"logits/{ci}/{c}"]
            total = len(logits)
            full = n - min_dis
            assert total == full * (full + 1) // 2                                    # == matcha_pair_count(lo, hi, min_dis)
            ii, jj = pair_index_to_ij(np.arange(total), lo, hi, min_dis)
            assert (ii - 1 == g[f"bin1/{ci}/{c}"]).all() and (jj - 1 == g[f"bin2/{ci}/{c}"]).all()
            proba = torch.sigmoid(torch.from_numpy(logits)).numpy()                   # denoise_contact.py:155
            weight = origin[ii - 1, jj - 1]                                           # :160
            tr = QuantileTransformer(n_quantiles=1000, output_distribution="uniform")
            my = denoise_matrix(n, ii - lo, jj - lo, proba, weight, tr)
            np.testing.assert_allclose(my, g[f"my/{ci}/{c}"], rtol=1e-6, atol=1e-7)
            np.testing.assert_allclose(my[ii - lo, jj - lo], g[f"balanced/{ci}/{c}"], rtol=1e-6, atol=1e-7)


def test_every_import_of_the_gpu_tests_resolves():
    """Round-1 regression: a GPU test imported a helper module that a cleanup had moved away, which only showed on the GPU
    box.  Walk every import statement of the GPU test files (module level AND inside test bodies) and resolve it here."""
    import ast
    import importlib
    here = os.path.dirname(os.path.abspath(__file__))
    extra = [here, os.path.join(ROOT, "scripts")]      # test_gpu_scripts.py inserts scripts/ itself (pickle module path)
    sys.path[:0] = extra
    try:
        for fn in sorted(os.listdir(here)):
            if not (fn.startswith("test_gpu") and fn.endswith(".py")):
                continue
            tree = ast.parse(open(os.path.join(here, fn)).read())
            for node in ast.walk(tree):
                if isinstance(node, ast.Import):
                    for a in node.names:
                        importlib.import_module(a.name)
                elif isinstance(node, ast.ImportFrom) and node.level == 0:
                    mod = importlib.import_module(node.module)
                    for a in node.names:
                        if not hasattr(mod, a.name):
                            importlib.import_module(node.module + "." + a.name)
    finally:
        for e in extra:
            sys.path.remove(e)


def test_gpu_suite_collects():
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests"), "--collect-only", "-q", "-m", "gpu"],
                       capture_output=True, text=True, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    n = [ln for ln in r.stdout.splitlines() if "::" in ln]
    assert len(n) >= 90, len(n)


def test_features_oracle_matches_the_reference_functions():
    """oracle/features_oracle.py (numpy restatement of process.py:90-105 and :148-170) against arrays written by the
    UNMODIFIED functions (oracle/make_features_golden.py -> tests/golden/features_small.npz): float64, bit for bit."""
    from oracle import features_oracle as FO
    g = np.load(os.path.join(ROOT, "tests", "golden", "features_small.npz"))
    N = int(g["chrom_range"][-1, 1]) - 1
    clusters = [g["members"][a:b] for a, b in zip(g["offsets"][:-1], g["offsets"][1:])]
    np.testing.assert_array_equal(FO.edgelist2adj(clusters, N), g["edge_adj"])
    c2n = {i: int(v) for i, v in enumerate(g["cool2node"]) if v > 0}
    n2c = {i: int(g["node2chrom"][i]) for i in range(1, N + 1)}
    intra, inter = FO.pixels2adj(g["bin1"], g["bin2"], g["count"], c2n, n2c, N)
    np.testing.assert_array_equal(intra, g["intra"])
    np.testing.assert_array_equal(inter, g["inter"])
