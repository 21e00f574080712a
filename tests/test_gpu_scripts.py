"""End-to-end run of the three entry points (scripts/main.py, denoise_contact.py, predict_multiway.py) on a
small synthetic SPRITE-like data set written in the reference's on-disk formats."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def workdir(tmp_path_factory):
    from matcha_b200.synthetic import make_dataset, write_temp_dir
    base = tmp_path_factory.mktemp("matcha_run")
    code = base / "Code"
    code.mkdir()
    ds = make_dataset((["chr21", "chr22"], 1_000_000, 64), kmers_per_size=6000, seed=3)
    cfg = write_temp_dir(ds, str(base / "Temp"))
    cfg["temp_dir"] = "../Temp"
    json.dump(cfg, open(code / "config.JSON", "w"))
    return base, ds


def _run(script, cwd, *args, env=None):
    e = dict(os.environ, PYTHONPATH=os.path.join(ROOT, "scripts") + os.pathsep + ROOT, **(env or {}))
    r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", script), *args], cwd=cwd, env=e, capture_output=True,
                       text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    return r.stdout


def test_main_trains_and_writes_the_reference_outputs(workdir):
    base, ds = workdir
    out = _run("main.py", str(base / "Code"), env={"MATCHA_EPOCHS1": "1", "MATCHA_EPOCHS2": "3",
                                                    "MATCHA_STEPS_PER_EPOCH": "120", "MATCHA_BATCH": "96"})
    emb = np.load(base / "embeddings.npy")
    assert emb.shape == (ds["N"], 64) and np.isfinite(emb).all()
    assert (base / "Temp" / "model.chkpt").exists() and (base / "Temp" / "model2load").exists()
    ck = torch.load(base / "Temp" / "model.chkpt", weights_only=False)
    assert "model_link" in ck and "epoch" in ck                              # main.py:316-318
    assert "node_embedding.Embedding_Linear0.tied weight_0" in ck["model_link"]
    auprs = [float(l.split("aupr: all ")[1].split(" ")[0]) for l in out.splitlines() if "Validation-hyper" in l]
    assert len(auprs) == 4
    # phase 2 learns to separate positives from corrupted tuples (the reference reaches AUPR ~0.73 after 1 epoch, 1:3 classes)
    assert max(auprs[1:]) > 0.45, out[-2000:]


def test_denoise_and_predict_multiway(workdir):
    base, ds = workdir
    code = str(base / "Code")
    _run("denoise_contact.py", code)
    for c in ds["chroms"]:
        m = np.load(base / ("%s_denoise.npy" % c))
        assert m.shape[0] == m.shape[1] and np.isfinite(m).all() and m.min() >= 0 and m.max() <= 1.0 + 1e-6
    px = np.load(base / "denoised_pixels.npz")
    n_pairs = sum(n * (n + 1) // 2 for n in ds["nums"])
    assert len(px["bin1_id"]) == len(px["bin2_id"]) == len(px["balanced"]) == n_pairs
    # predict_multiway on tuples of mixed sizes: each batch padded to its own longest tuple
    lines, tuples = [], []
    rng = np.random.default_rng(0)
    for k in (2, 3, 5, 4, 2):
        ids = np.sort(rng.choice(np.arange(1, ds["N"] + 1), size=k, replace=False))
        tuples.append(ids)
        names = []
        for i in ids:
            c = int(np.searchsorted(ds["chrom_range"][:, 1], i, side="right"))
            names.append("%s:%d" % (ds["chroms"][c], (i - ds["chrom_range"][c, 0]) * ds["res"] + 17))
        lines.append("\t".join(names))
    (base / "Code" / "tuples.txt").write_text("\n".join(lines) + "\n")
    _run("predict_multiway.py", code, "-i", "tuples.txt", "-o", "out.txt")
    got = np.loadtxt(base / "Code" / "out.txt")
    assert got.shape == (5,) and ((got > 0) & (got < 1)).all()
    sys.path.insert(0, os.path.join(ROOT, "scripts"))
    import Modules  # noqa: F401
    model = torch.load(base / "Temp" / "model2load", map_location="cuda", weights_only=False)
    model.eval()
    x = np.zeros((5, 5), dtype=np.int64)
    for i, t in enumerate(tuples):
        x[i, :len(t)] = t
    with torch.no_grad():
        want = torch.sigmoid(model(torch.from_numpy(x).cuda())).cpu().numpy().reshape(-1)
    np.testing.assert_allclose(got, want, rtol=1e-5, atol=1e-6)


def test_generate_kmers_script_matches_the_oracle(tmp_path):
    """scripts/generate_kmers.py on an edge_list.npy written as process.py:87 writes it (object array of lists)."""
    from oracle import kmer_oracle as KO
    rng = np.random.default_rng(4)
    clusters = []
    for _ in range(1500):
        size = int(min(28, 2 + rng.geometric(0.3)))
        a = int(rng.integers(1, 600))
        clusters.append(sorted(set(int(np.clip(a + rng.integers(-10, 11), 1, 700)) for _ in range(size))))
    code, temp = tmp_path / "Code", tmp_path / "Temp"
    code.mkdir(); temp.mkdir()
    arr = np.empty(len(clusters), dtype=object)
    for i, c in enumerate(clusters):
        arr[i] = c
    np.save(temp / "edge_list.npy", arr, allow_pickle=True)
    cfg = {"temp_dir": "../Temp", "max_cluster_size": 25, "k-mer_size": [2, 3, 4], "min_distance": 1, "min_freq_cutoff": 2}
    json.dump(cfg, open(code / "config.JSON", "w"))
    out = _run("generate_kmers.py", str(code))
    assert "Quick summarize" in out
    for k in (2, 3, 4):
        rows = np.load(temp / ("all_%d_counter.npy" % k))
        freq = np.load(temp / ("all_%d_freq_counter.npy" % k))
        want_rows, want_freq = KO.count_kmers(clusters, k, 1, 25, 2)
        assert rows.shape == want_rows.shape and (rows == want_rows).all() and (freq == want_freq).all(), k
