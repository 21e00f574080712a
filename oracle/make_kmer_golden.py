"""Golden vectors for k-mer enumeration / counting, produced by EXECUTING the reference's own ``build_dict``
(authoring container only; `/root/reference` does not exist on the GPU box).

    python oracle/make_kmer_golden.py           # writes tests/golden/kmer_small.npz

``Code/generate_kmers.py`` is a script (its body reads ./config.JSON and ``edge_list.npy`` at import, ``:73-84``), so its
one function, ``build_dict`` (``:8-69``), is lifted verbatim with ``ast`` and run in a namespace that provides the module
globals it reads (``node2usefulindex``, ``new_data``, ``min_dis``, ``min_freq_cutoff``) built exactly as the script body
builds them (``:88-97``).  The process pool around it (``:103-132``) only concatenates per-node results.
"""
import ast
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import kmer_oracle as KO  # noqa: E402

REF = "/root/reference/Code/generate_kmers.py"


def lift_build_dict():
    tree = ast.parse(open(REF).read())
    fns = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "build_dict"]
    assert len(fns) == 1
    ns = {}
    exec("from itertools import combinations\nfrom collections import Counter\nimport numpy as np\n"
         "def tqdm(x, *a, **k):\n    return x\n", ns)
    exec(compile(ast.Module(body=fns, type_ignores=[]), "generate_kmers.py(lifted)", "exec"), ns)
    return ns


def reference_kmers(ns, clusters, node_num, k, min_dis, max_size, min_freq):
    """The script body of generate_kmers.py:86-141 for one k, without the process pool."""
    new_data = [np.array(d) for d in clusters if (len(d) >= k) & (len(d) <= max_size)]            # :88-91
    node2usefulindex = [[] for _ in range(node_num)]                                                # :93-96
    for i, datum in enumerate(new_data):
        for n in datum:
            node2usefulindex[n].append(i)
    ns.update(new_data=new_data, node2usefulindex=node2usefulindex, min_dis=min_dis, min_freq_cutoff=min_freq)
    rows, freqs = [], []
    for batch in np.array_split(np.arange(node_num).astype("int"), max(1, node_num // 50)):         # :110-114
        _, temp, temp_freq = ns["build_dict"](k, batch)
        if len(temp) > 0:
            rows.append(np.asarray(temp)); freqs.append(np.asarray(temp_freq))
    if not rows:
        return np.zeros((0, k), dtype=np.int64), np.zeros((0,), dtype=np.int64)
    return KO.sort_rows(np.concatenate(rows, axis=0), np.concatenate(freqs, axis=0))


def toy_clusters(seed=0, n_clusters=900, node_num=400):
    """SPRITE-like clusters: unique ascending node ids, sizes 2..30 (some above max_cluster_size), local + far members."""
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n_clusters):
        size = int(min(30, 2 + rng.geometric(0.3)))
        anchor = int(rng.integers(1, node_num))
        ids = {anchor}
        while len(ids) < size:
            if rng.random() < 0.85:
                ids.add(int(np.clip(anchor + rng.integers(-9, 10), 1, node_num - 1)))
            else:
                ids.add(int(rng.integers(1, node_num)))
        out.append(sorted(ids))
    return out


def main():
    ns = lift_build_dict()
    clusters = toy_clusters()
    node_num = 400
    flat = np.concatenate([np.asarray(c, dtype=np.int64) for c in clusters])
    offsets = np.concatenate([[0], np.cumsum([len(c) for c in clusters])]).astype(np.int64)
    out = {"members": flat, "offsets": offsets, "node_num": np.int64(node_num)}
    cases = [(2, 0, 25, 2), (3, 0, 25, 2), (4, 0, 25, 2), (5, 0, 12, 2), (3, 2, 25, 2), (2, 5, 25, 3), (4, 1, 25, 1)]
    out["cases"] = np.asarray(cases, dtype=np.int64)
    for ci, (k, min_dis, max_size, min_freq) in enumerate(cases):
        rows, freq = reference_kmers(ns, clusters, node_num, k, min_dis, max_size, min_freq)
        mine_rows, mine_freq = KO.count_kmers(clusters, k, min_dis, max_size, min_freq)
        assert rows.shape == mine_rows.shape and (rows == mine_rows).all() and (freq == mine_freq).all(), (k, min_dis)
        out[f"rows/{ci}"], out[f"freq/{ci}"] = rows, freq
        print(f"k={k} min_dis={min_dis} max_size={max_size} min_freq={min_freq}: {len(freq)} k-mers, max freq {freq.max() if len(freq) else 0}")
    path = os.path.join(os.path.dirname(HERE), "tests", "golden", "kmer_small.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
