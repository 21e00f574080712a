"""Denoise intra-chromosomal contact maps with a trained model (interface of Code/denoise_contact.py).

Reads ./config.JSON, `temp_dir/model2load`, `temp_dir/chrom_range.npy`, `temp_dir/intra_adj.npy`; for every
chromosome scores ALL bin pairs (i, j), j >= i + min_distance, in `generate_pair_wise` order
(denoise_contact.py:67-74) with the streaming GPU pair scorer -- no pair list is built on the host -- then
applies the reference's normalisation (denoise_contact.py:160-192) and writes `../<chrom>_denoise.npy`
plus `../denoised_pixels.npz` (bin1_id, bin2_id, balanced: the datasets of the reference's HDF5 group;
`h5py` / plotting libraries are optional and absent from this image).  Pair ranges shard across ranks
when launched with torchrun (rank 0 gathers nothing: each rank writes its own chromosomes).
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import Modules  # noqa: F401,E402  (needed by torch.load of the whole-module pickle)
from utils import get_config  # noqa: E402

from matcha_b200.denoise import denoise_matrix  # noqa: E402
from matcha_b200.parallel import init_from_env  # noqa: E402
from matcha_b200.scorer import PairScorer, pair_count, pair_index_to_ij  # noqa: E402


def main():
    config = get_config()
    min_dis, temp_dir = config["min_distance"], config["temp_dir"]
    rank, world, local = init_from_env()
    dev = "cuda:%d" % local
    chrom_range = np.load(os.path.join(temp_dir, "chrom_range.npy"))
    model = torch.load(os.path.join(temp_dir, "model2load"), map_location=dev, weights_only=False)
    model.eval()
    origin = np.load(os.path.join(temp_dir, "intra_adj.npy")).astype("float32")
    from sklearn.preprocessing import QuantileTransformer
    transformer = QuantileTransformer(n_quantiles=1000, output_distribution="uniform")
    scorer = PairScorer(model)
    chrom_name = config["chrom_list"]
    bin1, bin2, balanced = [], [], []
    for i in range(rank, len(chrom_name), world):
        lo, hi = int(chrom_range[i, 0]), int(chrom_range[i, 1])
        n = hi - lo
        total = pair_count(lo, hi, min_dis)
        proba = scorer.score_range(lo, hi, min_dis, sigmoid=True).cpu().numpy()         # denoise_contact.py:153-155
        ii, jj = pair_index_to_ij(np.arange(total), lo, hi, min_dis)
        ii, jj = ii - lo, jj - lo
        weight = origin[ii + lo - 1, jj + lo - 1]                                         # :160
        my = denoise_matrix(n, ii, jj, proba, weight, transformer)                        # :162-189
        np.save("../%s_denoise.npy" % chrom_name[i], my.astype("float32"))
        bin1.append(ii + lo - 1); bin2.append(jj + lo - 1); balanced.append(my[ii, jj])   # :152-153,205
        print("%s: %d pairs scored" % (chrom_name[i], total))
    if bin1:
        suffix = "" if world == 1 else ".rank%d" % rank
        np.savez("../denoised_pixels%s.npz" % suffix, bin1_id=np.concatenate(bin1), bin2_id=np.concatenate(bin2),
                 balanced=np.concatenate(balanced))


if __name__ == "__main__":
    main()
