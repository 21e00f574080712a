"""Denoise intra-chromosomal contact maps with a trained model (interface of Code/denoise_contact.py).

Reads ./config.JSON, `temp_dir/model2load`, `temp_dir/chrom_range.npy`, `temp_dir/intra_adj.npy`; for every
chromosome scores ALL bin pairs (i, j), j >= i + min_distance, in `generate_pair_wise` order
(denoise_contact.py:67-74) with the streaming GPU pair scorer -- no pair list is built on the host -- then
applies the reference's normalisation (denoise_contact.py:160-192) ON THE DEVICE (matcha_b200/denoise.py) and writes
`../<chrom>_denoise.npy` plus `../denoised_pixels.npz` (bin1_id, bin2_id, balanced: the datasets of the reference's HDF5
group; `h5py` / plotting libraries are optional and absent from this image).

Launched with torchrun, every chromosome's pairs shard by contiguous PAIR RANGE over all ranks (no communication while
scoring); the packed score slices are then gathered over NCCL to the chromosome's owner rank (chromosome index mod world),
which runs the post-processing and writes that chromosome's files.
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import Modules  # noqa: F401,E402  (needed by torch.load of the whole-module pickle)
from utils import get_config  # noqa: E402

from matcha_b200.denoise import QuantileUniform, denoise_matrix  # noqa: E402
from matcha_b200.parallel import init_from_env, range_shard  # noqa: E402
from matcha_b200.scorer import PairScorer, pair_count, pair_index_to_ij  # noqa: E402


def gather_pair_scores(part, total, rank, world, owner):
    """Packed scores of one chromosome: rank r holds pairs range_shard(total, r, world); returns the full vector on
    `owner` (None elsewhere).  One NCCL gather of equal-length (padded) slices."""
    if world == 1:
        return part
    import torch.distributed as dist
    width = (total + world - 1) // world + 1
    send = torch.zeros(width, dtype=torch.float32, device=part.device)
    send[:part.numel()] = part
    bufs = [torch.empty_like(send) for _ in range(world)] if rank == owner else None
    dist.gather(send, bufs, dst=owner)
    if rank != owner:
        return None
    full = torch.empty(total, dtype=torch.float32, device=part.device)
    for r in range(world):
        b, e = range_shard(total, r, world)
        full[b:e] = bufs[r][:e - b]
    return full


def main():
    config = get_config()
    min_dis, temp_dir = config["min_distance"], config["temp_dir"]
    rank, world, local = init_from_env()
    dev = "cuda:%d" % local
    torch.cuda.set_device(local)
    chrom_range = np.load(os.path.join(temp_dir, "chrom_range.npy"))
    model = torch.load(os.path.join(temp_dir, "model2load"), map_location=dev, weights_only=False)
    model.eval()
    origin = np.load(os.path.join(temp_dir, "intra_adj.npy"), mmap_mode="r")
    scorer = PairScorer(model)
    chrom_name = config["chrom_list"]
    bin1, bin2, balanced = [], [], []
    for i in range(len(chrom_name)):
        lo, hi = int(chrom_range[i, 0]), int(chrom_range[i, 1])
        n = hi - lo
        total = pair_count(lo, hi, min_dis)
        owner = i % world
        b, e = range_shard(total, rank, world)
        part = scorer.score_range(lo, hi, min_dis, b, e, sigmoid=True)                    # denoise_contact.py:153-155
        proba = gather_pair_scores(part, total, rank, world, owner)
        if rank != owner:
            continue
        block = torch.from_numpy(np.ascontiguousarray(origin[lo - 1:hi - 1, lo - 1:hi - 1], dtype=np.float32)).to(dev)   # :160
        my, pix = denoise_matrix(proba, block, n, min_dis, QuantileUniform(n_quantiles=1000), want_pixels=True)   # :162-189,205
        np.save("../%s_denoise.npy" % chrom_name[i], my.cpu().numpy())
        ii, jj = pair_index_to_ij(np.arange(total), lo, hi, min_dis)
        bin1.append(ii - 1); bin2.append(jj - 1); balanced.append(pix.cpu().numpy())     # :152-153,205
        print("%s: %d pairs scored" % (chrom_name[i], total))
    if bin1:
        suffix = "" if world == 1 else ".rank%d" % rank
        np.savez("../denoised_pixels%s.npz" % suffix, bin1_id=np.concatenate(bin1), bin2_id=np.concatenate(bin2),
                 balanced=np.concatenate(balanced))
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
