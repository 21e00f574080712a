"""Score user-supplied multi-way interactions (interface of Code/predict_multiway.py).

    python predict_multiway.py -i tuples.txt -o output.txt

Each input line is tab-separated `chrom:coordinate` entries; they are mapped to bins through
`temp_dir/bin2node.npy` exactly as predict_multiway.py:23-60 does (floor to the resolution, dedupe, sort,
keep tuples with more than one bin).  Batches of 1e4 tuples are zero-padded to the longest tuple IN THE
BATCH -- the reference's padding rule, which the score depends on -- scored on the GPU and written with
np.savetxt after a sigmoid (predict_multiway.py:112-114).
"""
import argparse
import math
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import Modules  # noqa: F401,E402
from utils import get_config  # noqa: E402

from matcha_b200.scorer import score_tuples  # noqa: E402


def parse_file(filepath, temp_dir, chrom_list, res):
    bin2node = np.load(os.path.join(temp_dir, "bin2node.npy"), allow_pickle=True).item()
    final = []
    with open(filepath, "r") as f:
        for line in f:
            temp = []
            for info in line.strip().split("\t"):
                if not info:
                    continue
                try:
                    chrom, bin_ = info.split(":")
                except ValueError:
                    print(info)
                    raise EOFError
                if chrom not in chrom_list:
                    continue
                bin_ = int(math.floor(int(bin_) / res)) * res
                temp.append(bin2node["%s:%d" % (chrom, bin_)])        # KeyError on an unknown bin, as the reference
            temp = sorted(set(temp))
            if len(temp) > 1:
                final.append(temp)
    return final


def main():
    ap = argparse.ArgumentParser(description="predict multi-way interactions")
    ap.add_argument("-i", "--file", type=str)
    ap.add_argument("-o", "--output", type=str, default="./output.txt")
    args = ap.parse_args()
    if type(args.file) != str:
        print("invalid filepath")
        raise EOFError
    config = get_config()
    temp_dir, res, chrom_list = config["temp_dir"], config["resolution"], config["chrom_list"]
    samples = parse_file(args.file, temp_dir, chrom_list, res)
    dev = "cuda:%d" % torch.cuda.current_device()
    model = torch.load(os.path.join(temp_dir, "model2load"), map_location=dev, weights_only=False)
    outs = score_tuples(model, samples, batch_size=int(1e4), sigmoid=True)
    proba = np.concatenate([o.cpu().numpy() for _, o in sorted(outs, key=lambda t: t[0])], axis=0)
    np.savetxt(args.output, proba)


if __name__ == "__main__":
    main()
