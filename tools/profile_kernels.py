"""Launches each hot kernel a few times on synthetic Criteo-shaped columns so that
`ncu -k regex:<kernel>` can capture them without the data-generation noise.

    python tools/profile_kernels.py --rows 33554432 --which insert --cards 39043,100000,1766023
"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nvtabular_b200 import engine  # noqa: E402
from nvtabular_b200.column import Column, pack_validity  # noqa: E402
from nvtabular_b200.synth import power_law_ids, scatter_ids  # noqa: E402


def make_col(rows, k, seed, null_frac=0.02):
    g = torch.Generator(device="cuda")
    g.manual_seed(seed)
    keys = scatter_ids(power_law_ids(rows, k, g, "cuda"))
    valid = torch.rand(rows, generator=g, device="cuda") >= null_frac
    return Column(keys, pack_validity(valid))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=1 << 25)
    ap.add_argument("--which", default="insert,encode,moments,normalize")
    ap.add_argument("--cards", default="3,1543,39043,100000,1766023")
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--batches", type=int, default=2)
    a = ap.parse_args()
    which = a.which.split(",")
    cards = [int(x) for x in a.cards.split(",")]
    for i, k in enumerate(cards):
        col = make_col(a.rows, k, 100 + i)
        if "insert" in which or "encode" in which:
            agg = engine.HashAgg(0, capacity_hint=k)
            for _ in range(a.reps):
                agg.reset()
                for b in range(a.batches):        # a fit = several batches (sorted accumulators stage them)
                    agg.insert(col)
            # the vocabulary straight from the handle: the path Categorify.fit takes on one GPU
            v = engine.Vocab.build_from_agg(agg, key_bits=32, size_bound=a.rows * a.batches)
            torch.cuda.synchronize()
            print(f"card {k}: mode {agg.mode}, {v.n_kept} kept keys, null {v.null_size}")
        if "encode" in which:
            for _ in range(a.reps):
                out = v.encode(col, 1, 2, 3)
            torch.cuda.synchronize()
            del out
    if "moments" in which or "normalize" in which:
        g = torch.Generator(device="cuda"); g.manual_seed(7)
        cols = []
        for j in range(13):
            data = torch.randint(0, 1 << 20, (a.rows,), generator=g, device="cuda", dtype=torch.int32)
            valid = torch.rand(a.rows, generator=g, device="cuda") >= 0.03 * j
            c = Column(data, pack_validity(valid)); c.fill = 0.0
            cols.append(c)
        for _ in range(a.reps):
            m = engine.Moments(13); m.accumulate(cols)
            outs = engine.normalize_apply(cols, [1.0] * 13, [2.0] * 13)
        torch.cuda.synchronize()


if __name__ == "__main__":
    main()
