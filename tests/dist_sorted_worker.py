"""gloo worker (any world size) for tests/test_dist_cpu.py::test_sorted_exchange_world_sizes: runs the
product's key-range exchange of sorted packed pairs (nvtabular_b200.dist.global_merge_sorted) with
numpy stand-ins for the device primitives and writes each rank's result to disk.  Shards are
deliberately unequal (one rank is empty at world >= 3) so that padding segments, empty runs in the
merge tree and the quantile splitters of live ranks only are exercised."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _u(t):
    return t.numpy().view(np.uint64)


class PackedAgg:
    def __init__(self, keys_i32, nulls):
        u, c = np.unique(keys_i32.astype(np.int64), return_counts=True)
        uk = (u + (1 << 31)).astype(np.uint64)
        self.pairs = torch.from_numpy(((uk << np.uint64(32)) | c.astype(np.uint64)).view(np.int64))
        self.nulls = nulls

    def size(self):
        return self.pairs.numel(), self.nulls

    def export_packed(self, device=None):
        return self.pairs.clone()


class SortedEngine:
    @staticmethod
    def pairs_lower_bounds(pairs, bounds):
        k = _u(pairs) >> np.uint64(32)
        return torch.from_numpy(np.searchsorted(k, bounds.numpy().astype(np.uint64), side="left").astype(np.int64))

    @staticmethod
    def pairs_merge(a, b):
        w = np.concatenate([_u(a), _u(b)])
        k, c = w >> np.uint64(32), w & np.uint64(0xFFFFFFFF)
        uk, inv = np.unique(k, return_inverse=True)
        cs = np.zeros(len(uk), dtype=np.uint64)
        np.add.at(cs, inv, c)
        return torch.from_numpy(((uk << np.uint64(32)) | cs).view(np.int64))

    @staticmethod
    def radix_sort(data, lo_bit=0, hi_bit=None, descending=False):
        w = _u(data)
        f = ((w >> np.uint64(lo_bit)) & np.uint64((1 << (hi_bit - lo_bit)) - 1)).astype(np.int64)
        order = np.argsort(-f if descending else f, kind="stable")
        return torch.from_numpy(w[order].view(np.int64))

    @staticmethod
    def segment_copy(src, dst, seg_src, seg_dst):
        ss, sd = seg_src.tolist(), seg_dst.tolist()
        for j in range(len(sd)):
            if sd[j] >= 0:
                dst[sd[j]: sd[j] + ss[j + 1] - ss[j]] = src[ss[j]: ss[j + 1]]


def main():
    out_dir = sys.argv[1]
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    from nvtabular_b200.dist import global_merge_sorted
    rng = np.random.default_rng(500 + rank)
    cols, local = [], []
    for c in range(3):
        n = 0 if (world >= 3 and rank == 1 and c == 0) else 4000 + 1500 * rank + 700 * c
        kk = (rng.integers(0, 2500 * (c + 1), n) * 2654435761 % (1 << 32) - (1 << 31)).astype(np.int32)
        if c == 2:                                   # heavy hitters: counts far above the rest
            kk[: n // 3] = 12345
        cols.append(PackedAgg(kk, 2 + rank + c))
        local.append(kk.tolist())
    res = []
    for ordered, nsz in global_merge_sorted(cols, engine=SortedEngine, device=torch.device("cpu")):
        w = _u(ordered)
        res.append({"keys": ((w >> np.uint64(32)).astype(np.int64) - (1 << 31)).tolist(),
                    "sizes": (w & np.uint64(0xFFFFFFFF)).astype(np.int64).tolist(), "null": nsz})
    with open(os.path.join(out_dir, f"rank{rank}.json"), "w") as f:
        json.dump({"sorted": res, "local": local}, f)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
