"""Cross-GPU exchange for the fit statistics.

One process per GPU (torch.distributed, NCCL over NVLink5/NVSwitch).  Rows are
data-parallel; the only exchanges are
  * moments: one all-reduce of 5 doubles per column (engine.Moments.allreduce)
  * group-by tables: (key, size[, payload]) rows are routed to owner =
    mix(key) % world with ONE all-to-all per table, merged by the owner (exact
    global aggregates over disjoint keys), and the merged shards are
    all-gathered so every rank builds the identical vocabulary / stats table.
This replaces the reference's dask tree reduction over TCP/UCX and its shared
-filesystem "broadcast" (nvtabular/ops/categorify.py:1399-1540, 1627-1643).
On CPU test runs the same code path is exercised with the gloo backend
(tests/test_dist_cpu.py) with torch ops standing in for the device kernels.
"""
from typing import Optional

import numpy as np
import torch


def world():
    import os
    import torch.distributed as dist
    if os.environ.get("NVTB_DISABLE_DIST"):      # single-process reference fits inside a rank
        return 1, 0
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(), dist.get_rank()
    return 1, 0


def exchange_by_owner(keys, sizes, vals, perm, counts):
    """all-to-all of rows already grouped by owner (perm/counts from
    nvtb_partition_by_owner).  Returns the rows this rank owns."""
    import torch.distributed as dist
    dev = keys.device
    send_counts = torch.tensor(counts, dtype=torch.int64, device=dev)
    recv_counts = torch.empty_like(send_counts)
    dist.all_to_all_single(recv_counts, send_counts)
    rc = [int(x) for x in recv_counts.cpu().tolist()]
    total = sum(rc)

    def a2a(t, width=1):
        out = torch.empty((total,) + tuple(t.shape[1:]), dtype=t.dtype, device=dev)
        dist.all_to_all_single(out, t.contiguous(),
                               output_split_sizes=rc, input_split_sizes=list(counts))
        return out

    rk = a2a(keys[perm] if perm is not None else keys)
    rs = a2a(sizes[perm] if perm is not None else sizes)
    rv = a2a(vals[perm] if perm is not None else vals) if vals is not None else None
    return rk, rs, rv


def allgather_var(t: torch.Tensor):
    """all-gather of a variable-length (dim 0) tensor -> concatenation in rank order."""
    import torch.distributed as dist
    w, _ = world()
    dev = t.device
    n_local = torch.tensor([t.shape[0]], dtype=torch.int64, device=dev)
    n_all = [torch.empty_like(n_local) for _ in range(w)]
    dist.all_gather(n_all, n_local)
    n_all = [int(x.item()) for x in n_all]
    mx = max(max(n_all), 1)
    pad = torch.zeros((mx,) + tuple(t.shape[1:]), dtype=t.dtype, device=dev)
    pad[: t.shape[0]] = t
    out = torch.empty((w * mx,) + tuple(t.shape[1:]), dtype=t.dtype, device=dev)
    dist.all_gather_into_tensor(out, pad)
    return torch.cat([out[r * mx: r * mx + n_all[r]] for r in range(w)])


def global_merge(agg, engine=None):
    """engine.HashAgg -> (keys, sizes, vals|None, null_size, null_vals|None),
    globally merged and identical on every rank.  `engine` is the kernel provider
    (nvtabular_b200.engine); the gloo/CPU tests inject a stand-in to exercise the
    exchange plumbing without a GPU."""
    if engine is None:
        from . import engine
    import torch.distributed as dist
    keys, sizes, vals, null_size, null_vals = agg.export()
    w, _ = world()
    if w == 1:
        return keys, sizes, vals, null_size, null_vals
    perm, counts = engine.partition_by_owner(keys, w)
    sk = engine.gather_i64(keys, perm)
    ss = engine.gather_i64(sizes, perm)
    sv = None
    if vals is not None:
        width = agg.n_agg * 4
        sv = engine.gather_f64_rows(vals.reshape(-1, width), perm, width)
    rk, rs, rv = exchange_by_owner(sk, ss, sv, None, counts)
    owner = engine.HashAgg(agg.n_agg, capacity_hint=max(rk.numel(), 1))
    owner.merge(rk, rs, rv.reshape(-1) if rv is not None else None)
    ok, os_, ov, _, _ = owner.export()
    all_k = allgather_var(ok)
    all_s = allgather_var(os_)
    all_v = allgather_var(ov) if ov is not None else None
    dev = keys.device
    ns = torch.tensor([null_size], dtype=torch.int64, device=dev)
    dist.all_reduce(ns, op=dist.ReduceOp.SUM)
    nv = None
    if null_vals is not None:
        t = torch.tensor(null_vals, dtype=torch.float64, device=dev)
        sums = t[:, 0:2].contiguous()
        mn = torch.nan_to_num(t[:, 2], nan=float("inf")).contiguous()
        mx = torch.nan_to_num(t[:, 3], nan=float("-inf")).contiguous()
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
        dist.all_reduce(mn, op=dist.ReduceOp.MIN)
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        mn[torch.isinf(mn)] = float("nan")
        mx[torch.isinf(mx)] = float("nan")
        nv = torch.cat([sums, mn[:, None], mx[:, None]], dim=1).cpu().numpy()
    return all_k, all_s, all_v, int(ns.item()), nv
