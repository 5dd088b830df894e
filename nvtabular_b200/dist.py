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


def all_gather_object(obj):
    """[obj of rank 0, obj of rank 1, ...] on every rank (host metadata only: fit modes,
    string dictionaries — never row data)."""
    import torch.distributed as dist
    w, _ = world()
    if w == 1:
        return [obj]
    out = [None] * w
    dist.all_gather_object(out, obj)
    return out


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


def global_merge_many(aggs, engine=None, owner_pool=None):
    """Cross-GPU merge of MANY keys-only tables (one per Categorify column) with a constant
    number of collectives: the per-column partials of all columns travel in ONE all-to-all
    (keys) + ONE (sizes), each owner merges its shard of every column, and the merged shards
    come back in ONE all-gather pair.  A 26-column fit costs ~8 NCCL calls instead of ~200,
    so the exchange is bound by NVLink bytes (O(#distinct keys)), not by launch latency.
    Returns [(keys, sizes, null_size)] per table, identical on every rank."""
    if engine is None:
        from . import engine
    import os
    import time
    import torch.distributed as dist
    w, rank = world()
    trace = bool(os.environ.get("NVTB_TRACE")) and w > 1
    marks = []

    def mark(name):
        if trace:
            torch.cuda.synchronize()
            marks.append((name, time.perf_counter()))

    mark("start")
    exported = [a.export() for a in aggs]            # (keys, sizes, None, null_size, None)
    mark("export")
    if w == 1:
        return [(k, s, ns) for (k, s, _, ns, _) in exported]
    nc = len(aggs)
    dev = exported[0][0].device
    # 1. group every column's rows by owner rank
    # (no host round trip per column: counts stay on the device until all 26 are queued)
    grouped = []
    if hasattr(engine, "partition_by_owner_async"):
        counts_dev = torch.zeros((nc, w), dtype=torch.int64, device=dev)
        for c, (k, s, _, _, _) in enumerate(exported):
            perm = engine.partition_by_owner_async(k, w, counts_dev[c])
            grouped.append((engine.gather_i64(k, perm), engine.gather_i64(s, perm)))
        counts_nc = counts_dev.cpu()
    else:                                   # stand-in engines of the CPU tests
        counts_nc = torch.zeros((nc, w), dtype=torch.int64)
        for c, (k, s, _, _, _) in enumerate(exported):
            perm, cnt = engine.partition_by_owner(k, w)
            grouped.append((engine.gather_i64(k, perm), engine.gather_i64(s, perm)))
            counts_nc[c] = torch.tensor(cnt, dtype=torch.int64)
    send_k = [[None] * nc for _ in range(w)]
    send_s = [[None] * nc for _ in range(w)]
    counts = counts_nc.t().contiguous()     # [owner rank, column]
    for c, (gk, gs) in enumerate(grouped):
        off = 0
        for r in range(w):
            n_rc = int(counts[r, c])
            send_k[r][c] = gk[off: off + n_rc]
            send_s[r][c] = gs[off: off + n_rc]
            off += n_rc
    mark("partition")
    sk = torch.cat([t for r in range(w) for t in send_k[r]])
    ss = torch.cat([t for r in range(w) for t in send_s[r]])
    mark("cat")
    # 2. exchange the count matrix, then keys and sizes
    cm_send = counts.to(dev).reshape(-1)
    cm_recv = torch.empty_like(cm_send)              # cm_recv[src, c] = rows src sends me for column c
    dist.all_to_all_single(cm_recv, cm_send)
    cm_recv_h = cm_recv.view(w, nc).cpu()
    in_split = [int(x) for x in counts.sum(dim=1).tolist()]
    out_split = [int(x) for x in cm_recv_h.sum(dim=1).tolist()]
    rk = torch.empty(sum(out_split), dtype=torch.int64, device=dev)
    rs = torch.empty(sum(out_split), dtype=torch.int64, device=dev)
    dist.all_to_all_single(rk, sk, output_split_sizes=out_split, input_split_sizes=in_split)
    dist.all_to_all_single(rs, ss, output_split_sizes=out_split, input_split_sizes=in_split)
    mark("all_to_all")
    # 3. owner merge per column (exact global sizes over disjoint keys)
    src_off = [0]
    for r in range(w):
        src_off.append(src_off[-1] + out_split[r])
    owned_k, owned_s, owners = [], [], []
    for c in range(nc):
        segs_k, segs_s = [], []
        for r in range(w):
            o = src_off[r] + int(cm_recv_h[r, :c].sum())
            n = int(cm_recv_h[r, c])
            segs_k.append(rk[o: o + n])
            segs_s.append(rs[o: o + n])
        ck, cs = torch.cat(segs_k), torch.cat(segs_s)
        # owner tables are pooled by the caller: creating a table (pinned mailbox, event,
        # device counters) costs far more than merging a shard into it
        if owner_pool is not None and c < len(owner_pool) and owner_pool[c] is not None:
            owner = owner_pool[c]
            owner.reset()
        else:
            owner = engine.HashAgg(0, capacity_hint=max(ck.numel(), 1))
            if owner_pool is not None:
                while len(owner_pool) <= c:
                    owner_pool.append(None)
                owner_pool[c] = owner
        owner.merge(ck, cs)
        owners.append(owner)
    # exports in a second loop: an export needs the merged size on the host, i.e. a sync on
    # that column's merge; issued right after each merge it would drain the stream 26 times
    for owner in owners:
        ok, os_, _, _, _ = owner.export()
        owned_k.append(ok)
        owned_s.append(os_)
    mark("owner_merge")
    # 4. all-gather the merged shards of all columns at once
    n_local = torch.tensor([t.numel() for t in owned_k], dtype=torch.int64, device=dev)
    n_all = torch.empty(w * nc, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(n_all, n_local)
    n_all_h = n_all.view(w, nc).cpu()
    tot = [int(x) for x in n_all_h.sum(dim=1).tolist()]
    mx = max(max(tot), 1)
    pk = torch.zeros(mx, dtype=torch.int64, device=dev)
    ps = torch.zeros(mx, dtype=torch.int64, device=dev)
    if tot[rank]:
        pk[: tot[rank]] = torch.cat(owned_k)
        ps[: tot[rank]] = torch.cat(owned_s)
    gk = torch.empty(w * mx, dtype=torch.int64, device=dev)
    gs = torch.empty(w * mx, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(gk, pk)
    dist.all_gather_into_tensor(gs, ps)
    ns = torch.tensor([e[3] for e in exported], dtype=torch.int64, device=dev)
    dist.all_reduce(ns, op=dist.ReduceOp.SUM)
    ns_h = ns.cpu().tolist()
    mark("allgather")
    out = []
    for c in range(nc):
        ks, szs = [], []
        for r in range(w):
            o = r * mx + int(n_all_h[r, :c].sum())
            n = int(n_all_h[r, c])
            ks.append(gk[o: o + n])
            szs.append(gs[o: o + n])
        out.append((torch.cat(ks), torch.cat(szs), int(ns_h[c])))
    mark("split")
    if trace and rank == 0:
        print("[nvtb trace] merge_many: " + ", ".join(
            f"{marks[i][0]} {1e3 * (marks[i][1] - marks[i - 1][1]):.2f} ms" for i in range(1, len(marks)))
              + f"; keys sent {sk.numel()}, owned {sum(t.numel() for t in owned_k)}", flush=True)
    return out
