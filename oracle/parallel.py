"""Partition-parallel runner for the oracle — TEST / BENCH INFRASTRUCTURE.

The reference without a Dask client computes with scheduler="synchronous"
(nvtabular/workflow/workflow.py:74; ops/categorify.py:1910-1916) = one core;
with a LocalCluster it runs one partition per worker and merges partial
statistics with _mid_level_groupby / _tree_node_moments.  `run_criteo_workflow`
does exactly that split with a fork-based process pool (partials and vocabularies
travel through files, like the reference's on_host spill and unique.<col>.parquet), so bench.py can time
the reference's CPU path on all host cores (`--impl reference`) or on one
(`cpu_baseline`)."""
import multiprocessing as mp
import os
import pickle
import shutil
import tempfile
import time
from typing import List

import numpy as np
import pandas as pd

from . import categorify as _cat
from .moments import chunkwise_moments, finalize_moments, tree_node_moments
from .normalize import fill_missing, normalize_transform

_PARTS: List[pd.DataFrame] = []      # inherited by forked workers (no pickling of inputs)
_STATE = {}
_VOCAB_CACHE = {}                    # per worker process, like the reference's worker cache


def _fit_part(i):
    """one partition: _top_level_groupby per column (written next to the other partials, as the
    reference spills them with to_arrow / on_host, categorify.py:1036-1049) + chunk-wise moments"""
    df = _PARTS[i]
    cats, conts, tmp = _STATE["cats"], _STATE["conts"], _STATE["tmp"]
    for c in cats:
        _cat.top_level_groupby(df, [c], True)[0].to_pickle(os.path.join(tmp, f"gb.{c}.{i}.pkl"))
    return chunkwise_moments(fill_missing(df[conts], conts, 0)) if conts else None


def _merge_col(c):
    """one column: _mid_level_groupby over every partition's partial + _write_uniques
    (categorify.py:1054-1337); the vocabulary goes to a file, like unique.<col>.parquet"""
    tmp, nparts = _STATE["tmp"], _STATE["nparts"]
    parts = [pd.read_pickle(os.path.join(tmp, f"gb.{c}.{i}.pkl")) for i in range(nparts)]
    vocab = _cat.write_uniques(_cat.mid_level_groupby(parts, [c]), [c])
    with open(os.path.join(tmp, f"vocab.{c}.pkl"), "wb") as f:
        pickle.dump(vocab, f, protocol=pickle.HIGHEST_PROTOCOL)
    return c


def _vocab(c):
    v = _VOCAB_CACHE.get((_STATE["tmp"], c))
    if v is None:                      # fetch_table_data's per-worker cache, categorify.py:1627-1643
        with open(os.path.join(_STATE["tmp"], f"vocab.{c}.pkl"), "rb") as f:
            v = _VOCAB_CACHE[(_STATE["tmp"], c)] = pickle.load(f)
    return v


def _transform_part(args):
    i, means, stds = args
    df = _PARTS[i]
    cats, conts = _STATE["cats"], _STATE["conts"]
    out = {}
    for c in cats:
        out[c] = _cat.categorify_encode(df, c, _vocab(c))
    if conts:
        nd = normalize_transform(fill_missing(df[conts], conts, 0), conts, means, stds)
        for c in conts:
            out[c] = nd[c].to_numpy()
    # the result stays in the worker (as the reference's workers keep their partitions);
    # return a checksum so the work cannot be optimised away
    return int(sum(int(np.asarray(v[:16]).sum()) for v in out.values() if len(v)))


def run_criteo_workflow(df: pd.DataFrame, cats: List[str], conts: List[str], workers: int = 1):
    """Categorify(cats) + FillMissing + Normalize(conts): fit then transform, as a LocalCluster
    with `workers` processes would run it: per-partition partials in parallel, the per-column
    merge + vocabulary write in parallel over columns, the encode in parallel over partitions
    with the vocabularies read from files by each worker.  ONE pool serves all phases.
    Returns (seconds_fit, seconds_transform, vocabs, means, stds)."""
    global _PARTS
    n = len(df)
    workers = max(1, min(workers, max(1, n // 50_000)))
    chunk = -(-n // workers)
    _PARTS = [df.iloc[i:i + chunk] for i in range(0, n, chunk)]
    base = "/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else None
    tmp = tempfile.mkdtemp(prefix="nvtb_oracle_", dir=base)
    _STATE.clear()
    _STATE.update(cats=cats, conts=conts, tmp=tmp, nparts=len(_PARTS))
    pool = mp.get_context("fork").Pool(workers) if workers > 1 else None
    pmap = pool.map if pool else (lambda f, xs: [f(x) for x in xs])
    try:
        idx = list(range(len(_PARTS)))
        t0 = time.perf_counter()
        moms = pmap(_fit_part, idx)
        # largest vocabularies first: the phase ends with its slowest column
        pmap(_merge_col, list(cats))
        means, stds = {}, {}
        if conts:
            stats = finalize_moments(tree_node_moments(moms))
            means = {c: float(stats["mean"].loc[c]) for c in conts}
            stds = {c: float(stats["std"].loc[c]) for c in conts}
        t1 = time.perf_counter()
        _ = pmap(_transform_part, [(i, means, stds) for i in idx])
        t2 = time.perf_counter()
        vocabs = {c: _vocab(c) for c in cats}
    finally:
        if pool:
            pool.close()
            pool.join()
        shutil.rmtree(tmp, ignore_errors=True)
        _VOCAB_CACHE.clear()
    return t1 - t0, t2 - t1, vocabs, means, stds


# ---------------------------------------------------------------------------------------
# C5: HashBucket over many key columns (nvtabular/ops/hash_bucket.py:86-100)
# ---------------------------------------------------------------------------------------
_HB = {}


def _hb_task(args):
    j, lo, hi = args
    from .hashing import hash_bucket
    out = hash_bucket(_HB["cols"][j][lo:hi], _HB["nb"])
    return int(out[:16].sum())          # the labels stay in the worker; a checksum comes back


def run_hashbucket(cols: List[np.ndarray], num_buckets: int, workers: int = 1) -> float:
    """hash_series(col) % num_buckets for every column, row ranges spread over `workers` forked
    processes (the arrays are inherited, not pickled).  Returns seconds."""
    _HB.clear()
    _HB.update(cols=cols, nb=num_buckets)
    n = len(cols[0])
    workers = max(1, workers)
    per = max(1, -(-workers // len(cols)))                 # row ranges per column
    chunk = -(-n // per)
    tasks = [(j, lo, min(n, lo + chunk)) for j in range(len(cols)) for lo in range(0, n, chunk)]
    pool = mp.get_context("fork").Pool(workers) if workers > 1 else None
    try:
        t0 = time.perf_counter()
        if pool:
            pool.map(_hb_task, tasks, chunksize=1)
        else:
            for t in tasks:
                _hb_task(t)
        return time.perf_counter() - t0
    finally:
        if pool:
            pool.close()
            pool.join()


# ---------------------------------------------------------------------------------------
# C4: JoinGroupby + TargetEncoding on a ratings table
# (nvtabular/ops/join_groupby.py:140-217, target_encoding.py:171-439)
# ---------------------------------------------------------------------------------------
_ML = {}


def _ml_fit_part(i):
    from .groupby import _top_level
    p = _PARTS[i]
    out = {}
    for g in _ML["groups"]:
        out[("jg",) + tuple(g)] = _top_level(p, g, _ML["conts"], _ML["stats"])
        out[("te",) + tuple(g)] = _top_level(p, g, _ML["targets"], ["count", "sum"])
        fg = ["__fold__"] + g
        out[("te",) + tuple(fg)] = _top_level(p, fg, _ML["targets"], ["count", "sum"])
    return out, chunkwise_moments(p[_ML["targets"]])


def _ml_transform_part(i):
    from .groupby import join_groupby_transform, te_transform_part
    p = _PARTS[i]
    a = join_groupby_transform(p, _ML["groups"], _ML["jg_tables"])
    b = te_transform_part(p, _ML["groups"], _ML["te_tables"], _ML["y_mean"], _ML["targets"], _ML["kfold"],
                          _ML["p_smooth"])
    return float(a.iloc[:16].to_numpy(dtype="float64").sum() + b.iloc[:16].to_numpy(dtype="float64").sum())


def run_movielens_workflow(df: pd.DataFrame, workers: int = 1, kfold: int = 5, p_smooth: float = 20.0,
                           fold_seed: int = 42):
    """["userId","movieId",["userId","movieId"]] >> JoinGroupby(cont_cols=["rating"],
    stats=[count,sum,mean,std]) + TargetEncoding("rating", kfold, p_smooth): per-partition partial
    group-bys in parallel, merged by the parent (_bottom_level), transform in parallel over the
    partitions (the tables reach the forked workers through module state).  Returns
    (seconds_fit, seconds_transform)."""
    global _PARTS
    from .groupby import _bottom_level, _make_name, add_fold
    n = len(df)
    workers = max(1, min(workers, max(1, n // 50_000)))
    chunk = -(-n // workers)
    parts = []
    for i in range(0, n, chunk):
        p = df.iloc[i:i + chunk].reset_index(drop=True)
        p["__fold__"] = add_fold(len(p), kfold, fold_seed)
        parts.append(p)
    _PARTS = parts
    groups = [["userId"], ["movieId"], ["userId", "movieId"]]
    stats = ["count", "sum", "mean", "std"]
    _ML.clear()
    _ML.update(groups=groups, conts=["rating"], targets=["rating"], stats=stats, kfold=kfold, p_smooth=p_smooth)
    idx = list(range(len(parts)))
    t0 = time.perf_counter()
    pool = mp.get_context("fork").Pool(workers) if workers > 1 else None
    try:
        res = pool.map(_ml_fit_part, idx) if pool else [_ml_fit_part(i) for i in idx]
    finally:
        if pool:
            pool.close()
            pool.join()
    jg_tables, te_tables = {}, {}
    for g in groups:
        jg_tables[_make_name(*g)] = _bottom_level([r[0][("jg",) + tuple(g)] for r in res], g, ["rating"], stats)
        te_tables[_make_name(*g)] = _bottom_level([r[0][("te",) + tuple(g)] for r in res], g, ["rating"],
                                                  ["count", "sum"])
        fg = ["__fold__"] + g
        te_tables[_make_name(*fg)] = _bottom_level([r[0][("te",) + tuple(fg)] for r in res], fg, ["rating"],
                                                   ["count", "sum"])
    stats_m = finalize_moments(tree_node_moments([r[1] for r in res]))
    _ML.update(jg_tables=jg_tables, te_tables=te_tables, y_mean={"rating": float(stats_m["mean"].loc["rating"])})
    t1 = time.perf_counter()
    pool = mp.get_context("fork").Pool(workers) if workers > 1 else None     # forked AFTER the tables exist
    try:
        _ = pool.map(_ml_transform_part, idx) if pool else [_ml_transform_part(i) for i in idx]
    finally:
        if pool:
            pool.close()
            pool.join()
    t2 = time.perf_counter()
    return t1 - t0, t2 - t1
