"""Partition-parallel runner for the oracle — TEST / BENCH INFRASTRUCTURE.

The reference without a Dask client computes with scheduler="synchronous"
(nvtabular/workflow/workflow.py:74; ops/categorify.py:1910-1916) = one core;
with a LocalCluster it runs one partition per worker and merges partial
statistics with _mid_level_groupby / _tree_node_moments.  `run_criteo_workflow`
does exactly that split with a fork-based process pool, so bench.py can time
the reference's CPU path on all host cores (`--impl reference`) or on one
(`cpu_baseline`)."""
import multiprocessing as mp
import os
import time
from typing import List

import numpy as np
import pandas as pd

from . import categorify as _cat
from .moments import chunkwise_moments, finalize_moments, tree_node_moments
from .normalize import fill_missing, normalize_transform

_PARTS: List[pd.DataFrame] = []      # inherited by forked workers (no pickling of inputs)
_STATE = {}


def _fit_part(i):
    df = _PARTS[i]
    cats, conts = _STATE["cats"], _STATE["conts"]
    gbs = {c: _cat.top_level_groupby(df, [c], True)[0] for c in cats}
    mom = chunkwise_moments(fill_missing(df[conts], conts, 0)) if conts else None
    return gbs, mom


def _transform_part(i):
    df = _PARTS[i]
    cats, conts = _STATE["cats"], _STATE["conts"]
    vocabs, means, stds = _STATE["vocabs"], _STATE["means"], _STATE["stds"]
    out = {}
    for c in cats:
        out[c] = _cat.categorify_encode(df, c, vocabs[c])
    if conts:
        nd = normalize_transform(fill_missing(df[conts], conts, 0), conts, means, stds)
        for c in conts:
            out[c] = nd[c].to_numpy()
    # the result stays in the worker (as the reference's workers keep their partitions);
    # return a checksum so the work cannot be optimised away
    return int(sum(int(np.asarray(v[:16]).sum()) for v in out.values() if len(v)))


def run_criteo_workflow(df: pd.DataFrame, cats: List[str], conts: List[str], workers: int = 1):
    """Categorify(cats) + FillMissing + Normalize(conts): fit then transform.
    Returns (seconds_fit, seconds_transform, vocabs, means, stds)."""
    global _PARTS
    n = len(df)
    workers = max(1, min(workers, max(1, n // 50_000)))
    chunk = -(-n // workers)
    _PARTS = [df.iloc[i:i + chunk] for i in range(0, n, chunk)]
    _STATE.clear()
    _STATE.update(cats=cats, conts=conts)
    pool = mp.get_context("fork").Pool(workers) if workers > 1 else None
    try:
        t0 = time.perf_counter()
        idx = list(range(len(_PARTS)))
        res = pool.map(_fit_part, idx) if pool else [_fit_part(i) for i in idx]
        vocabs = {}
        for c in cats:
            gb = _cat.mid_level_groupby([r[0][c] for r in res], [c])
            vocabs[c] = _cat.write_uniques(gb, [c])
        means, stds = {}, {}
        if conts:
            stats = finalize_moments(tree_node_moments([r[1] for r in res]))
            means = {c: float(stats["mean"].loc[c]) for c in conts}
            stds = {c: float(stats["std"].loc[c]) for c in conts}
        t1 = time.perf_counter()
    finally:
        if pool:
            pool.close()
            pool.join()
    # transform: a new pool so the fitted state is inherited by fork
    _STATE.update(vocabs=vocabs, means=means, stds=stds)
    pool = mp.get_context("fork").Pool(workers) if workers > 1 else None
    try:
        t2 = time.perf_counter()
        _ = pool.map(_transform_part, idx) if pool else [_transform_part(i) for i in idx]
        t3 = time.perf_counter()
    finally:
        if pool:
            pool.close()
            pool.join()
    return t1 - t0, t3 - t2, vocabs, means, stds
