"""Oracle for JoinGroupby / TargetEncoding — TEST INFRASTRUCTURE.

Follows the reference on pandas:
  _top_level_groupby / _bottom_level_groupby  nvtabular/ops/categorify.py:955-1137
  JoinGroupby.fit/transform                   nvtabular/ops/join_groupby.py:140-217
  TargetEncoding.fit/_op_group_logic/_add_fold nvtabular/ops/target_encoding.py:171-439
"""
from typing import Dict, List

import numpy as np
import pandas as pd

from .categorify import _make_name
from .moments import chunkwise_moments, finalize_moments, tree_node_moments

AGG_DTYPES = {"count": np.int32, "std": np.float32, "var": np.float32, "mean": np.float32}  # join_groupby.py:29-34


def _top_level(df, group: List[str], agg_cols: List[str], agg_list: List[str], sep="_"):
    """categorify.py:955-1051 with continuous aggregations."""
    sum_sq = "std" in agg_list or "var" in agg_list
    df_gb = df[group + [c for c in agg_cols if c not in group]].copy(deep=False)
    agg_dict = {}
    base_aggs = []
    if "size" in agg_list:
        base_aggs.append("size")
    if set(agg_list).difference({"size", "min", "max"}):
        base_aggs.append("count")                                  # :995-998
    agg_dict[group[0]] = base_aggs                                  # :999 (first key column!)
    for col in agg_cols:
        agg_dict[col] = ["sum"]
        if sum_sq:
            name = _make_name(col, "pow2", sep=sep)
            df_gb[name] = df_gb[col].pow(2)
            agg_dict[name] = ["sum"]
        if "min" in agg_list:
            agg_dict[col].append("min")
        if "max" in agg_list:
            agg_dict[col].append("max")
    gb = df_gb.groupby(group, dropna=False).agg(agg_dict)
    gb.columns = [
        _make_name(*(tuple(group) + name[1:]), sep=sep) if name[0] == group[0]
        else _make_name(*(tuple(group) + name), sep=sep)
        for name in gb.columns.to_flat_index()
    ]                                                               # :1019-1024
    gb.reset_index(inplace=True, drop=False)
    return gb


def _agg_type(col):                                                 # :1140-1146
    if col.endswith("_min"):
        return "min"
    if col.endswith("_max"):
        return "max"
    return "sum"


def _bottom_level(dfs, group, agg_cols, agg_list, sep="_"):
    """categorify.py:1054-1137."""
    df = pd.concat(dfs, ignore_index=True)
    gb = df.groupby(group, dropna=False).agg(
        {c: _agg_type(c) for c in df.columns if c not in group})
    gb.reset_index(drop=False, inplace=True)
    name_count = _make_name(*(group + ["count"]), sep=sep)
    name_size = _make_name(*(group + ["size"]), sep=sep)
    required = list(group)
    if "count" in agg_list:
        required.append(name_count)
    if "size" in agg_list:
        required.append(name_size)
    for cont in agg_cols:
        name_sum = _make_name(*(group + [cont, "sum"]), sep=sep)
        if "sum" in agg_list:
            required.append(name_sum)
        if "mean" in agg_list:
            nm = _make_name(*(group + [cont, "mean"]), sep=sep)
            required.append(nm)
            gb[nm] = gb[name_sum] / gb[name_count]
        if "min" in agg_list:
            required.append(_make_name(*(group + [cont, "min"]), sep=sep))
        if "max" in agg_list:
            required.append(_make_name(*(group + [cont, "max"]), sep=sep))
        if "var" in agg_list or "std" in agg_list:
            n = gb[name_count]
            x = gb[name_sum]
            x2 = gb[_make_name(*(group + [cont, "pow2", "sum"]), sep=sep)]
            result = x2 - x**2 / n
            div = (n - 1).copy()
            div[div < 1] = 1
            result = result / div
            result[(n - 1) == 0] = np.nan
            if "var" in agg_list:
                nm = _make_name(*(group + [cont, "var"]), sep=sep)
                required.append(nm)
                gb[nm] = result
            if "std" in agg_list:
                nm = _make_name(*(group + [cont, "std"]), sep=sep)
                required.append(nm)
                gb[nm] = np.sqrt(result)
    return gb[required]


def groupby_stats(partitions, group: List[str], agg_cols: List[str], agg_list: List[str], sep="_"):
    """_category_stats with agg cols (categorify.py:1543-1555): cat_stats.<name>.parquet."""
    if isinstance(partitions, pd.DataFrame):
        partitions = [partitions]
    if not agg_list:
        agg_list = ["count"]
    return _bottom_level([_top_level(p, group, agg_cols, agg_list, sep) for p in partitions],
                         group, agg_cols, agg_list, sep)


def join_groupby_transform(df: pd.DataFrame, groups, stat_tables: Dict[str, pd.DataFrame], sep="_"):
    """JoinGroupby.transform (join_groupby.py:175-217)."""
    new_df = pd.DataFrame()
    tmp = "__tmp__"
    df = df.copy(deep=False)
    df[tmp] = np.arange(len(df), dtype="int32")
    for g in groups:
        g = [g] if isinstance(g, str) else list(g)
        name = _make_name(*g, sep=sep)
        stat_df = stat_tables[name]
        tran = df[g + [tmp]].merge(stat_df, left_on=g, right_on=g, how="left").sort_values(tmp)
        tran = tran.drop(columns=g + [tmp])
        new_cols = [c for c in tran.columns if c not in new_df.columns]
        part = tran[new_cols].reset_index(drop=True)
        for col in part.columns:
            for agg, dt in AGG_DTYPES.items():
                if col.endswith(f"{sep}{agg}"):
                    part[col] = part[col].astype(dt)
        new_df = pd.concat([new_df, part], axis=1)
    return new_df


def add_fold(n, kfold, fold_seed=None):
    """_add_fold (target_encoding.py:427-439)."""
    typ = np.min_scalar_type(kfold * 2)
    if fold_seed is None:
        fold = np.arange(n, dtype=typ)
        np.mod(fold, kfold, out=fold)
        return fold
    state = np.random.RandomState(fold_seed)
    return state.choice(np.arange(kfold, dtype=typ), n)


def te_transform_part(p, groups, tables, y_mean, target_cols, kfold, p_smooth, out_dtype=None, sep="_",
                      fold_name="__fold__"):
    """TargetEncoding.transform of ONE partition that already carries its fold column
    (target_encoding.py:301-424)."""
    tmp = "__tmp__"
    p = p.copy(deep=False)
    p[tmp] = np.arange(len(p), dtype="int32")
    new_df = None
    for g in groups:
        out_col = [f"TE_{_make_name(*g, sep=sep)}_{x}" for x in target_cols]
        agg_all = tables[_make_name(*g, sep=sep)].copy()
        agg_all.columns = g + ["count_y_all"] + [x + "_sum_y_all" for x in target_cols]
        if kfold > 1:
            cols = [fold_name] + g
            agg_f = tables[_make_name(*cols, sep=sep)].copy()
            agg_f.columns = cols + ["count_y"] + [x + "_sum_y" for x in target_cols]
            agg_f = agg_f.merge(agg_all, on=g, how="left")
            agg_f["count_y_all"] = agg_f["count_y_all"] - agg_f["count_y"]
            for i, x in enumerate(target_cols):
                agg_f[x + "_sum_y_all"] = agg_f[x + "_sum_y_all"] - agg_f[x + "_sum_y"]
                agg_f[out_col[i]] = (agg_f[x + "_sum_y_all"] + p_smooth * y_mean[x]) / (
                    agg_f["count_y_all"] + p_smooth)
            agg_f = agg_f.drop(["count_y_all", "count_y"] + [x + "_sum_y" for x in target_cols]
                               + [x + "_sum_y_all" for x in target_cols], axis=1)
            tran = p[cols + [tmp]].merge(agg_f, on=cols, how="left")
        else:
            cols = g
            for i, x in enumerate(target_cols):
                agg_all[out_col[i]] = (agg_all[x + "_sum_y_all"] + p_smooth * y_mean[x]) / (
                    agg_all["count_y_all"] + p_smooth)
            agg_all = agg_all.drop(["count_y_all"] + [x + "_sum_y_all" for x in target_cols], axis=1)
            tran = p[cols + [tmp]].merge(agg_all, on=cols, how="left")
        for i, x in enumerate(target_cols):
            tran[out_col[i]] = tran[out_col[i]].fillna(y_mean[x])
        if out_dtype is not None:
            tran[out_col] = tran[out_col].astype(out_dtype)
        tran = tran.sort_values(tmp, ignore_index=True).drop(columns=cols + [tmp])
        tran.index = p.index
        tran = tran.astype(out_dtype or np.float32)
        new_df = tran if new_df is None else pd.concat([new_df, tran], axis=1)
    return new_df


def target_encoding(partitions, cat_groups, target_cols: List[str], kfold=3, fold_seed=42,
                    p_smooth=20, out_dtype=None, target_mean=None, sep="_"):
    """TargetEncoding fit + transform on a list of pandas partitions
    (target_encoding.py:171-214, 301-424).  Returns the list of transformed
    partitions and the fitted pieces."""
    if isinstance(partitions, pd.DataFrame):
        partitions = [partitions]
    fold_name = "__fold__"
    parts = []
    for p in partitions:
        p = p.copy(deep=False)
        if kfold > 1:
            p[fold_name] = add_fold(len(p), kfold, fold_seed)     # same seed per partition (:182-188)
        parts.append(p)
    if target_mean is None:
        level = [chunkwise_moments(p[target_cols]) for p in parts]
        stats = finalize_moments(tree_node_moments(level))
        y_mean = {c: float(stats["mean"].loc[c]) for c in target_cols}
    else:
        y_mean = target_mean
    groups = [[g] if isinstance(g, str) else list(g) for g in cat_groups]
    tables = {}
    for g in groups:
        tables[_make_name(*g, sep=sep)] = groupby_stats(parts, g, target_cols, ["count", "sum"], sep)
        if kfold > 1:
            fg = [fold_name] + g
            tables[_make_name(*fg, sep=sep)] = groupby_stats(parts, fg, target_cols, ["count", "sum"], sep)
    outs = [te_transform_part(p, groups, tables, y_mean, target_cols, kfold, p_smooth, out_dtype, sep, fold_name)
            for p in parts]
    return outs, tables, y_mean
