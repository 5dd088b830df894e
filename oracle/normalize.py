"""Oracle for Normalize / NormalizeMinMax / FillMissing — TEST INFRASTRUCTURE.
Follows nvtabular/ops/normalize.py:61-90,150-178 and nvtabular/ops/fill.py:49-57."""
import numpy as np
import pandas as pd

from .moments import chunkwise_moments, finalize_moments, tree_node_moments


def fill_missing(df: pd.DataFrame, cols, fill_val=0, add_binary_cols=False) -> pd.DataFrame:
    """FillMissing.transform (fill.py:49-57)."""
    df = df.copy(deep=False)
    if add_binary_cols:
        for col in cols:
            df[f"{col}_filled"] = df[col].isna()
            df[col] = df[col].fillna(fill_val)
    else:
        df[cols] = df[cols].fillna(fill_val)
    return df


def normalize_fit(partitions, cols, split_every=32):
    """Normalize.fit/_custom_moments/fit_finalize (normalize.py:61-68; moments.py:28-61)."""
    if isinstance(partitions, pd.DataFrame):
        partitions = [partitions]
    level = [chunkwise_moments(p[cols]) for p in partitions]
    while len(level) > 1:
        level = [tree_node_moments(level[i:i + split_every]) for i in range(0, len(level), split_every)]
    stats = finalize_moments(tree_node_moments(level))
    means = {c: float(stats["mean"].loc[c]) for c in stats.index}
    stds = {c: float(stats["std"].loc[c]) for c in stats.index}
    return means, stds


def normalize_transform(df: pd.DataFrame, cols, means, stds, out_dtype=None) -> pd.DataFrame:
    """Normalize.transform (normalize.py:71-90)."""
    new_df = pd.DataFrame()
    for name in cols:
        values = df[name]
        if stds[name] > 0:
            values = (values - means[name]) / (stds[name])
        else:
            values = values - means[name]
        new_df[name] = values.astype(out_dtype or np.float64)
    return new_df


def minmax_fit(partitions, cols):
    """NormalizeMinMax.fit (normalize.py:164-178)."""
    if isinstance(partitions, pd.DataFrame):
        partitions = [partitions]
    mins = pd.concat([p[cols].min() for p in partitions], axis=1).min(axis=1)
    maxs = pd.concat([p[cols].max() for p in partitions], axis=1).max(axis=1)
    return {c: mins[c] for c in cols}, {c: maxs[c] for c in cols}


def minmax_transform(df, cols, mins, maxs, out_dtype=None):
    """NormalizeMinMax.transform (normalize.py:150-161)."""
    new_df = pd.DataFrame()
    for name in cols:
        dif = maxs[name] - mins[name]
        if dif > 0:
            new_df[name] = (df[name] - mins[name]) / dif
        elif dif == 0:
            new_df[name] = df[name] / (2 * df[name])
        new_df[name] = new_df[name].astype(out_dtype or np.float64)
    return new_df
