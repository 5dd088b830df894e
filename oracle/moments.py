"""Oracle for Normalize.fit statistics — TEST INFRASTRUCTURE.
Follows nvtabular/ops/moments.py:64-116 on pandas."""
import numpy as np
import pandas as pd


def chunkwise_moments(df: pd.DataFrame):
    """moments.py:64-77: count (non-null), sum cast to f64, f64 squared sum."""
    vals = {name: pd.DataFrame() for name in ["count", "sum", "squaredsum"]}
    for name in df.columns:
        column = df[name]
        if column.dtype == object:
            column = pd.Series([x for row in column for x in row])
        vals["count"][name] = [column.count()]
        vals["sum"][name] = [np.float64(column.sum())]
        vals["squaredsum"][name] = [column.astype("float64").pow(2).sum()]
    return vals


def tree_node_moments(inputs):
    """moments.py:80-86."""
    out = {}
    for val in ["count", "sum", "squaredsum"]:
        df_list = [x.get(val, None) for x in inputs]
        df_list = [df for df in df_list if df is not None]
        out[val] = pd.concat(df_list, ignore_index=True).sum().to_frame().transpose()
    return out


def finalize_moments(inp, ddof=1):
    """moments.py:89-116."""
    n = inp["count"].iloc[0].astype("float64")
    x = inp["sum"].iloc[0]
    x2 = inp["squaredsum"].iloc[0]
    var = x2 - x**2 / n
    div = (n - ddof).copy()
    div[div < 1] = 1
    var = var / div
    var[(n - ddof) == 0] = np.nan
    out = pd.DataFrame(index=inp["count"].columns)
    out["count"] = n
    out["sum"] = x
    out["sum2"] = x2
    out["mean"] = x / n
    out["var"] = var
    out["std"] = np.sqrt(var)
    return out
