"""FillMissing (reference nvtabular/ops/fill.py:23-80).

The transform is DEFERRED: FillMissing marks each column with the fill value
(`Column.fill`) and the next kernel that reads the column applies it while it
streams the data (K1 moments / K2 normalize — SURVEY.md §2b "fused").  Any
other consumer (or the end of the workflow) materialises it with the
standalone fill kernel (nvtb_fill_apply)."""
from .. import engine
from ..column import Column, DeviceFrame
from ..graph import ColumnSelector
from .base import Operator


def materialize(col: Column) -> Column:
    """Apply a deferred fill now (one pass of the fill kernel)."""
    if col.fill is None:
        return col
    leaf = Column(col.data, col.validity, None, None, None, col.is_bool)
    out, _ = engine.fill_apply([leaf], [float(col.fill)])
    res = out[0]
    res.offsets = col.offsets
    return res


def materialize_many(cols):
    """Materialise several deferred fills in ONE launch."""
    idx = [i for i, c in enumerate(cols) if c.fill is not None]
    if not idx:
        return list(cols)
    same_len = len({cols[i].data.numel() for i in idx}) == 1
    out = list(cols)
    if same_len:
        leaves = [Column(cols[i].data, cols[i].validity, None, None, None, cols[i].is_bool) for i in idx]
        res, _ = engine.fill_apply(leaves, [float(cols[i].fill) for i in idx])
        for i, r in zip(idx, res):
            r.offsets = cols[i].offsets
            out[i] = r
    else:
        for i in idx:
            out[i] = materialize(cols[i])
    return out


class FillMissing(Operator):
    """Replace nulls with a constant (default 0); `add_binary_cols` adds a
    boolean `<col>_filled` column per input (fill.py:49-57, 67-78)."""

    def __init__(self, fill_val=0, add_binary_cols=False):
        super().__init__()
        self.fill_val = fill_val
        self.add_binary_cols = add_binary_cols

    def transform(self, col_selector: ColumnSelector, df: DeviceFrame) -> DeviceFrame:
        names = col_selector.names
        if self.add_binary_cols:
            cols = [self._get(df, n) for n in names]
            filled, flags = engine.fill_apply(cols, [float(self.fill_val)] * len(cols), add_binary_cols=True)
            for n, f, fl, c in zip(names, filled, flags, cols):
                f.offsets = c.offsets
                df[n] = f
                df[f"{n}_filled"] = fl
            return df
        for n in names:
            c = df[n]
            if c.fill is not None:     # fill of an already-filled column is a no-op
                continue
            df[n] = Column(c.data, c.validity, c.offsets, c.dictionary, self.fill_val, c.is_bool)
        return df

    def inference_initialize(self, col_selector, inference_config):
        """fill.py:59-65: the host dict-of-arrays transform for serving"""
        if self.add_binary_cols:
            return None
        from ..inference import FillTransform
        return FillTransform(self)

    def column_mapping(self, col_selector):
        mapping = super().column_mapping(col_selector)
        if self.add_binary_cols:
            for n in col_selector.names:
                mapping[f"{n}_filled"] = [n]
        return mapping

    def _compute_dtype(self, col_schema, input_schema):
        col_schema = super()._compute_dtype(col_schema, input_schema)
        if col_schema.name.endswith("_filled"):
            col_schema = col_schema.with_dtype(bool)
        return col_schema
