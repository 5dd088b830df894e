"""Normalize / NormalizeMinMax (reference nvtabular/ops/normalize.py:33-212,
statistics from nvtabular/ops/moments.py:28-116).

fit  = one fused scan per partition (K1: count, sum, sumsq, min, max of every
       column in a single launch, an upstream FillMissing folded in) + one NCCL
       all-reduce across GPUs; the dask tree of moments.py:45-55 disappears.
transform = one fused pass (K2), again with FillMissing folded in."""
import numpy as np

from .. import engine
from ..column import Column, DeviceFrame
from ..graph import ColumnSelector, Tags
from .base import StatOperator


def _leaf(col: Column) -> Column:
    return Column(col.data, col.validity, None, None, col.fill, col.is_bool)


class _MomentsOp(StatOperator):
    fit_blocks_host = True      # fit ends with the moments read-back (Workflow.fit orders on this)
    fuses_fill = True

    def __init__(self, out_dtype=None):
        super().__init__()
        self.out_dtype = out_dtype

    def _fit_moments(self, col_selector: ColumnSelector, ddf):
        names = col_selector.names
        m = None
        for df in ddf:
            cols = [_leaf(df[n]) for n in names]
            if m is None:
                m = engine.Moments(len(names), device=cols[0].data.device)
            if len({c.data.numel() for c in cols}) == 1:
                m.accumulate(cols)
            else:   # ragged list columns: one launch per column length
                raise NotImplementedError("columns of one Normalize must have equal leaf counts")
        if m is None:
            return names, None
        m.allreduce()
        return names, m.result()

    @property
    def output_tags(self):
        return [Tags.CONTINUOUS]

    @property
    def output_dtype(self):
        return self.out_dtype or np.float64

    @property
    def supports(self):
        from ..inference import Supports
        return Supports.CPU_DICT_ARRAY | Supports.GPU_DICT_ARRAY | Supports.CPU_DATAFRAME | Supports.GPU_DATAFRAME

    @property
    def supported_formats(self):          # normalize.py:100-108
        from ..inference import DataFormats
        return (DataFormats.PANDAS_DATAFRAME | DataFormats.CUDF_DATAFRAME | DataFormats.NUMPY_DICT_ARRAY
                | DataFormats.CUPY_DICT_ARRAY)


class Normalize(_MomentsOp):
    """(x - mean) / std with ddof=1 statistics (normalize.py:61-90)."""

    def __init__(self, out_dtype=None):
        super().__init__(out_dtype)
        self.means = {}
        self.stds = {}

    def fit(self, col_selector: ColumnSelector, ddf):
        return self._fit_moments(col_selector, ddf)

    def fit_finalize(self, stats):
        names, r = stats
        if r is None:
            return
        for i, col in enumerate(names):
            self.means[col] = float(r["mean"][i])
            self.stds[col] = float(r["std"][i])

    def transform(self, col_selector: ColumnSelector, df) -> DeviceFrame:
        names = col_selector.names
        if isinstance(df, dict):     # dict-of-arrays path (normalize.py:101-108)
            frame = DeviceFrame.from_dict(df)
            out = self.transform(col_selector, frame)
            return {n: out[n].data for n in names}
        new_df = DeviceFrame()
        cols = [df[n] for n in names]
        outs = engine.normalize_apply([_leaf(c) for c in cols], [self.means[n] for n in names],
                                      [self.stds[n] for n in names], self.output_dtype)
        for n, o, c in zip(names, outs, cols):
            o.offsets = c.offsets
            new_df[n] = o
        return new_df

    def clear(self):
        self.means = {}
        self.stds = {}


class NormalizeMinMax(_MomentsOp):
    """(x - min) / (max - min) (normalize.py:150-178)."""

    def __init__(self, out_dtype=None):
        super().__init__(out_dtype)
        self.mins = {}
        self.maxs = {}

    def fit(self, col_selector: ColumnSelector, ddf):
        return self._fit_moments(col_selector, ddf)

    def fit_finalize(self, stats):
        names, r = stats
        if r is None:
            return
        for i, col in enumerate(names):
            self.mins[col] = float(r["min"][i])
            self.maxs[col] = float(r["max"][i])

    def transform(self, col_selector: ColumnSelector, df) -> DeviceFrame:
        names = col_selector.names
        new_df = DeviceFrame()
        cols = [df[n] for n in names]
        outs = engine.minmax_apply([_leaf(c) for c in cols], [self.mins[n] for n in names],
                                   [self.maxs[n] for n in names], self.output_dtype)
        for n, o, c in zip(names, outs, cols):
            o.offsets = c.offsets
            new_df[n] = o
        return new_df

    def clear(self):
        self.mins = {}
        self.maxs = {}
