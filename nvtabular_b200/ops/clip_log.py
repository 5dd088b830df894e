"""Clip and LogOp (reference nvtabular/ops/clip.py:23-60, nvtabular/ops/logop.py:31-66): the
continuous-column steps of the published Criteo workflow
(bench/examples/dask-nvtabular-criteo-benchmark.py:201-204: FillMissing >> Clip(min_value=0)
>> LogOp).  One kernel pass each (csrc/scan_kernels.cu transform_column<OP_CLIP | OP_CLIPLOG>);
an upstream FillMissing is consumed inside the pass instead of being materialised first."""
import numpy as np

from .. import engine
from ..column import Column, DeviceFrame
from ..graph import ColumnSelector, Tags
from .base import Operator


class Clip(Operator):
    fuses_fill = True

    def __init__(self, min_value=None, max_value=None):
        if min_value is None and max_value is None:          # clip.py:43-44
            raise ValueError("Must specify a min or max value to clip to")
        super().__init__()
        self.min_value = min_value
        self.max_value = max_value

    def transform(self, col_selector: ColumnSelector, df: DeviceFrame) -> DeviceFrame:
        names = col_selector.names
        cols = [df[n] for n in names]
        outs = engine.cliplog_apply([_leaf(c) for c in cols], self.min_value, self.max_value, False)
        new_df = DeviceFrame()
        for n, o, c in zip(names, outs, cols):
            o.offsets = c.offsets
            new_df[n] = o
        return new_df


class LogOp(Operator):
    """log(1 + x) of continuous columns, float32 out (logop.py:47-56)."""
    fuses_fill = True

    def transform(self, col_selector: ColumnSelector, df: DeviceFrame) -> DeviceFrame:
        names = col_selector.names
        cols = [df[n] for n in names]
        outs = engine.cliplog_apply([_leaf(c) for c in cols], None, None, True, self.output_dtype)
        for n, o, c in zip(names, outs, cols):
            o.offsets = c.offsets
            df[n] = o
        return df

    @property
    def output_tags(self):
        return [Tags.CONTINUOUS]

    @property
    def output_dtype(self):
        return np.float32


def _leaf(col: Column) -> Column:
    """the leaves of a (list) column with the deferred fill still attached: the kernel applies it"""
    return Column(col.data, col.validity, None, None, col.fill, col.is_bool)
