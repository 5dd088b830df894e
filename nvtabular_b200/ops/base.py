"""Operator / StatOperator contract of the reference (merlin.dag BaseOperator /
StatOperator, aliased at reference nvtabular/ops/operator.py:16-27 and
nvtabular/ops/stat_operator.py:16): transform(col_selector, df),
fit(col_selector, ddf) -> stats, fit_finalize(stats), clear(),
set_storage_path(), column_mapping(), output dtype/tags/properties hooks.
`df` is a DeviceFrame; `ddf` is an iterable of DeviceFrames (partitions)."""
from typing import Dict, List

from ..column import Column, DeviceFrame
from ..graph import ColumnSchema, ColumnSelector, Schema


class _OperatorMeta(type):
    """lets an operator CLASS stand on the right of `>>` (`cols >> ops.Categorify`)."""

    def __rrshift__(cls, other):
        return ColumnSelector(other) >> cls()


class Operator(metaclass=_OperatorMeta):
    #: ops that consume a deferred FillMissing inside their own kernel
    fuses_fill = False

    def __init__(self):
        pass

    # ------------------------------------------------------------------ contract
    def transform(self, col_selector: ColumnSelector, df: DeviceFrame) -> DeviceFrame:
        raise NotImplementedError

    def column_mapping(self, col_selector: ColumnSelector) -> Dict[str, List[str]]:
        return {name: [name] for name in col_selector.names}

    @property
    def dependencies(self):
        return None

    @property
    def output_dtype(self):
        return None

    @property
    def output_tags(self):
        return []

    @property
    def label(self):
        return type(self).__name__

    def __rrshift__(self, other):
        return ColumnSelector(other) >> self

    # ------------------------------------------------ serving hooks (reference operator contract)
    def inference_initialize(self, col_selector, inference_config):
        """an inference-time replacement of this op for dict-of-arrays requests, or None
        (merlin.dag BaseOperator hook; overridden by Categorify / FillMissing)"""
        return None

    def compute_selector(self, input_schema, selector, parents_selector=None, dependencies_selector=None):
        """the columns this op reads: its parents' outputs (reference categorify.py:589-597)"""
        sel = parents_selector if parents_selector is not None else selector
        missing = [n for n in sel.names if input_schema is not None and n not in input_schema]
        if missing:
            raise ValueError(f"Missing columns {missing} found in operator {type(self).__name__} "
                             "during computing input selector.")
        return sel

    @property
    def supports(self):
        from ..inference import Supports
        return Supports.GPU_DATAFRAME | Supports.CPU_DATAFRAME

    @property
    def supported_formats(self):
        """frames in, frames out by default; ops with a dict-of-arrays path add the *_DICT_ARRAY bits
        (reference nvtabular/ops/normalize.py:92-108)"""
        from ..inference import DataFormats
        return DataFormats.PANDAS_DATAFRAME | DataFormats.CUDF_DATAFRAME

    # -------------------------------------------------------------------- schema
    def _compute_dtype(self, col_schema: ColumnSchema, input_schema: Schema) -> ColumnSchema:
        dtype = col_schema.dtype
        is_list, is_ragged = col_schema.is_list, col_schema.is_ragged
        if input_schema.column_names:
            src = input_schema[input_schema.column_names[0]]
            dtype = dtype or src.dtype
            is_list, is_ragged = src.is_list, src.is_ragged
        if self.output_dtype is not None:
            dtype = self.output_dtype
        return col_schema.with_dtype(dtype, is_list, is_ragged)

    def _compute_tags(self, col_schema, input_schema):
        tags = []
        if input_schema.column_names:
            tags = list(input_schema[input_schema.column_names[0]].tags)
        return col_schema.with_tags(tags + list(self.output_tags))

    def _compute_properties(self, col_schema, input_schema):
        props = {}
        if input_schema.column_names:
            props = dict(input_schema[input_schema.column_names[0]].properties)
        return col_schema.with_properties(props)

    def compute_output_schema(self, input_schema: Schema, col_selector: ColumnSelector) -> Schema:
        out = []
        for name, sources in self.column_mapping(col_selector).items():
            src_schema = Schema([input_schema[s] for s in sources if s in input_schema])
            cs = ColumnSchema(name)
            cs = self._compute_dtype(cs, src_schema)
            cs = self._compute_tags(cs, src_schema)
            cs = self._compute_properties(cs, src_schema)
            out.append(cs)
        return Schema(out)

    # --------------------------------------------------------------------- helpers
    @staticmethod
    def _get(df: DeviceFrame, name: str, fuse_fill=False) -> Column:
        col = df[name]
        if col.fill is not None and not fuse_fill:
            from .fill import materialize
            col = materialize(col)
        return col


class StatOperator(Operator):
    def fit(self, col_selector: ColumnSelector, ddf):
        raise NotImplementedError

    def fit_finalize(self, stats):
        raise NotImplementedError

    def clear(self):
        raise NotImplementedError

    def set_storage_path(self, new_path, copy=False):
        """Certain stat operators may need external storage"""
