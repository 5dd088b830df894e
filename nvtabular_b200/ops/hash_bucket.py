"""HashBucket (reference nvtabular/ops/hash_bucket.py:32-131): int32(hash(x) % nb)
with the pandas value hash (see include/nvtb200.h nvtb_hash_bucket_apply)."""
from typing import Dict, Union

import numpy as np
import pandas as pd

from .. import engine
from ..column import Column, DeviceFrame
from ..graph import ColumnSelector, Tags
from .base import Operator


def emb_sz_rule(n_cat: int, minimum_size=16, maximum_size=512):
    """reference nvtabular/ops/categorify.py:687-688"""
    return n_cat, min(max(minimum_size, round(1.6 * n_cat**0.56)), maximum_size)


def string_hash_column(col: Column) -> Column:
    """Strings are hashed on the HOST DICTIONARY (U entries, not N rows) with
    pandas' own string hash — exactly what the reference's CPU branch of
    hash_series does — and carried to the rows as an int64 column of hashes."""
    import torch
    h = pd.util.hash_array(np.asarray(col.dictionary, dtype=object)).view(np.int64) if len(col.dictionary) \
        else np.zeros(0, dtype=np.int64)
    table = torch.from_numpy(h.copy()).to(col.data.device)
    idx = col.data.to(torch.int64).clamp_(0, max(len(h) - 1, 0))
    out = Column(table[idx] if len(h) else torch.zeros_like(idx), col.validity, col.offsets)
    out.prehashed = True
    return out


class HashBucket(Operator):
    def __init__(self, num_buckets: Union[int, Dict[str, int]]):
        if not isinstance(num_buckets, (dict, int)):
            raise TypeError(f"`num_buckets` must be dict, iterable, or int, got type {type(num_buckets)}")
        self.num_buckets = num_buckets
        super().__init__()

    def transform(self, col_selector: ColumnSelector, df: DeviceFrame) -> DeviceFrame:
        if isinstance(self.num_buckets, int):
            num_buckets = {name: self.num_buckets for name in col_selector.names}
        else:
            num_buckets = self.num_buckets
        for col, nb in num_buckets.items():
            c = self._get(df, col)
            if c.is_string:
                leaf = string_hash_column(Column(c.data, c.validity, None, c.dictionary))
            else:
                leaf = Column(c.data, c.validity, None, None, None, c.is_bool)
            out = engine.hash_bucket([leaf], nb, 0, np.int32)
            df[col] = Column(out, None, c.offsets)
        return df

    def get_embedding_sizes(self, columns):
        if isinstance(self.num_buckets, int):
            return {col: emb_sz_rule(self.num_buckets) for col in columns}
        return {col: emb_sz_rule(self.num_buckets[col]) for col in columns}

    def _compute_properties(self, col_schema, input_schema):
        cardinality, dimensions = self.get_embedding_sizes([col_schema.name])[col_schema.name]
        props = dict(input_schema[input_schema.column_names[0]].properties) if input_schema.column_names else {}
        props.update({"domain": {"min": 0, "max": cardinality},
                      "embedding_sizes": {"cardinality": cardinality, "dimension": dimensions}})
        return col_schema.with_properties(props)

    @property
    def output_tags(self):
        return [Tags.CATEGORICAL]

    @property
    def output_dtype(self):
        return np.int32
