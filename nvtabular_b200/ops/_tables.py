"""Host frame <-> device key helpers shared by the operators that keep their fitted state in
parquet files (reference layout: unique.<col>.parquet / cat_stats.<name>.parquet,
nvtabular/ops/categorify.py:719-822, 1073-1137): a workflow reloaded from disk rebuilds its
device tables from those files."""
from typing import List, Tuple

import numpy as np
import pandas as pd
import torch

from ..column import DeviceFrame, unpack_validity
from .keyspace import ComboKeySpace, KeySpace, _leaf


def keys_from_frame(df: pd.DataFrame, key_names: List[str]) -> Tuple[object, torch.Tensor, torch.Tensor]:
    """-> (key space, int64 device keys [n], bool device mask [n] of rows whose key is null).
    The key space is rebuilt from the values present in the file (unseen values are
    out-of-vocabulary by construction)."""
    frame = DeviceFrame.from_pandas(df[key_names].reset_index(drop=True))
    cols = [_leaf(frame[n]) for n in key_names]
    n = len(df)
    if len(key_names) > 1:
        space = ComboKeySpace.fit([cols], ncomp=len(key_names))
        key = space.keys_for(cols)
    else:
        space = KeySpace.for_columns(cols)
        key = space.keys_for(cols[0])
    data = key.data.to(torch.int64)
    if key.validity is not None:
        isnull = ~unpack_validity(key.validity, n)
    else:
        isnull = torch.zeros(n, dtype=torch.bool, device=data.device)
    return space, data, isnull


def key_columns(space, key_names: List[str], keys: np.ndarray, with_null_row: bool = False) -> dict:
    """decoded key columns of a table ({name: Series}); `with_null_row` appends the null group.
    The dtype of the key survives the null: integers become pandas nullable integers (parquet
    then holds an int column with a null, as cuDF writes it — not float64), floats get NaN,
    strings None."""
    comps = space.decode(keys) if isinstance(space, ComboKeySpace) else [space.decode(keys)]
    out = {}
    for name, v in zip(key_names, comps):
        v = np.asarray(v)
        if v.dtype == object:
            isn = np.array([x is None for x in v], dtype=bool)
            vals = [x for x in v if x is not None]
            if vals and all(isinstance(x, (int, np.integer)) for x in vals):       # ints with nulls (combo decode)
                arr = pd.array([pd.NA if m else int(x) for x, m in zip(v, isn)], dtype="Int64")
                ser = pd.Series(arr)
            else:
                ser = pd.Series(v, dtype=object)
        else:
            ser = pd.Series(v)
        if with_null_row:
            if ser.dtype == object:
                ser = pd.concat([ser, pd.Series([None], dtype=object)], ignore_index=True)
            elif pd.api.types.is_integer_dtype(ser.dtype):
                name_dt = ser.dtype.name if pd.api.types.is_extension_array_dtype(ser.dtype) else \
                    ser.dtype.name.replace("int", "Int").replace("uInt", "UInt")
                ser = pd.concat([ser.astype(name_dt), pd.Series(pd.array([pd.NA], dtype=name_dt))], ignore_index=True)
            else:
                ser = pd.concat([ser, pd.Series([np.nan], dtype=ser.dtype)], ignore_index=True)
        out[name] = ser
    return out
