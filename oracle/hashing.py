"""Oracle for dispatch.hash_series on the CPU branch — TEST INFRASTRUCTURE.

The reference calls merlin.core.dispatch.hash_series (un-vendored,
merlin-core>=23.4.0, requirements/base.txt:1) at nvtabular/ops/categorify.py:
1840,1849 and nvtabular/ops/hash_bucket.py:95,98.  Its CPU branch resolves to
pandas.util.hash_array / hash_pandas_object(index=False); for numeric data that
is pandas/core/util/hashing.py::_hash_ndarray: the value's bits viewed as
u{itemsize}, zero-extended to u64, then

    v ^= v >> 30; v *= 0xBF58476D1CE4E5B9; v ^= v >> 27;
    v *= 0x94D049BB133111EB; v ^= v >> 31

This file restates that arithmetic in numpy (it does NOT call pandas' hasher;
tests/golden/hash_golden.json, generated from pandas itself, pins it).
Nulls: a null is what the pandas path sees, the float64 NaN bit pattern.
"""
import numpy as np
import pandas as pd

NAN_BITS = np.uint64(0x7FF8000000000000)


def _mix(v: np.ndarray) -> np.ndarray:
    v = v.astype(np.uint64, copy=True)
    with np.errstate(over="ignore"):
        v ^= v >> np.uint64(30)
        v *= np.uint64(0xBF58476D1CE4E5B9)
        v ^= v >> np.uint64(27)
        v *= np.uint64(0x94D049BB133111EB)
        v ^= v >> np.uint64(31)
    return v


def hash_values(values, mask=None) -> np.ndarray:
    """uint64 hash per element.  `values`: numeric ndarray / Series; `mask`:
    True = null.  A float Series' NaNs are nulls (pandas convention)."""
    if isinstance(values, pd.Series):
        if pd.api.types.is_extension_array_dtype(values.dtype) and \
                pd.api.types.is_numeric_dtype(values.dtype):
            m = values.isna().to_numpy()
            base = np.dtype(str(values.dtype).lower())
            arr = values.fillna(0).to_numpy(dtype=base)
            return hash_values(arr, m if mask is None else (mask | m))
        values = values.to_numpy()
    arr = np.ascontiguousarray(values)
    if arr.dtype == bool:
        bits = arr.astype(np.uint64)
    elif np.issubdtype(arr.dtype, np.number) and arr.dtype.itemsize <= 8:
        bits = arr.view(f"u{arr.dtype.itemsize}").astype(np.uint64)
    else:
        raise TypeError(f"oracle hash only covers numeric dtypes, got {arr.dtype}")
    if np.issubdtype(arr.dtype, np.floating):
        nan = np.isnan(arr)
        mask = nan if mask is None else (np.asarray(mask) | nan)
    if mask is not None:
        bits = bits.copy()
        bits[np.asarray(mask, dtype=bool)] = NAN_BITS
    return _mix(bits)


def hash_bucket(values, num_buckets: int, mask=None) -> np.ndarray:
    """HashBucket.transform (nvtabular/ops/hash_bucket.py:86-100): int32(hash % nb)."""
    return (hash_values(values, mask) % np.uint64(num_buckets)).astype(np.int32)
