"""Device column model: typed data + Arrow validity bitmask (+ list offsets,
+ string dictionary).  This replaces merlin.core.dispatch's DataFrameType
(reference nvtabular/dispatch.py:21) for the hot path: kernels read these
buffers through the C-ABI (include/nvtb200.h, nvtb_col_t).

Host<->device conversion lives here and is NOT on the measured hot path; it
uses torch only as a memory/stream provider.
"""
from typing import Dict, List, Optional

import numpy as np
import pandas as pd
import torch

from . import _lib

_TORCH2CODE = {
    torch.int32: _lib.I32, torch.int64: _lib.I64, torch.float32: _lib.F32,
    torch.float64: _lib.F64, torch.uint8: _lib.U8, torch.bool: _lib.U8,
}
_NP2TORCH = {
    np.dtype("int32"): torch.int32, np.dtype("int64"): torch.int64,
    np.dtype("float32"): torch.float32, np.dtype("float64"): torch.float64,
    np.dtype("uint8"): torch.uint8, np.dtype("bool"): torch.bool,
}
_BIT_WEIGHTS = (1, 2, 4, 8, 16, 32, 64, 128)


def default_device():
    return torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")


def pack_validity(valid: torch.Tensor) -> torch.Tensor:
    """bool[n] -> uint8[ceil(n/8)] bitmask, LSB-first (Arrow layout), padded to 32 B."""
    n = valid.numel()
    nbytes = (n + 7) // 8
    padded = ((nbytes + 31) // 32) * 32
    bits = torch.zeros(padded * 8, dtype=torch.uint8, device=valid.device)
    bits[:n] = valid.to(torch.uint8)
    w = torch.tensor(_BIT_WEIGHTS, dtype=torch.uint8, device=valid.device)
    return (bits.view(-1, 8) * w).sum(dim=1, dtype=torch.int32).to(torch.uint8)


def unpack_validity(mask: Optional[torch.Tensor], n: int, device=None) -> torch.Tensor:
    """uint8 bitmask -> bool[n]."""
    if mask is None:
        return torch.ones(n, dtype=torch.bool, device=device)
    w = torch.tensor(_BIT_WEIGHTS, dtype=torch.uint8, device=mask.device)
    bits = (mask.view(-1, 1) & w) != 0
    return bits.reshape(-1)[:n]


class Column:
    """One column resident in device memory.

    data      : 1-D tensor (int32/int64/float32/float64/uint8); for a list
                column the flattened leaf values; for a string column the
                order-preserving dictionary codes.
    validity  : uint8 bitmask (bit=1 -> non-null) or None
    offsets   : int64[nrows+1] for list (multi-hot) columns, else None
    dictionary: numpy object array of the sorted distinct strings, else None
    fill      : a deferred FillMissing value (fused into the next kernel that
                consumes the column), else None
    """

    __slots__ = ("data", "validity", "offsets", "dictionary", "fill", "is_bool", "prehashed")

    def __init__(self, data, validity=None, offsets=None, dictionary=None, fill=None, is_bool=False):
        if data.dtype == torch.bool:
            data = data.to(torch.uint8)
            is_bool = True
        self.data = data
        self.validity = validity
        self.offsets = offsets
        self.dictionary = dictionary
        self.fill = fill
        self.is_bool = is_bool
        self.prehashed = False   # int64 data that already IS a value hash (NVTB_H64)

    # ------------------------------------------------------------------ meta
    @property
    def nrows(self) -> int:
        return int(self.offsets.numel() - 1) if self.offsets is not None else int(self.data.numel())

    def __len__(self):
        return self.nrows

    @property
    def is_list(self) -> bool:
        return self.offsets is not None

    @property
    def is_string(self) -> bool:
        return self.dictionary is not None

    @property
    def dtype_code(self) -> int:
        if self.prehashed:
            return _lib.H64
        return _TORCH2CODE[self.data.dtype]

    @property
    def np_dtype(self):
        if self.is_string:
            return np.dtype("object")
        if self.is_bool:
            return np.dtype("bool")
        return np.dtype(str(self.data.dtype).replace("torch.", ""))

    def desc(self):
        """(data_ptr, validity_ptr|None, dtype_code) for _lib.col_array."""
        return (self.data.data_ptr() if self.data.numel() else None,
                self.validity.data_ptr() if self.validity is not None else None,
                self.dtype_code)

    def leaves(self) -> "Column":
        """The flattened values of a list column (itself otherwise)."""
        if not self.is_list:
            return self
        return Column(self.data, self.validity, None, self.dictionary, self.fill, self.is_bool)

    def with_data(self, data, validity=None, keep_list=True) -> "Column":
        return Column(data, validity, self.offsets if keep_list else None, None)

    def to(self, device, non_blocking=True) -> "Column":
        """Copy the buffers to `device` (pinned host <-> HBM copies are asynchronous)."""
        out = Column(self.data.to(device, non_blocking=non_blocking),
                     self.validity.to(device, non_blocking=non_blocking) if self.validity is not None else None,
                     self.offsets.to(device, non_blocking=non_blocking) if self.offsets is not None else None,
                     self.dictionary, self.fill, self.is_bool)
        out.prehashed = self.prehashed
        return out

    def pin(self) -> "Column":
        """A pinned-host copy (the staging format of the end-to-end path)."""
        def _p(t):
            if t is None:
                return None
            buf = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
            buf.copy_(t)
            return buf
        out = Column(_p(self.data), _p(self.validity), _p(self.offsets), self.dictionary, self.fill, self.is_bool)
        out.prehashed = self.prehashed
        return out

    @property
    def device(self):
        return self.data.device

    def nbytes(self) -> int:
        n = self.data.numel() * self.data.element_size()
        if self.validity is not None:
            n += self.validity.numel()
        if self.offsets is not None:
            n += self.offsets.numel() * 8
        return n

    def null_count(self) -> int:
        if self.validity is None:
            return 0
        n = self.data.numel()
        return int(n - unpack_validity(self.validity, n).sum().item())

    # ------------------------------------------------------------- host -> device
    @classmethod
    def from_numpy(cls, arr: np.ndarray, mask: Optional[np.ndarray] = None, device=None) -> "Column":
        """arr: numeric ndarray; mask: bool ndarray, True = NULL (pandas convention)."""
        device = device or default_device()
        arr = np.ascontiguousarray(arr)
        if arr.dtype not in _NP2TORCH:
            if np.issubdtype(arr.dtype, np.integer):
                # small / unsigned ints widen to the next supported signed type
                arr = arr.astype("int32" if arr.dtype.itemsize < 4 else "int64")
            elif np.issubdtype(arr.dtype, np.floating):
                arr = arr.astype("float32" if arr.dtype.itemsize < 4 else "float64")
            else:
                raise TypeError(f"unsupported column dtype {arr.dtype}")
        if not arr.flags.writeable:
            arr = arr.copy()
        t = torch.from_numpy(arr).to(device)
        validity = None
        if mask is not None and mask.any():
            validity = pack_validity(torch.from_numpy(~np.asarray(mask, dtype=bool)).to(device))
        return cls(t, validity)

    @classmethod
    def from_strings(cls, values, device=None) -> "Column":
        """Dictionary-encode strings with ORDER-PRESERVING codes (code order ==
        string order), so the (size desc, key asc) vocabulary rule holds on codes."""
        device = device or default_device()
        ser = pd.Series(values, dtype="object")
        isnull = ser.isna().to_numpy()
        uniq = np.array(sorted(set(ser[~isnull].tolist())), dtype=object)
        lut = {s: i for i, s in enumerate(uniq)}
        codes = np.fromiter((lut[v] if not m else 0 for v, m in zip(ser.tolist(), isnull)),
                            dtype=np.int32, count=len(ser))
        col = cls.from_numpy(codes, isnull, device)
        col.dictionary = uniq
        return col

    @classmethod
    def from_pandas(cls, ser: pd.Series, device=None) -> "Column":
        device = device or default_device()
        dt = ser.dtype
        if isinstance(dt, pd.CategoricalDtype):
            ser = ser.astype(object)
            dt = ser.dtype
        if pd.api.types.is_bool_dtype(dt) and not pd.api.types.is_extension_array_dtype(dt):
            return cls.from_numpy(ser.to_numpy(), None, device)
        if pd.api.types.is_extension_array_dtype(dt) and (
                pd.api.types.is_integer_dtype(dt) or pd.api.types.is_float_dtype(dt)
                or pd.api.types.is_bool_dtype(dt)):
            mask = ser.isna().to_numpy()
            base = np.dtype(str(dt).lower().replace("boolean", "bool"))
            vals = ser.fillna(0).to_numpy(dtype=base)
            return cls.from_numpy(vals, mask, device)
        if pd.api.types.is_numeric_dtype(dt):
            arr = ser.to_numpy()
            mask = np.isnan(arr) if np.issubdtype(arr.dtype, np.floating) else None
            return cls.from_numpy(arr, mask, device)
        # object / string: strings or lists
        vals = ser.tolist()
        first = next((v for v in vals if v is not None and not (isinstance(v, float) and np.isnan(v))), None)
        if isinstance(first, (list, tuple, np.ndarray)):
            return cls.from_lists(vals, device)
        return cls.from_strings(vals, device)

    @classmethod
    def from_arrow(cls, arr, device=None, pin=False) -> "Column":
        """pyarrow Array / ChunkedArray -> Column without going through pandas: the Arrow
        validity bitmap already IS this engine's bitmask (LSB-first, 1 = non-null), and
        nullable integers stay integers (pandas would turn them into float64 + NaN).
        `pin=True` stages the buffers in pinned host memory (the parquet ingest path,
        SURVEY.md 8f-1; call site in the reference: bench/examples/dask-nvtabular-criteo-
        benchmark.py:216 `Dataset(path, engine="parquet", part_size=...)`)."""
        import pyarrow as pa
        if isinstance(arr, pa.ChunkedArray):
            arr = arr.combine_chunks() if arr.num_chunks != 1 else arr.chunk(0)
        t = arr.type
        if pa.types.is_dictionary(t):
            arr = arr.dictionary_decode()
            t = arr.type
        device = device or (torch.device("cpu") if pin else default_device())

        def place(np_arr):
            ten = torch.from_numpy(np_arr)
            if pin and torch.cuda.is_available():
                buf = torch.empty(ten.shape, dtype=ten.dtype, pin_memory=True)
                buf.copy_(ten)
                return buf
            return ten.to(device)

        if pa.types.is_string(t) or pa.types.is_large_string(t):
            return cls.from_strings(arr.to_pylist(), device)
        if pa.types.is_list(t) or pa.types.is_large_list(t):
            flat = arr.flatten()                                  # honours slices
            off = arr.offsets.to_numpy(zero_copy_only=False).astype(np.int64)
            off = off - off[0]
            if arr.null_count:                                    # a null row is an empty row
                lens = np.where(arr.is_valid().to_numpy(zero_copy_only=False), np.diff(off), 0)
                off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
                flat = pa.concat_arrays([x.values for x in arr if x.is_valid]) if len(arr) else flat
            leaf = cls.from_arrow(flat, device, pin)
            leaf.offsets = place(off)
            return leaf
        n = len(arr)
        bufs = arr.buffers()
        if pa.types.is_boolean(t):
            vals = arr.fill_null(False).to_numpy(zero_copy_only=False).astype(np.uint8)
            is_bool = True
        else:
            np_dt = np.dtype(t.to_pandas_dtype())
            if np_dt not in _NP2TORCH:
                if np.issubdtype(np_dt, np.integer):
                    arr = arr.cast(pa.int32() if np_dt.itemsize < 4 else pa.int64())
                elif np.issubdtype(np_dt, np.floating):
                    arr = arr.cast(pa.float32() if np_dt.itemsize < 4 else pa.float64())
                else:
                    raise TypeError(f"unsupported arrow type {t}")
                bufs = arr.buffers()
                np_dt = np.dtype(arr.type.to_pandas_dtype())
            vals = np.frombuffer(bufs[1], dtype=np_dt)[arr.offset: arr.offset + n] if n else np.zeros(0, np_dt)
            vals = np.array(vals)            # own the memory (the table may be dropped)
            is_bool = False
        validity = None
        if arr.null_count:
            nbytes = (n + 7) // 8
            padded = ((nbytes + 31) // 32) * 32
            out = np.zeros(padded, dtype=np.uint8)
            if arr.offset % 8 == 0:
                raw = np.frombuffer(bufs[0], dtype=np.uint8)[arr.offset // 8: arr.offset // 8 + nbytes]
                out[:nbytes] = raw
            else:
                out[:nbytes] = np.packbits(arr.is_valid().to_numpy(zero_copy_only=False), bitorder="little")
            if n % 8:                            # bits past the last row are zero, like pack_validity
                out[nbytes - 1] &= (1 << (n % 8)) - 1
            validity = place(out)
        col = cls(place(vals), validity)
        col.is_bool = is_bool
        return col

    def to_arrow(self):
        """-> pyarrow Array (host).  Numeric columns reuse the data and bitmask buffers."""
        import pyarrow as pa
        if self.dictionary is not None or self.offsets is not None:
            return pa.Array.from_pandas(self.to_pandas())
        vals = self.data.detach().cpu().numpy()
        n = len(vals)
        if self.is_bool:
            mask = None
            if self.validity is not None:
                mask = ~unpack_validity(self.validity, n).cpu().numpy()
            return pa.array(vals.astype(bool), mask=mask)
        vbuf = None
        nulls = 0
        if self.validity is not None:
            bits = self.validity.detach().cpu().numpy()
            nulls = n - int(unpack_validity(self.validity, n).sum().item())
            vbuf = pa.py_buffer(bits[: (n + 7) // 8].tobytes()) if nulls else None
        return pa.Array.from_buffers(pa.from_numpy_dtype(vals.dtype), n, [vbuf, pa.py_buffer(vals)], null_count=nulls)

    @classmethod
    def from_lists(cls, rows, device=None) -> "Column":
        device = device or default_device()
        lens = np.fromiter((0 if r is None else len(r) for r in rows), dtype=np.int64, count=len(rows))
        offsets = np.zeros(len(rows) + 1, dtype=np.int64)
        np.cumsum(lens, out=offsets[1:])
        flat: List = []
        for r in rows:
            if r is not None:
                flat.extend(list(r))
        leaf = cls.from_pandas(pd.Series(flat, dtype=object if (flat and isinstance(flat[0], str)) else None), device) \
            if flat else cls(torch.zeros(0, dtype=torch.int64, device=device))
        leaf.offsets = torch.from_numpy(offsets).to(device)
        return leaf

    # ------------------------------------------------------------- device -> host
    def materialize_fill(self) -> "Column":
        """Apply a deferred FillMissing on the host side of a conversion."""
        return self

    def to_numpy(self):
        """(values ndarray, null-mask ndarray|None) of the leaf values."""
        vals = self.data.detach().cpu().numpy()
        if self.is_bool:
            vals = vals.astype(bool)
        mask = None
        if self.validity is not None:
            mask = ~unpack_validity(self.validity, self.data.numel()).cpu().numpy()
            if not mask.any():
                mask = None
        return vals, mask

    def to_pandas(self, name=None) -> pd.Series:
        vals, mask = self.to_numpy()
        if self.dictionary is not None:
            out = np.empty(len(vals), dtype=object)
            if len(vals):
                out[:] = self.dictionary[np.clip(vals, 0, max(len(self.dictionary) - 1, 0))] \
                    if len(self.dictionary) else None
            if mask is not None:
                out[mask] = None
            leaf = out
        elif mask is not None:
            # pandas' own convention: numeric column with nulls -> float64 NaN
            leaf = vals.astype("float64") if not np.issubdtype(vals.dtype, np.floating) else vals.copy()
            leaf[mask] = np.nan
        else:
            leaf = vals
        if self.offsets is not None:
            off = self.offsets.cpu().numpy()
            rows = [leaf[off[i]:off[i + 1]] for i in range(len(off) - 1)]
            return pd.Series(rows, name=name, dtype=object)
        return pd.Series(leaf, name=name)


class DeviceFrame:
    """An ordered set of equal-length Columns — the engine's DataFrame."""

    def __init__(self, columns: Optional[Dict[str, Column]] = None):
        self._cols: Dict[str, Column] = dict(columns or {})

    @classmethod
    def from_pandas(cls, df: pd.DataFrame, device=None, columns=None) -> "DeviceFrame":
        names = list(columns) if columns is not None else list(df.columns)
        return cls({n: Column.from_pandas(df[n], device) for n in names})

    @classmethod
    def from_arrow(cls, table, device=None, pin=False, columns=None) -> "DeviceFrame":
        names = list(columns) if columns is not None else list(table.column_names)
        return cls({n: Column.from_arrow(table.column(n), device, pin) for n in names})

    def to_arrow(self):
        import pyarrow as pa
        return pa.table({k: c.to_arrow() for k, c in self._cols.items()})

    @classmethod
    def from_dict(cls, d, device=None) -> "DeviceFrame":
        out = {}
        for k, v in d.items():
            if isinstance(v, Column):
                out[k] = v
            elif isinstance(v, torch.Tensor):
                out[k] = Column(v)
            elif isinstance(v, tuple) and len(v) == 2 and isinstance(v[0], torch.Tensor):
                out[k] = Column(v[0], v[1])
            else:
                out[k] = Column.from_pandas(pd.Series(v), device)
        return cls(out)

    @property
    def columns(self) -> List[str]:
        return list(self._cols)

    def __contains__(self, name):
        return name in self._cols

    def __getitem__(self, name) -> Column:
        if isinstance(name, (list, tuple)):
            return DeviceFrame({n: self._cols[n] for n in name})
        return self._cols[name]

    def __setitem__(self, name, col: Column):
        self._cols[name] = col

    def __len__(self):
        for c in self._cols.values():
            return c.nrows
        return 0

    def copy(self) -> "DeviceFrame":
        return DeviceFrame(dict(self._cols))

    def slice_rows(self, start: int, stop: int) -> "DeviceFrame":
        """Rows [start, stop) as views; `start` must be a multiple of 64 (keeps the 32-byte
        alignment of the data and the byte alignment of the validity bitmask)."""
        assert start % 64 == 0, "partition boundaries must be multiples of 64 rows"
        out = {}
        for k, c in self._cols.items():
            assert c.offsets is None, "list columns cannot be row-sliced"
            v = None
            if c.validity is not None:
                nb = (stop - start + 7) // 8
                v = c.validity[start // 8: start // 8 + ((nb + 31) // 32) * 32]
                if v.numel() < nb:
                    v = c.validity[start // 8: start // 8 + nb]
            col = Column(c.data[start:stop], v, None, c.dictionary, c.fill, c.is_bool)
            col.prehashed = c.prehashed
            out[k] = col
        return DeviceFrame(out)

    def to(self, device, non_blocking=True) -> "DeviceFrame":
        return DeviceFrame({k: c.to(device, non_blocking) for k, c in self._cols.items()})

    def pin(self) -> "DeviceFrame":
        return DeviceFrame({k: c.pin() for k, c in self._cols.items()})

    @property
    def is_host(self) -> bool:
        for c in self._cols.values():
            return c.data.device.type == "cpu"
        return False

    def nbytes(self) -> int:
        return sum(c.nbytes() for c in self._cols.values())

    def drop(self, names) -> "DeviceFrame":
        return DeviceFrame({k: v for k, v in self._cols.items() if k not in set(names)})

    def items(self):
        return self._cols.items()

    def to_pandas(self) -> pd.DataFrame:
        from .ops.fill import materialize  # deferred fills are applied by the fill kernel
        out = {}
        for k, c in self._cols.items():
            out[k] = materialize(c).to_pandas(k)
        return pd.DataFrame(out) if out else pd.DataFrame()
