"""Inference-time transforms over dict-of-arrays requests — the twin of the reference's
`nvtabular_cpp.inference` pybind11 module (cpp/nvtabular/inference/categorify.cc:288-347,
fill.cc:108-124), returned by `op.inference_initialize(...)` (nvtabular/ops/categorify.py:602-609,
ops/fill.py:59-65).

A request is {column: numpy array | (values, offsets) | device tensor}.  Host arrays are encoded by
the native host table (csrc/infer.cu: built once from the device vocabulary, probed by host
threads — a serving batch is too small to pay for a PCIe round trip); device tensors take the
device encode kernel.  Both give the labels `Workflow.transform` gives."""
import ctypes
from ctypes import byref, c_void_p

import numpy as np
import torch

from . import _lib
from .column import Column


class DataFormats:
    """bit flags of reference merlin.dag DataFormats (names only; nvtabular/ops/normalize.py:100-108)"""
    PANDAS_DATAFRAME, CUDF_DATAFRAME, NUMPY_DICT_ARRAY, CUPY_DICT_ARRAY = 1, 2, 4, 8


class Supports:
    CPU_DATAFRAME, GPU_DATAFRAME, CPU_DICT_ARRAY, GPU_DICT_ARRAY = 1, 2, 4, 8


class _HostVocab:
    def __init__(self, fv):
        _lib.require_cuda()
        self.lib = _lib.load()
        self.h = c_void_p()
        _lib.check(self.lib.nvtb_infer_vocab_from_device(byref(self.h), fv.vocab.h, _lib.stream_ptr()))

    def __del__(self):
        try:
            if getattr(self, "h", None) is not None and self.h.value:
                self.lib.nvtb_infer_vocab_destroy(self.h)
                self.h = c_void_p()
        except Exception:
            pass

    def encode(self, keys: np.ndarray, valid, null_label, oov_label, first_label, num_buckets, out_dtype, threads=0):
        keys = np.ascontiguousarray(keys)
        code = _lib.I32 if keys.dtype == np.int32 else _lib.I64
        if keys.dtype not in (np.dtype("int32"), np.dtype("int64")):
            keys = keys.astype(np.int64)
            code = _lib.I64
        out = np.empty(len(keys), dtype=out_dtype)
        mask = None
        if valid is not None:
            mask = np.packbits(np.asarray(valid, dtype=bool), bitorder="little")
        _lib.check(self.lib.nvtb_infer_categorify_host(
            self.h, keys.ctypes.data_as(c_void_p), code,
            mask.ctypes.data_as(c_void_p) if mask is not None else None, len(keys), int(null_label), int(oov_label),
            int(first_label), int(num_buckets or 0), out.ctypes.data_as(c_void_p),
            _lib.I32 if np.dtype(out_dtype) == np.int32 else _lib.I64, int(threads)))
        return out


class CategorifyTransform:
    """`nvtabular_cpp.inference.CategorifyTransform(op)` (categorify.cc:288-329)."""

    def __init__(self, op):
        self.op = op
        self._host = {}

    @property
    def supports(self):
        return Supports.CPU_DICT_ARRAY | Supports.GPU_DICT_ARRAY

    @property
    def supported_formats(self):
        return DataFormats.NUMPY_DICT_ARRAY | DataFormats.CUPY_DICT_ARRAY

    def _labels(self, name):
        op = self.op
        storage = op.storage_name.get(name, name)
        fv = op._fitted(storage)
        buckets = op.num_buckets
        nb = (buckets.get(storage, 0) if isinstance(buckets, dict) else buckets) or 0
        null_label = fv.index_start if op.single_table else 1
        oov_label = null_label + 1
        first_label = fv.index_start if op.single_table else oov_label + (nb or 1)
        return storage, fv, nb, null_label, oov_label, first_label

    def _encode(self, name, values):
        storage, fv, nb, null_label, oov_label, first_label = self._labels(name)
        out_dtype = np.dtype(self.op.output_dtype)
        if isinstance(values, torch.Tensor) and values.is_cuda:           # device request: the encode kernel
            key = fv.space.keys_for(Column(values))
            return fv.vocab.encode(key, null_label, oov_label, first_label, nb, [Column(values)] if nb else [], out_dtype)
        arr = np.asarray(values)
        space = fv.space
        valid = None
        if getattr(space, "kind", "int") == "int" and arr.dtype.kind in "iu":
            keys = arr
        else:                                    # strings / floats: host key space (O(batch)), nulls = None / NaN
            import pandas as pd
            ser = pd.Series(arr)
            valid = ~ser.isna().to_numpy()
            keys = space.encode_values(ser.where(valid, ser[valid].iloc[0] if valid.any() else 0)).astype(np.int64)
            if getattr(space, "kind", "") == "str":
                keys = np.where(keys < 0, np.iinfo(np.int64).max, keys)      # unseen strings: never a key
            if nb:
                raise NotImplementedError("hash buckets for string / float keys take the device path")
        hv = self._host.get(storage)
        if hv is None:
            hv = self._host[storage] = _HostVocab(fv)
        return hv.encode(keys, valid, null_label, oov_label, first_label, nb, out_dtype)

    def transform(self, col_selector, tensors: dict):
        for name in list(tensors):
            if name not in self.op.storage_name and dict.get(self.op.categories, name) is None \
                    and name not in self.op.categories.fitted:
                raise ValueError(f"Unknown column for CategorifyTransform {name}")     # categorify.cc:303-308
            v = tensors[name]
            if isinstance(v, tuple):                                      # (values, offsets): list column
                tensors[name] = (self._encode(name, v[0]), v[1])
            else:
                tensors[name] = self._encode(name, v)
        return tensors


class FillTransform:
    """`nvtabular_cpp.inference.FillTransform(op)` (fill.cc:91-124): NaN -> fill_val, in place."""

    def __init__(self, op):
        self.fill_val = float(op.fill_val)
        self.lib = _lib.load()

    @property
    def supports(self):
        return Supports.CPU_DICT_ARRAY

    @property
    def supported_formats(self):
        return DataFormats.NUMPY_DICT_ARRAY

    def _fill(self, arr):
        if isinstance(arr, torch.Tensor):
            return torch.nan_to_num(arr, nan=self.fill_val) if arr.is_floating_point() else arr
        arr = np.ascontiguousarray(arr)
        if arr.dtype.kind != "f":
            return arr
        if not arr.flags.writeable:
            arr = arr.copy()
        code = _lib.F32 if arr.dtype == np.float32 else _lib.F64
        if arr.dtype not in (np.dtype("float32"), np.dtype("float64")):
            arr = arr.astype(np.float64)
            code = _lib.F64
        _lib.check(self.lib.nvtb_infer_fill_host(arr.ctypes.data_as(c_void_p), code, arr.size, self.fill_val))
        return arr

    def transform(self, col_selector, tensors: dict):
        for name in list(tensors):
            v = tensors[name]
            tensors[name] = (self._fill(v[0]), v[1]) if isinstance(v, tuple) else self._fill(v)
        return tensors
