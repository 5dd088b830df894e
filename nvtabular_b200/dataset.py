"""Dataset: the partitioned, device-resident table a Workflow runs on.

Stands in for merlin.io.Dataset (un-vendored; call sites reference
nvtabular/workflow/workflow.py:195-248, bench/examples/
dask-nvtabular-criteo-benchmark.py:216).  A partition is a DeviceFrame in HBM;
there is no dask graph — `to_ddf().compute()` simply brings a (lazily
transformed) dataset back as a pandas frame, as the reference's tests do.
"""
from typing import Callable, Iterable, List, Optional, Union

import numpy as np
import pandas as pd
import torch

from .column import Column, DeviceFrame
from .graph import ColumnSchema, Schema


def _schema_of(frame: DeviceFrame) -> Schema:
    cols = []
    for name, c in frame.items():
        cols.append(ColumnSchema(name, dtype=c.np_dtype, is_list=c.is_list, is_ragged=c.is_list))
    return Schema(cols)


class _Lazy:
    """What `Dataset.to_ddf()` returns: `.compute()` -> pandas DataFrame."""

    def __init__(self, ds: "Dataset", columns=None):
        self._ds = ds
        self._columns = columns

    def compute(self, scheduler=None, **kwargs) -> pd.DataFrame:
        frames = []
        for part in self._ds.partitions():
            if self._columns is not None:
                part = part[list(self._columns)]
            frames.append(part.to_pandas())
        if not frames:
            return pd.DataFrame()
        return pd.concat(frames, ignore_index=True) if len(frames) > 1 else frames[0]

    @property
    def npartitions(self):
        return self._ds.npartitions

    @property
    def columns(self):
        return self._ds.schema.column_names

    def head(self, n=5):
        return self.compute().head(n)

    def __getitem__(self, cols):
        return _Lazy(self._ds, [cols] if isinstance(cols, str) else list(cols))


class Dataset:
    """`Dataset(df)`, `Dataset([df0, df1])`, `Dataset(DeviceFrame)`, `Dataset(dict of tensors)`,
    `Dataset("file.parquet" | [paths])`.  `npartitions` splits a single host frame by rows
    (what `dd.from_pandas(df, npartitions=k)` does in the reference's tests)."""

    def __init__(self, data, engine=None, npartitions: Optional[int] = None, cpu: bool = False,
                 part_size=None, schema: Optional[Schema] = None, device=None,
                 _transform: Optional[Callable] = None, base_dataset=None, **kwargs):
        self._device = device
        self._transform = _transform
        self.base_dataset = base_dataset or self
        self.cpu = cpu     # accepted for API compatibility; there is no CPU engine
        self._parts: Optional[List[DeviceFrame]] = None
        self._source = data
        self._npartitions = npartitions
        if isinstance(data, Dataset):
            self._source = data._source
            self._parts = data._parts
            self._transform = _transform or data._transform
        self._schema = schema

    # --------------------------------------------------------------- ingestion
    def _ingest(self) -> List[DeviceFrame]:
        src = self._source
        if isinstance(src, (str, bytes)) or (isinstance(src, (list, tuple)) and src
                                             and all(isinstance(s, str) for s in src)):
            paths = [src] if isinstance(src, (str, bytes)) else list(src)
            host = [pd.read_parquet(p) if str(p).endswith(".parquet") or not str(p).endswith(".csv")
                    else pd.read_csv(p) for p in paths]
        elif isinstance(src, pd.DataFrame):
            host = [src]
        elif isinstance(src, DeviceFrame):
            return [src]
        elif isinstance(src, dict):
            return [DeviceFrame.from_dict(src, self._device)]
        elif isinstance(src, (list, tuple)):
            out = []
            for s in src:
                out += Dataset(s, device=self._device)._ingest()
            return out
        elif hasattr(src, "to_pandas"):          # pyarrow.Table and friends
            host = [src.to_pandas()]
        else:
            raise TypeError(f"cannot build a Dataset from {type(src)}")
        if self._npartitions and self._npartitions > 1 and len(host) == 1:
            df = host[0]
            n = len(df)
            k = self._npartitions
            # dask's from_pandas split: chunks of ceil(n / k) rows
            chunk = -(-n // k) if n else 0
            host = [df.iloc[i:i + chunk] for i in range(0, n, chunk)] if chunk else [df]
        return [DeviceFrame.from_pandas(h.reset_index(drop=True), self._device) for h in host]

    def partitions(self) -> Iterable[DeviceFrame]:
        if self._parts is None:
            self._parts = self._ingest()
        for p in self._parts:
            yield self._transform(p) if self._transform is not None else p

    @property
    def npartitions(self):
        if self._parts is None:
            self._parts = self._ingest()
        return len(self._parts)

    @property
    def num_rows(self):
        if self._parts is None:
            self._parts = self._ingest()
        return sum(len(p) for p in self._parts)

    @property
    def schema(self) -> Schema:
        if self._schema is None:
            if self._parts is None:
                self._parts = self._ingest()
            first = self._parts[0] if self._parts else DeviceFrame()
            if self._transform is not None:
                first = self._transform(first)
            self._schema = _schema_of(first)
        return self._schema

    # ------------------------------------------------------------------ egress
    def to_ddf(self, columns=None, **kwargs) -> _Lazy:
        return _Lazy(self, columns)

    def compute(self, **kwargs) -> pd.DataFrame:
        return self.to_ddf().compute()

    def to_cpu(self):
        self.cpu = True
        return self

    def to_parquet(self, output_path, **kwargs):
        import os
        os.makedirs(output_path, exist_ok=True)
        for i, part in enumerate(self.partitions()):
            part.to_pandas().to_parquet(os.path.join(output_path, f"part_{i}.parquet"))
        return output_path
