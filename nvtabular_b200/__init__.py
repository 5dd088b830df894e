"""nvtabular_b200 — a B200-native engine behind the nvtabular.ops operator API
for the Categorify / FillMissing / Normalize / HashBucket / JoinGroupby /
TargetEncoding hot path.  Python is host glue; all row-level
work happens in hand-written sm_100a kernels reached through the C-ABI of
include/nvtb200.h (nvtabular_b200/lib/libnvtb200.so).

    import nvtabular_b200 as nvt            # or: import nvtabular as nvt
    cats = ["C1", "C2"] >> nvt.ops.Categorify()
    conts = ["I1"] >> nvt.ops.FillMissing() >> nvt.ops.Normalize()
    wf = nvt.Workflow(cats + conts).fit(nvt.Dataset(df))
    out = wf.transform(nvt.Dataset(df)).to_ddf().compute()
"""
__version__ = "0.1.0"

from . import ops  # noqa: F401
from . import serialize  # noqa: F401
from .column import Column, DeviceFrame  # noqa: F401
from .dataset import Dataset  # noqa: F401
from .graph import ColumnSchema, ColumnSelector, Node, Schema, Tags  # noqa: F401
from .workflow import Workflow  # noqa: F401

WorkflowNode = Node


class Shuffle:
    """merlin.io.Shuffle as the reference's benchmark passes it to Dataset.to_parquet
    (bench/examples/dask-nvtabular-criteo-benchmark.py:225-237)"""
    PER_PARTITION = "PER_PARTITION"
    PER_WORKER = "PER_WORKER"
    FULL = "FULL"
