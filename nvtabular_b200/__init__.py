"""nvtabular_b200 — a B200-native engine behind the nvtabular.ops operator API
for the Categorify / FillMissing / Normalize / HashBucket / JoinGroupby /
TargetEncoding hot path.  Python is host glue; all row-level
work happens in hand-written sm_100a kernels reached through the C-ABI of
include/nvtb200.h (nvtabular_b200/lib/libnvtb200.so).
"""
__version__ = "0.1.0"

from .column import Column, DeviceFrame  # noqa: F401
