"""Drop-in alias: `import nvtabular as nvt` resolves to the B200 engine, so
pipelines written against the reference API (nvtabular/__init__.py:19-56 of the
reference) run unchanged on the hot path this repo implements."""
import sys

import nvtabular_b200 as _impl
from nvtabular_b200 import *  # noqa: F401,F403
from nvtabular_b200 import (ColumnSchema, ColumnSelector, Dataset, Schema, Workflow,  # noqa: F401
                            Shuffle, WorkflowNode, ops)

__version__ = _impl.__version__
sys.modules.setdefault("nvtabular.ops", ops)
