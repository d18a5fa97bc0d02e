"""`torch_sparse := pytorch_sparse_b200` — import shim used ONLY to run the reference's own, unmodified test files
against this package (tests/test_reference_suite_gpu.py). The reference's tests import `torch_sparse`,
`torch_sparse.matmul`, `torch_sparse.tensor`, `torch_sparse.storage` and `torch_sparse.testing`; each name is
bound to the corresponding module of pytorch_sparse_b200 (testing.py is the shim's own: devices = cuda only)."""
import importlib
import os
import sys

os.environ["TSB200_REGISTER_TORCH_SPARSE"] = "1"   # this shim IS the importable `torch_sparse`

import pytorch_sparse_b200 as _pkg
from pytorch_sparse_b200 import *  # noqa: F401,F403
from pytorch_sparse_b200 import SparseStorage, SparseTensor, __version__  # noqa: F401

for _name in ("matmul", "tensor", "storage", "transpose", "add", "reduce", "index_select", "functional", "ops"):
    sys.modules[f"{__name__}.{_name}"] = importlib.import_module(f"pytorch_sparse_b200.{_name}")
# the reference keeps these in modules of their own
sys.modules[f"{__name__}.coalesce"] = sys.modules[f"{__name__}.functional"]
sys.modules[f"{__name__}.spmm"] = sys.modules[f"{__name__}.functional"]
sys.modules[f"{__name__}.spspmm"] = sys.modules[f"{__name__}.functional"]
sys.modules[f"{__name__}.mul"] = sys.modules[f"{__name__}.add"]
sys.modules[f"{__name__}.narrow"] = sys.modules[f"{__name__}.add"]
