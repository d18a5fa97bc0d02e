"""Stand-in for torch_sparse/testing.py:1-21 with the device list reduced to the GPU (this package has no CPU path)."""
from typing import Any

import torch

reductions = ['sum', 'add', 'mean', 'min', 'max']
dtypes = [torch.half, torch.float, torch.double, torch.int, torch.long, torch.bfloat16]
grad_dtypes = [torch.half, torch.float, torch.double, torch.bfloat16]
devices = [torch.device('cuda:0')]


def tensor(x: Any, dtype: torch.dtype, device: torch.device):
    return None if x is None else torch.tensor(x, dtype=dtype, device=device)
