"""Index bookkeeping for CPU-resident SparseStorage objects.

A SparseStorage may be *constructed* from CPU tensors (e.g. `SparseTensor(...).cuda()` as PyG
does); ordering and pointer arrays for such host-resident data are derived here with stock
PyTorch index ops. This is construction convenience only — every arithmetic operator of the hot
path (spmm, value gradients, coalesce reductions, spspmm) exists solely as a CUDA kernel in
libtsb200 and raises for CPU tensors.
"""
from typing import Optional, Tuple

import torch
from torch import Tensor


def ind2ptr(ind: Tensor, M: int) -> Tensor:
    return torch.searchsorted(ind, torch.arange(M + 1, dtype=ind.dtype, device=ind.device))


def ptr2ind(ptr: Tensor, E: int) -> Tensor:
    counts = ptr[1:] - ptr[:-1]
    return torch.repeat_interleave(torch.arange(counts.numel(), dtype=ptr.dtype, device=ptr.device), counts,
                                   output_size=E)


def sort_perm(row: Tensor, col: Tensor, N: int) -> Optional[Tensor]:
    key = row * N + col
    if key.numel() < 2 or bool((key[1:] >= key[:-1]).all()):
        return None
    return torch.sort(key, stable=True)[1]


def csr2csc(row: Tensor, col: Tensor, M: int) -> Tensor:
    return torch.sort(col * M + row, stable=True)[1]


def is_coalesced(row: Tensor, col: Tensor, N: int) -> bool:
    key = row * N + col
    return key.numel() < 2 or bool((key[1:] > key[:-1]).all())
