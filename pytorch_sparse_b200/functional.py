"""The legacy functional `(index, value)` API: `coalesce`, `spmm`, `spspmm`
(torch_sparse/coalesce.py:5-25, torch_sparse/spmm.py:5-31, torch_sparse/spspmm.py:6-33)."""
from __future__ import annotations

import torch
from torch import Tensor

from . import ops
from .matmul import matmul
from .tensor import SparseTensor


def coalesce(index, value, m, n, op="add"):
    """Row-major sort of the entries and reduction of duplicates with `op` in
    {"add"/"sum", "mean", "min", "max"}; `value` may be None or have trailing dimensions."""
    row, col, value = ops.coalesce(index[0], index[1], value, m, n, op)
    return torch.stack([row, col], dim=0), value


def spmm(index: Tensor, value: Tensor, m: int, n: int, matrix: Tensor) -> Tensor:
    """Sparse (m x n, COO `index`/`value`, any order, duplicates allowed) times dense `matrix`.

    The reference materialises an E x K gather and scatter-adds it (torch_sparse/spmm.py:27-29);
    here the entries are ordered once (skipped when already sorted) and the CSR SpMM kernel runs,
    duplicates simply accumulate. Gradients flow to `value` and `matrix`."""
    assert n == matrix.size(-2)
    squeeze = matrix.dim() == 1
    if squeeze:
        matrix = matrix.unsqueeze(-1)
    src = SparseTensor(row=index[0], col=index[1], value=value, sparse_sizes=(m, n), is_sorted=False,
                       trust_data=True)
    return matmul(src, matrix, "sum")


def spspmm(indexA, valueA, indexB, valueB, m, k, n, coalesced=False):
    """Sparse (m x k) times sparse (k x n) -> (index, value) of the sorted, coalesced product.

    The reference passes `is_sorted=not coalesced` and then hands COO to torch.sparse.mm, which tolerates
    unsorted / uncoalesced indices either way (torch_sparse/spspmm.py:25-28). Here the product runs on the
    CSR views, which need row-major order, so the order is always verified: the native check is one kernel
    and returns without sorting when the entries are already ordered. Duplicate entries need no merging —
    their products accumulate, as in torch.sparse.mm."""
    A = SparseTensor(row=indexA[0], col=indexA[1], value=valueA, sparse_sizes=(m, k), is_sorted=False)
    B = SparseTensor(row=indexB[0], col=indexB[1], value=valueB, sparse_sizes=(k, n), is_sorted=False)
    C = matmul(A, B)
    row, col, value = C.coo()
    return torch.stack([row, col], dim=0), value
