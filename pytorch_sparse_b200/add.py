"""Sparse + sparse addition: `spadd(indexA, valueA, indexB, valueB, m, n)` and `add(SparseTensor, SparseTensor)`
(torch_sparse/spadd.py:5-18, torch_sparse/add.py:38-56) — both are "concatenate the entries, coalesce with
sum", so they ride on the native coalesce kernels. `narrow` slices rows (pointer arithmetic) or columns
(torch_sparse/narrow.py:8-77)."""
from __future__ import annotations

from typing import Optional

import torch
from torch import Tensor

from . import ops
from .storage import SparseStorage
from .tensor import SparseTensor


def spadd(indexA, valueA, indexB, valueB, m, n):
    row = torch.cat([indexA[0], indexB[0]])
    col = torch.cat([indexA[1], indexB[1]])
    value = torch.cat([valueA, valueB], dim=0)
    row, col, value = ops.coalesce(row, col, value, m, n, "add")
    return torch.stack([row, col], dim=0), value


def add(src: SparseTensor, other: SparseTensor) -> SparseTensor:
    if not isinstance(other, SparseTensor):
        raise NotImplementedError("only SparseTensor + SparseTensor is on the sparse-matmul path")
    rowA, colA, valueA = src.coo()
    rowB, colB, valueB = other.coo()
    value: Optional[Tensor] = None
    if valueA is not None and valueB is not None:
        value = torch.cat([valueA, valueB], dim=0)
    M = max(src.size(0), other.size(0))
    N = max(src.size(1), other.size(1))
    row, col, value = ops.coalesce(torch.cat([rowA, rowB]), torch.cat([colA, colB]), value, M, N, "sum")
    return SparseTensor(row=row, col=col, value=value, sparse_sizes=(M, N), is_sorted=True, trust_data=True)


SparseTensor.add = lambda self, other: add(self, other)
SparseTensor.__add__ = lambda self, other: add(self, other)


def narrow(src: SparseTensor, dim: int, start: int, length: int) -> SparseTensor:
    if dim < 0:
        dim += src.dim()
    if start < 0:
        start += src.size(dim)
    st = src.storage
    if dim == 0:
        rowptr, col, value = src.csr()
        sub = rowptr[start:start + length + 1]
        lo = int(sub[0])
        hi = int(sub[-1])
        row = st._row
        storage = SparseStorage(row=None if row is None else row[lo:hi] - start, rowptr=sub - lo, col=col[lo:hi],
                                value=None if value is None else value[lo:hi],
                                sparse_sizes=(length, src.sparse_size(1)),
                                rowcount=None if st._rowcount is None else st._rowcount[start:start + length],
                                is_sorted=True, trust_data=True)
        return src.from_storage(storage)
    if dim == 1:
        row, col, value = src.coo()
        keep = (col >= start) & (col < start + length)
        colptr = st._colptr
        if colptr is not None:
            colptr = colptr[start:start + length + 1]
            colptr = colptr - colptr[0]
        storage = SparseStorage(row=row[keep], col=col[keep] - start, value=None if value is None else value[keep],
                                sparse_sizes=(src.sparse_size(0), length), colptr=colptr,
                                colcount=None if st._colcount is None else st._colcount[start:start + length],
                                is_sorted=True, trust_data=True)
        return src.from_storage(storage)
    value = st.value()
    if value is None:
        raise ValueError
    return src.set_value(value.narrow(dim - 1, start, length), layout="coo")


SparseTensor.narrow = lambda self, dim, start, length: narrow(self, dim, start, length)


def mul(src: SparseTensor, other: Tensor) -> SparseTensor:
    """Scale the stored values row-wise (`other` of shape [M, 1]) or column-wise ([1, N]) — the GCN
    normalisation step either side of SpMM (torch_sparse/mul.py:22-40, dense-operand branch). A pure
    gather-multiply on the values; structure and caches are shared with `src`."""
    if not isinstance(other, Tensor):
        raise NotImplementedError("sparse * sparse is not on the sparse-matmul path")
    row, col, value = src.coo()
    if other.dim() >= 2 and other.size(0) == src.size(0) and other.size(1) == 1:
        factor = other.squeeze(1)[row]
    elif other.dim() >= 2 and other.size(0) == 1 and other.size(1) == src.size(1):
        factor = other.squeeze(0)[col]
    else:
        raise ValueError(f"Size mismatch: Expected size ({src.size(0)}, 1, ...) or (1, {src.size(1)}, ...), "
                         f"but got size {tuple(other.size())}.")
    value = factor if value is None else factor.to(value.dtype) * value
    return src.set_value(value, layout="coo")


SparseTensor.mul = lambda self, other: mul(self, other)
SparseTensor.__mul__ = lambda self, other: mul(self, other)
