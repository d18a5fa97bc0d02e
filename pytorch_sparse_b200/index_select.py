"""Row / column gather and symmetrisation — the mini-batch step in front of SpMM and the other
"concatenate + coalesce" caller:

  * `index_select(src, dim, idx)` / `index_select_nnz(src, idx, layout)` (torch_sparse/index_select.py:9-99),
  * `SparseTensor.to_symmetric(reduce)` (torch_sparse/tensor.py:404-438).

Row gathers are pointer arithmetic plus one gather of `col`/`value` (`ptr2ind` on the native kernel gives the
new COO rows); column gathers additionally need the new row-major order, which is the native (row, col) sort;
`to_symmetric` is `cat([A, A^T])` + the native coalesce with the requested reduction.
"""
from __future__ import annotations

from typing import Optional

import torch
from torch import Tensor

from . import ops
from .storage import SparseStorage
from .tensor import SparseTensor


def _expand_ptr(old_ptr: Tensor, count: Tensor, idx: Tensor):
    """new pointer, new major index per entry and the gather permutation for selecting the major slices `idx`
    (torch_sparse/index_select.py:18-31)."""
    count = count[idx]
    ptr = count.new_zeros(idx.size(0) + 1)
    torch.cumsum(count, dim=0, out=ptr[1:])
    E = int(ptr[-1])
    major = ops.ptr2ind(ptr, E) if ptr.is_cuda else torch.repeat_interleave(
        torch.arange(idx.size(0), device=ptr.device), count)
    # entry e of new slice s sits at old_ptr[idx[s]] + (e - ptr[s])
    perm = torch.arange(E, device=ptr.device) + (old_ptr[idx] - ptr[:-1])[major]
    return ptr, count, major, perm


def index_select(src: SparseTensor, dim: int, idx: Tensor) -> SparseTensor:
    dim = src.dim() + dim if dim < 0 else dim
    assert idx.dim() == 1
    if dim == 0:
        old_rowptr, col, value = src.csr()
        rowptr, rowcount, row, perm = _expand_ptr(old_rowptr, src.storage.rowcount(), idx)
        storage = SparseStorage(row=row, rowptr=rowptr, col=col[perm], value=None if value is None else value[perm],
                                sparse_sizes=(idx.size(0), src.sparse_size(1)), rowcount=rowcount,
                                is_sorted=True, trust_data=True)
        return src.from_storage(storage)
    if dim == 1:
        old_colptr, row, value = src.csc()
        colptr, colcount, col, perm = _expand_ptr(old_colptr, src.storage.colcount(), idx)
        row = row[perm]
        M, Nn = src.sparse_size(0), idx.size(0)
        # back to row-major order: the native (row, col) sort (the reference argsorts idx.size(0) * row + col)
        csc2csr = ops.sort_perm(row, col, M, Nn) if row.is_cuda else torch.argsort(Nn * row + col)
        if csc2csr is None:  # already row-major
            csc2csr = torch.arange(row.numel(), device=row.device)
        if value is not None:
            value = value[perm][csc2csr]
        storage = SparseStorage(row=row[csc2csr], col=col[csc2csr], value=value, sparse_sizes=(M, Nn),
                                colptr=colptr, colcount=colcount, csc2csr=csc2csr, is_sorted=True, trust_data=True)
        return src.from_storage(storage)
    value = src.storage.value()
    if value is None:
        raise ValueError
    return src.set_value(value.index_select(dim - 1, idx), layout="coo")


def index_select_nnz(src: SparseTensor, idx: Tensor, layout: Optional[str] = None) -> SparseTensor:
    assert idx.dim() == 1
    if layout == "csc":
        idx = src.storage.csc2csr()[idx]
    row, col, value = src.coo()
    return SparseTensor(row=row[idx], col=col[idx], value=None if value is None else value[idx],
                        sparse_sizes=src.sparse_sizes(), is_sorted=True)


def to_symmetric(src: SparseTensor, reduce: str = "sum") -> SparseTensor:
    """A ∪ Aᵀ with duplicate (row, col) entries reduced — on square inputs the diagonal is hit twice, exactly as
    in the reference (torch_sparse/tensor.py:404-438)."""
    N = max(src.size(0), src.size(1))
    row, col, value = src.coo()
    r2, c2 = torch.cat([row, col]), torch.cat([col, row])
    v2 = None if value is None else torch.cat([value, value])
    r2, c2, v2 = ops.coalesce(r2, c2, v2, N, N, reduce)
    return SparseTensor(row=r2, col=c2, value=v2, sparse_sizes=(N, N), is_sorted=True, trust_data=True)


SparseTensor.index_select = lambda self, dim, idx: index_select(self, dim, idx)
SparseTensor.index_select_nnz = lambda self, idx, layout=None: index_select_nnz(self, idx, layout)
SparseTensor.to_symmetric = lambda self, reduce="sum": to_symmetric(self, reduce)
