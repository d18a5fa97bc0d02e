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


def masked_select(src: SparseTensor, dim: int, mask: Tensor) -> SparseTensor:
    """Keep the slices of dimension `dim` where `mask` is set (torch_sparse/masked_select.py: same as an
    index_select with the set positions)."""
    return index_select(src, dim, mask.nonzero().view(-1))


def select(src: SparseTensor, dim: int, idx: int) -> SparseTensor:
    """One slice of dimension `dim`, kept as a length-1 dimension (torch_sparse/select.py:4-5)."""
    return src.narrow(dim, idx, 1)


def _getitem(self: SparseTensor, index) -> SparseTensor:
    """`mat[rows, cols]` with ints, slices, index / bool tensors, lists, numpy arrays and one Ellipsis
    (torch_sparse/tensor.py:624-671): each item is a narrow / index_select / masked_select on the next dimension."""
    items = list(index) if isinstance(index, tuple) else [index]

    def is_ellipsis(i):
        return i is Ellipsis

    if sum(1 for i in items if is_ellipsis(i)) > 1:
        raise SyntaxError
    dim, out = 0, self
    while items:
        item = items.pop(0)
        if isinstance(item, (list, tuple)):
            item = torch.tensor(item, device=self.device())
        elif type(item).__module__ == "numpy" and hasattr(item, "dtype"):
            item = torch.from_numpy(item).to(self.device())
        if isinstance(item, int) and not isinstance(item, bool):
            out = select(out, dim, item)
            dim += 1
        elif isinstance(item, slice):
            if item.step is not None:
                raise ValueError("Step parameter not yet supported.")
            start = 0 if item.start is None else item.start
            start = self.size(dim) + start if start < 0 else start
            stop = self.size(dim) if item.stop is None else item.stop
            stop = self.size(dim) + stop if stop < 0 else stop
            out = out.narrow(dim, start, max(stop - start, 0))
            dim += 1
        elif torch.is_tensor(item):
            if item.dtype == torch.bool:
                out = masked_select(out, dim, item)
                dim += 1
            elif item.dtype == torch.long:
                out = index_select(out, dim, item)
                dim += 1
        elif is_ellipsis(item):
            if self.dim() - len(items) < dim:
                raise SyntaxError
            dim = self.dim() - len(items)
        else:
            raise SyntaxError
    return out


SparseTensor.masked_select = lambda self, dim, mask: masked_select(self, dim, mask)
SparseTensor.select = lambda self, dim, idx: select(self, dim, idx)
SparseTensor.__getitem__ = _getitem
SparseTensor.index_select = lambda self, dim, idx: index_select(self, dim, idx)
SparseTensor.index_select_nnz = lambda self, idx, layout=None: index_select_nnz(self, idx, layout)
SparseTensor.to_symmetric = lambda self, reduce="sum": to_symmetric(self, reduce)
