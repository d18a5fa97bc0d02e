"""Sparse + sparse addition: `spadd(indexA, valueA, indexB, valueB, m, n)` and `add(SparseTensor, SparseTensor)`
(torch_sparse/spadd.py:5-18, torch_sparse/add.py:38-56) — both are "concatenate the entries, coalesce with
sum", so they ride on the native coalesce kernels. `narrow` slices rows (pointer arithmetic) or columns
(torch_sparse/narrow.py:8-77). `mul` / `mul_` / `mul_nnz` and the dense-vector `add` variants
(torch_sparse/mul.py, add.py:21-37, 59-104) complete the row/column scaling family."""
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


def _broadcast_to_nnz(src: SparseTensor, other: Tensor) -> Tensor:
    """`other` of shape [M, 1, ...] (one factor per row) or [1, N, ...] (per column) expanded to one per stored entry
    (torch_sparse/add.py:24-31, mul.py:24-33: gather_csr over rowptr / a gather through col)."""
    if other.dim() >= 2 and other.size(0) == src.size(0) and other.size(1) == 1:
        return other.squeeze(1)[src.storage.row()]
    if other.dim() >= 2 and other.size(0) == 1 and other.size(1) == src.size(1):
        return other.squeeze(0)[src.storage.col()]
    raise ValueError("Size mismatch: Expected size (" + str(src.size(0)) + ", 1, ...) or (1, " + str(src.size(1)) +
                     ", ...), but got size " + str(other.size()) + ".")


@torch.jit._overload  # noqa: F811
def add(src, other):  # noqa: F811
    # type: (SparseTensor, Tensor) -> SparseTensor
    pass


@torch.jit._overload  # noqa: F811
def add(src, other):  # noqa: F811
    # type: (SparseTensor, SparseTensor) -> SparseTensor
    pass


def add(src, other):  # noqa: F811
    """sparse + dense row/column vector (values shifted per row / column; a value-less tensor counts as ones) or
    sparse + sparse (concatenate + native coalesce with sum) — torch_sparse/add.py:21-56. TorchScript-compatible."""
    if isinstance(other, Tensor):
        term = _broadcast_to_nnz(src, other)
        value = src.storage.value()
        if value is None:
            value = term + 1
        else:
            value = term.to(value.dtype) + value
        return src.set_value(value, layout="coo")
    elif isinstance(other, SparseTensor):
        rowA, colA, valueA = src.coo()
        rowB, colB, valueB = other.coo()
        value: Optional[Tensor] = None
        if valueA is not None and valueB is not None:
            value = torch.cat([valueA, valueB], dim=0)
        M = max(src.size(0), other.size(0))
        N = max(src.size(1), other.size(1))
        row, col, value = torch.ops.tsb200.coalesce(torch.cat([rowA, rowB]), torch.cat([colA, colB]), value, M, N,
                                                    "sum")
        return SparseTensor(row, None, col, value, (M, N), True, True)
    else:
        raise NotImplementedError


def add_(src: SparseTensor, other: Tensor) -> SparseTensor:
    term = _broadcast_to_nnz(src, other)
    value = src.storage.value()
    value = term + 1 if value is None else value.add_(term.to(value.dtype))
    return src.set_value_(value, layout="coo")


def add_nnz(src: SparseTensor, other: Tensor, layout: Optional[str] = None) -> SparseTensor:
    value = src.storage.value()
    return src.set_value(other.add(1) if value is None else value.add(other.to(value.dtype)), layout=layout)


def add_nnz_(src: SparseTensor, other: Tensor, layout: Optional[str] = None) -> SparseTensor:
    value = src.storage.value()
    return src.set_value_(other.add(1) if value is None else value.add_(other.to(value.dtype)), layout=layout)


SparseTensor.add = lambda self, other: add(self, other)
SparseTensor.add_ = lambda self, other: add_(self, other)
SparseTensor.add_nnz = lambda self, other, layout=None: add_nnz(self, other, layout)
SparseTensor.add_nnz_ = lambda self, other, layout=None: add_nnz_(self, other, layout)
SparseTensor.__add__ = SparseTensor.add
SparseTensor.__radd__ = SparseTensor.add
SparseTensor.__iadd__ = SparseTensor.add_


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


@torch.jit._overload  # noqa: F811
def mul(src, other):  # noqa: F811
    # type: (SparseTensor, Tensor) -> SparseTensor
    pass


@torch.jit._overload  # noqa: F811
def mul(src, other):  # noqa: F811
    # type: (SparseTensor, SparseTensor) -> SparseTensor
    pass


def mul(src, other):  # noqa: F811
    """Scale the stored values row-wise (`other` of shape [M, 1]) or column-wise ([1, N]) — the GCN
    normalisation step either side of SpMM — or multiply two coalesced SparseTensors entry-wise (the result
    keeps the positions stored in BOTH operands): torch_sparse/mul.py:22-79. Structure and caches of the
    dense-operand branch are shared with `src`. TorchScript-compatible."""
    if isinstance(other, Tensor):
        factor = _broadcast_to_nnz(src, other)
        value = src.storage.value()
        if value is not None:
            factor = factor.to(value.dtype) * value
        return src.set_value(factor, layout="coo")
    assert isinstance(other, SparseTensor)
    if not src.is_coalesced():
        raise ValueError("The `src` tensor is not coalesced")
    if not other.is_coalesced():
        raise ValueError("The `other` tensor is not coalesced")
    rowA, colA, valueA = src.coo()
    rowB, colB, valueB = other.coo()
    if valueA is None or valueB is None:
        raise ValueError("Both sparse tensors must contain values")
    M, N = max(src.size(0), other.size(0)), max(src.size(1), other.size(1))
    row, col, value = torch.cat([rowA, rowB]), torch.cat([colA, colB]), torch.cat([valueA, valueB], dim=0)
    # native stable (row, col) sort of the concatenation: a position stored in both operands shows up as two
    # neighbours, A's entry first
    perm = torch.ops.tsb200.sort_perm(row, col, M, N)
    if perm is not None:
        row, col, value = row[perm], col[perm], value[perm]
    if row.numel() < 2:
        both = torch.zeros(0, dtype=torch.long, device=row.device)
    else:
        both = ((row[1:] == row[:-1]) & (col[1:] == col[:-1])).nonzero().view(-1) + 1
    return SparseTensor(row[both], None, col[both], value[both - 1] * value[both], (M, N), True, True)


def mul_(src: SparseTensor, other: Tensor) -> SparseTensor:
    factor = _broadcast_to_nnz(src, other)
    value = src.storage.value()
    return src.set_value_(factor if value is None else value.mul_(factor.to(value.dtype)), layout="coo")


def mul_nnz(src: SparseTensor, other: Tensor, layout: Optional[str] = None) -> SparseTensor:
    value = src.storage.value()
    return src.set_value(other if value is None else value.mul(other.to(value.dtype)), layout=layout)


def mul_nnz_(src: SparseTensor, other: Tensor, layout: Optional[str] = None) -> SparseTensor:
    value = src.storage.value()
    return src.set_value_(other if value is None else value.mul_(other.to(value.dtype)), layout=layout)


SparseTensor.mul = lambda self, other: mul(self, other)
SparseTensor.mul_ = lambda self, other: mul_(self, other)
SparseTensor.mul_nnz = lambda self, other, layout=None: mul_nnz(self, other, layout)
SparseTensor.mul_nnz_ = lambda self, other, layout=None: mul_nnz_(self, other, layout)
SparseTensor.__mul__ = SparseTensor.mul
SparseTensor.__rmul__ = SparseTensor.mul
SparseTensor.__imul__ = SparseTensor.mul_
