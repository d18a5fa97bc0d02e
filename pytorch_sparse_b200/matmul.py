"""Sparse @ dense (SpMM) and sparse @ sparse (SpSpMM) entry points.

Same functions, dispatch rules and `SparseTensor` method bindings as torch_sparse/matmul.py:
`spmm_{sum,add,mean,min,max}` (:9-76), `spmm` (:79-91), `spspmm_sum` (:94-111), `spspmm`
(:118-126), `matmul` (:141-163), bindings (:166-171). Which cached views are materialised
depends on who needs gradients, exactly as in the reference (:19-25, :46-53).
"""
from __future__ import annotations

from typing import Tuple, Union

import torch
from torch import Tensor

from . import ops
from .tensor import SparseTensor


def _csr_for(src: SparseTensor, other: Tensor, want_rowcount: bool):
    st = src.storage
    rowptr, col, value = src.csr()
    if value is not None:
        value = value.to(other.dtype)
    row, rowcount, colptr, csr2csc = st._row, st._rowcount, st._colptr, st._csr2csc
    if value is not None and value.requires_grad:
        row = st.row()
    if other.requires_grad:
        row = st.row()
        csr2csc = st.csr2csc()
        colptr = st.colptr()
        if want_rowcount:
            rowcount = st.rowcount()
    return row, rowptr, col, value, rowcount, colptr, csr2csc


def spmm_sum(src: SparseTensor, other: Tensor) -> Tensor:
    row, rowptr, col, value, _, colptr, csr2csc = _csr_for(src, other, False)
    return ops.spmm_sum(row, rowptr, col, value, colptr, csr2csc, other)


def spmm_add(src: SparseTensor, other: Tensor) -> Tensor:
    return spmm_sum(src, other)


def spmm_mean(src: SparseTensor, other: Tensor) -> Tensor:
    row, rowptr, col, value, rowcount, colptr, csr2csc = _csr_for(src, other, True)
    return ops.spmm_mean(row, rowptr, col, value, rowcount, colptr, csr2csc, other)


def spmm_min(src: SparseTensor, other: Tensor) -> Tuple[Tensor, Tensor]:
    rowptr, col, value = src.csr()
    if value is not None:
        value = value.to(other.dtype)
    return ops.spmm_min(rowptr, col, value, other)


def spmm_max(src: SparseTensor, other: Tensor) -> Tuple[Tensor, Tensor]:
    rowptr, col, value = src.csr()
    if value is not None:
        value = value.to(other.dtype)
    return ops.spmm_max(rowptr, col, value, other)


def spmm(src: SparseTensor, other: Tensor, reduce: str = "sum") -> Tensor:
    if reduce in ("sum", "add"):
        return spmm_sum(src, other)
    if reduce == "mean":
        return spmm_mean(src, other)
    if reduce == "min":
        return spmm_min(src, other)[0]
    if reduce == "max":
        return spmm_max(src, other)[0]
    raise ValueError


def spspmm_sum(src: SparseTensor, other: SparseTensor) -> SparseTensor:
    assert src.sparse_size(1) == other.sparse_size(0)
    rowptr_a, col_a, val_a = src.csr()
    rowptr_b, col_b, val_b = other.csr()
    M, Kd, N = src.sparse_size(0), src.sparse_size(1), other.sparse_size(1)
    want_value = src.has_value() or other.has_value()
    if want_value:  # like torch.sparse.mm on the COO tensors: a missing value means ones
        if val_a is not None and val_a.dim() > 1 or val_b is not None and val_b.dim() > 1:
            raise RuntimeError("spspmm: multi-dimensional values are not supported")
        if val_a is None:
            val_a = None  # implicit ones (handled in the kernel)
        if val_a is not None and val_b is not None and val_a.dtype != val_b.dtype:
            val_b = val_b.to(val_a.dtype)
    rowptr_c, row_c, col_c, val_c = ops.spspmm(rowptr_a, col_a, val_a, rowptr_b, col_b, val_b, M, Kd, N,
                                               want_value)
    return SparseTensor(row=row_c, rowptr=rowptr_c, col=col_c, value=val_c, sparse_sizes=(M, N),
                        is_sorted=True, trust_data=True)


def spspmm_add(src: SparseTensor, other: SparseTensor) -> SparseTensor:
    return spspmm_sum(src, other)


def spspmm(src: SparseTensor, other: SparseTensor, reduce: str = "sum") -> SparseTensor:
    if reduce in ("sum", "add"):
        return spspmm_sum(src, other)
    if reduce in ("mean", "min", "max"):
        raise NotImplementedError
    raise ValueError


def matmul(src: SparseTensor, other: Union[Tensor, SparseTensor], reduce: str = "sum"):
    """Matrix product of a sparse tensor with a dense tensor (reduce in sum/add/mean/min/max) or with
    another sparse tensor (sum only)."""
    if isinstance(other, Tensor):
        return spmm(src, other, reduce)
    if isinstance(other, SparseTensor):
        return spspmm(src, other, reduce)
    raise ValueError


SparseTensor.spmm = lambda self, other, reduce="sum": spmm(self, other, reduce)
SparseTensor.spspmm = lambda self, other, reduce="sum": spspmm(self, other, reduce)
SparseTensor.matmul = lambda self, other, reduce="sum": matmul(self, other, reduce)
SparseTensor.__matmul__ = lambda self, other: matmul(self, other, "sum")
