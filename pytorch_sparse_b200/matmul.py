"""Sparse @ dense (SpMM) and sparse @ sparse (SpSpMM) entry points.

Public names and behaviour follow torch_sparse/matmul.py (`spmm_sum/add/mean/min/max` :9-76, `spmm` :79-91,
`spspmm_sum` :94-111, `spspmm` :118-126, `matmul` :141-163, `SparseTensor` bindings :166-171), and like the
reference's these functions are TorchScript-compatible (`torch.jit.script(spspmm)`, test/test_matmul.py:79): every
native step is a registered operator. A reduction materialises only the cached CSR/CSC views a requested gradient
will read (the reference's rule, :19-25, :46-53).
"""
from typing import Optional, Tuple

import torch
from torch import Tensor

from .tensor import SparseTensor


def spmm_sum(src: SparseTensor, other: Tensor) -> Tensor:
    rowptr, col, value = src.csr()
    st = src.storage
    row, colptr, csr2csc = st._row, st._colptr, st._csr2csc
    row_csc: Optional[Tensor] = None
    if value is not None:
        value = value.to(other.dtype)
        if value.requires_grad:
            row = st.row()
    if other.requires_grad:     # A^T @ grad_out runs on the CSC view
        row = st.row()
        csr2csc = st.csr2csc()
        colptr = st.colptr()
        row_csc = st.row_csc()  # structure-only: gathered once per matrix instead of once per backward
    return torch.ops.tsb200.spmm_sum_csc(row, rowptr, col, value, colptr, csr2csc, other, row_csc)


def spmm_add(src: SparseTensor, other: Tensor) -> Tensor:
    return spmm_sum(src, other)


def spmm_mean(src: SparseTensor, other: Tensor) -> Tensor:
    rowptr, col, value = src.csr()
    st = src.storage
    row, rowcount, colptr, csr2csc = st._row, st._rowcount, st._colptr, st._csr2csc
    row_csc: Optional[Tensor] = None
    if value is not None:
        value = value.to(other.dtype)
        if value.requires_grad:
            row = st.row()
    if other.requires_grad:
        row = st.row()
        rowcount = st.rowcount()
        csr2csc = st.csr2csc()
        colptr = st.colptr()
        row_csc = st.row_csc()
    return torch.ops.tsb200.spmm_mean_csc(row, rowptr, col, value, rowcount, colptr, csr2csc, other, row_csc)


def spmm_min(src: SparseTensor, other: Tensor) -> Tuple[Tensor, Tensor]:
    rowptr, col, value = src.csr()
    if value is not None:
        value = value.to(other.dtype)
    return torch.ops.tsb200.spmm_min(rowptr, col, value, other)


def spmm_max(src: SparseTensor, other: Tensor) -> Tuple[Tensor, Tensor]:
    rowptr, col, value = src.csr()
    if value is not None:
        value = value.to(other.dtype)
    return torch.ops.tsb200.spmm_max(rowptr, col, value, other)


def spmm(src: SparseTensor, other: Tensor, reduce: str = "sum") -> Tensor:
    if reduce == "sum" or reduce == "add":
        return spmm_sum(src, other)
    elif reduce == "mean":
        return spmm_mean(src, other)
    elif reduce == "min":
        return spmm_min(src, other)[0]     # values only; arg_out stays internal
    elif reduce == "max":
        return spmm_max(src, other)[0]
    else:
        raise ValueError("unknown reduce '" + reduce + "'")


def spspmm_sum(src: SparseTensor, other: SparseTensor) -> SparseTensor:
    M, Kd = src.sparse_sizes()
    Kb, N = other.sparse_sizes()
    assert Kd == Kb, "inner dimensions differ"
    rowptr_a, col_a, val_a = src.csr()
    rowptr_b, col_b, val_b = other.csr()
    want_value = val_a is not None or val_b is not None   # a missing side counts as all-ones
    if val_a is not None and val_a.dim() > 1:
        raise RuntimeError("spspmm: multi-dimensional values are not supported")
    if val_b is not None and val_b.dim() > 1:
        raise RuntimeError("spspmm: multi-dimensional values are not supported")
    if val_a is not None and val_b is not None and val_a.dtype != val_b.dtype:
        val_b = val_b.to(val_a.dtype)
    rowptr_c, row_c, col_c, val_c = torch.ops.tsb200.spspmm(rowptr_a, col_a, val_a, rowptr_b, col_b, val_b, M, Kd, N,
                                                            want_value)
    return SparseTensor(row_c, rowptr_c, col_c, val_c, (M, N), True, True)


def spspmm_add(src: SparseTensor, other: SparseTensor) -> SparseTensor:
    return spspmm_sum(src, other)


def spspmm(src: SparseTensor, other: SparseTensor, reduce: str = "sum") -> SparseTensor:
    if reduce == "sum" or reduce == "add":
        return spspmm_sum(src, other)
    elif reduce == "mean" or reduce == "min" or reduce == "max":
        raise NotImplementedError("spspmm with reduce='" + reduce + "'")
    else:
        raise ValueError("unknown reduce '" + reduce + "'")


@torch.jit._overload  # noqa: F811
def matmul(src, other, reduce):  # noqa: F811
    # type: (SparseTensor, Tensor, str) -> Tensor
    pass


@torch.jit._overload  # noqa: F811
def matmul(src, other, reduce):  # noqa: F811
    # type: (SparseTensor, SparseTensor, str) -> SparseTensor
    pass


def matmul(src, other, reduce="sum"):  # noqa: F811
    """`src @ other`: a dense `other` gives a dense result (reduce in sum/add/mean/min/max over each row's
    entries); a SparseTensor `other` gives a SparseTensor (sum only)."""
    if isinstance(other, Tensor):
        return spmm(src, other, reduce)
    elif isinstance(other, SparseTensor):
        return spspmm(src, other, reduce)
    raise ValueError("cannot multiply a SparseTensor with this operand")


SparseTensor.spmm = lambda self, other, reduce="sum": spmm(self, other, reduce)
SparseTensor.spspmm = lambda self, other, reduce="sum": spspmm(self, other, reduce)
SparseTensor.matmul = lambda self, other, reduce="sum": matmul(self, other, reduce)
SparseTensor.__matmul__ = lambda self, other: matmul(self, other, "sum")
