"""Sparse @ dense (SpMM) and sparse @ sparse (SpSpMM) entry points.

Public names and behaviour follow torch_sparse/matmul.py (`spmm_sum/add/mean/min/max` :9-76, `spmm` :79-91,
`spspmm_sum` :94-111, `spspmm` :118-126, `matmul` :141-163, `SparseTensor` bindings :166-171). The
implementation is table driven: one helper decides which cached CSR/CSC views a reduction needs — the
reference's rule (:19-25, :46-53) is "only what a requested gradient will read" — and hands them to the
operator layer in ops.py.
"""
from __future__ import annotations

from typing import Callable, Dict, Tuple, Union

from torch import Tensor

from . import ops
from .tensor import SparseTensor

_SUM_LIKE = ("sum", "add")
_ARG_LIKE = ("min", "max")


def _views_for_backward(src: SparseTensor, value, other: Tensor, with_rowcount: bool) -> dict:
    """Cached views to pass along; anything not yet materialised stays None unless a gradient needs it."""
    st = src.storage
    views = dict(row=st._row, rowcount=st._rowcount, colptr=st._colptr, csr2csc=st._csr2csc, row_csc=None)
    grad_value = value is not None and value.requires_grad
    grad_dense = other.requires_grad
    if grad_value or grad_dense:
        views["row"] = st.row()
    if grad_dense:                      # A^T @ grad_out runs on the CSC view
        views["csr2csc"] = st.csr2csc()
        views["colptr"] = st.colptr()
        views["row_csc"] = st.row_csc()  # structure-only: gathered once per matrix instead of once per backward
        if with_rowcount:
            views["rowcount"] = st.rowcount()
    return views


def _run_spmm(src: SparseTensor, other: Tensor, reduce: str):
    rowptr, col, value = src.csr()
    if value is not None:
        value = value.to(other.dtype)
    if reduce in _ARG_LIKE:
        op = ops.spmm_min if reduce == "min" else ops.spmm_max
        return op(rowptr, col, value, other)
    v = _views_for_backward(src, value, other, with_rowcount=(reduce == "mean"))
    if reduce == "mean":
        return ops.spmm_mean(v["row"], rowptr, col, value, v["rowcount"], v["colptr"], v["csr2csc"], other,
                             row_csc=v["row_csc"])
    return ops.spmm_sum(v["row"], rowptr, col, value, v["colptr"], v["csr2csc"], other, row_csc=v["row_csc"])


def spmm_sum(src: SparseTensor, other: Tensor) -> Tensor:
    return _run_spmm(src, other, "sum")


def spmm_add(src: SparseTensor, other: Tensor) -> Tensor:
    return _run_spmm(src, other, "sum")


def spmm_mean(src: SparseTensor, other: Tensor) -> Tensor:
    return _run_spmm(src, other, "mean")


def spmm_min(src: SparseTensor, other: Tensor) -> Tuple[Tensor, Tensor]:
    return _run_spmm(src, other, "min")


def spmm_max(src: SparseTensor, other: Tensor) -> Tuple[Tensor, Tensor]:
    return _run_spmm(src, other, "max")


def spmm(src: SparseTensor, other: Tensor, reduce: str = "sum") -> Tensor:
    if reduce in _SUM_LIKE:
        return _run_spmm(src, other, "sum")
    if reduce == "mean":
        return _run_spmm(src, other, "mean")
    if reduce in _ARG_LIKE:
        return _run_spmm(src, other, reduce)[0]     # values only; arg_out stays internal
    raise ValueError(f"unknown reduce '{reduce}'")


def spspmm_sum(src: SparseTensor, other: SparseTensor) -> SparseTensor:
    M, Kd = src.sparse_sizes()
    Kb, N = other.sparse_sizes()
    assert Kd == Kb, f"inner dimensions differ: {Kd} vs {Kb}"
    rowptr_a, col_a, val_a = src.csr()
    rowptr_b, col_b, val_b = other.csr()
    want_value = val_a is not None or val_b is not None   # a missing side counts as all-ones
    for v in (val_a, val_b):
        if v is not None and v.dim() > 1:
            raise RuntimeError("spspmm: multi-dimensional values are not supported")
    if val_a is not None and val_b is not None and val_a.dtype != val_b.dtype:
        val_b = val_b.to(val_a.dtype)
    rowptr_c, row_c, col_c, val_c = ops.spspmm(rowptr_a, col_a, val_a, rowptr_b, col_b, val_b, M, Kd, N, want_value)
    return SparseTensor(row=row_c, rowptr=rowptr_c, col=col_c, value=val_c, sparse_sizes=(M, N),
                        is_sorted=True, trust_data=True)


def spspmm_add(src: SparseTensor, other: SparseTensor) -> SparseTensor:
    return spspmm_sum(src, other)


def spspmm(src: SparseTensor, other: SparseTensor, reduce: str = "sum") -> SparseTensor:
    if reduce in _SUM_LIKE:
        return spspmm_sum(src, other)
    if reduce in ("mean",) + _ARG_LIKE:
        raise NotImplementedError(f"spspmm with reduce='{reduce}'")
    raise ValueError(f"unknown reduce '{reduce}'")


def matmul(src: SparseTensor, other: Union[Tensor, SparseTensor], reduce: str = "sum"):
    """`src @ other`: a dense `other` gives a dense result (reduce in sum/add/mean/min/max over each row's
    entries); a SparseTensor `other` gives a SparseTensor (sum only)."""
    if isinstance(other, SparseTensor):
        return spspmm(src, other, reduce)
    if isinstance(other, Tensor):
        return spmm(src, other, reduce)
    raise ValueError(f"cannot multiply a SparseTensor with {type(other).__name__}")


_BINDINGS: Dict[str, Callable] = {
    "spmm": lambda self, other, reduce="sum": spmm(self, other, reduce),
    "spspmm": lambda self, other, reduce="sum": spspmm(self, other, reduce),
    "matmul": lambda self, other, reduce="sum": matmul(self, other, reduce),
    "__matmul__": lambda self, other: matmul(self, other, "sum"),
}
for _name, _fn in _BINDINGS.items():
    setattr(SparseTensor, _name, _fn)
