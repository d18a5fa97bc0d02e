"""Reductions of a SparseTensor over its dimensions: `sum / mean / min / max (src, dim=None)`
(torch_sparse/reduce.py:8-93). dim=1 reduces each row's stored values (a segment reduce over `rowptr`),
dim=0 each column's (a segment reduce over `colptr` through the `csr2csc` permutation, instead of the
reference's scatter over `col`), dim>1 reduces trailing value dimensions, dim=None everything.
Without values the entries count as ones. Gradients flow to the stored values as in the reference (whose
segment_csr / scatter are differentiable): ops.segment_reduce is an autograd Function with a one-kernel backward."""
from __future__ import annotations

from typing import Optional

import torch
from torch import Tensor

from . import ops
from .tensor import SparseTensor

_ALL = {"sum": torch.sum, "add": torch.sum, "mean": torch.mean, "min": torch.min, "max": torch.max}


def reduction(src: SparseTensor, dim: Optional[int] = None, reduce: str = "sum") -> Tensor:
    if reduce not in _ALL:
        raise ValueError(f"unknown reduce '{reduce}'")
    value = src.storage.value()
    counting = reduce in ("sum", "add")
    if dim is None:
        if value is not None:
            return _ALL[reduce](value)
        return torch.tensor(src.nnz() if counting else 1, dtype=src.dtype(), device=src.device())
    if dim < 0:
        dim += src.dim()
    if dim > 1:
        if value is None:
            raise ValueError
        red = _ALL[reduce](value, dim=dim - 1)
        return red if isinstance(red, Tensor) else red[0]
    st = src.storage
    if value is None:        # structure only: counts for sum, ones otherwise
        if counting:
            return (st.rowcount() if dim == 1 else st.colcount()).to(src.dtype())
        # one result per column (dim=0) or per row (dim=1)
        return torch.ones(src.size(1) if dim == 0 else src.size(0), dtype=src.dtype(), device=src.device())
    # the segment id of every stored entry is its COO row (dim=1) / column (dim=0): what the backward gathers through
    grad = value.requires_grad and torch.is_grad_enabled()
    if dim == 1:
        return ops.segment_reduce(st.rowptr(), value, reduce, seg=st.row() if grad else None)
    return ops.segment_reduce(st.colptr(), value, reduce, perm=st.csr2csc(), seg=st.col() if grad else None)


def sum(src: SparseTensor, dim: Optional[int] = None) -> Tensor:  # noqa: A001
    return reduction(src, dim, "sum")


def mean(src: SparseTensor, dim: Optional[int] = None) -> Tensor:
    return reduction(src, dim, "mean")


def min(src: SparseTensor, dim: Optional[int] = None) -> Tensor:  # noqa: A001
    return reduction(src, dim, "min")


def max(src: SparseTensor, dim: Optional[int] = None) -> Tensor:  # noqa: A001
    return reduction(src, dim, "max")


SparseTensor.sum = lambda self, dim=None: reduction(self, dim, "sum")
SparseTensor.mean = lambda self, dim=None: reduction(self, dim, "mean")
SparseTensor.min = lambda self, dim=None: reduction(self, dim, "min")
SparseTensor.max = lambda self, dim=None: reduction(self, dim, "max")
