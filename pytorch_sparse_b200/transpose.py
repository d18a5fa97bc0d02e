"""Transposition: `SparseTensor.t()` and the functional `transpose(index, value, m, n)`
(torch_sparse/transpose.py:7-62)."""
from __future__ import annotations

import torch

from . import ops
from .storage import SparseStorage
from .tensor import SparseTensor


def t(src: SparseTensor) -> SparseTensor:
    st = src.storage
    perm = st.csr2csc()
    row, col, value = src.coo()
    M, N = st.sparse_sizes()
    storage = SparseStorage(row=col[perm], rowptr=st._colptr, col=row[perm],
                            value=None if value is None else value[perm], sparse_sizes=(N, M),
                            rowcount=st._colcount, colptr=st._rowptr, colcount=st._rowcount,
                            csr2csc=st._csc2csr, csc2csr=perm, is_sorted=True, trust_data=True)
    return src.from_storage(storage)


SparseTensor.t = lambda self: t(self)


def transpose(index, value, m, n, coalesced=True):
    """Swap the two index rows; with `coalesced=True` also sort + merge duplicates (sum)."""
    row, col = index[1], index[0]
    if coalesced:
        row, col, value = ops.coalesce(row, col, value, n, m, "add")
    return torch.stack([row, col], dim=0), value
