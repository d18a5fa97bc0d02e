"""Transposition of a SparseTensor (`t()`) and of an `(index, value)` pair (`transpose`), cf.
torch_sparse/transpose.py:7-62. The transposed storage is the CSC view of the original, so every cache
the source already holds is reused with rows and columns swapped; nothing is re-sorted."""
from __future__ import annotations

import torch

from . import ops
from .storage import SparseStorage
from .tensor import SparseTensor


def t(src: SparseTensor) -> SparseTensor:
    st = src.storage
    to_csc = st.csr2csc()                       # entry order of the transposed (column-major) layout
    M, N = st.sparse_sizes()
    value = st.value()
    swapped = SparseStorage(
        row=st.col()[to_csc], col=st.row_csc(), value=value[to_csc] if value is not None else None,
        sparse_sizes=(N, M),
        # CSR caches of the transpose are the CSC caches of the source and vice versa
        rowptr=st._colptr, rowcount=st._colcount, colptr=st._rowptr, colcount=st._rowcount,
        csr2csc=st._csc2csr, csc2csr=to_csc,
        is_sorted=True, trust_data=True)
    return src.from_storage(swapped)


SparseTensor.t = t


def transpose(index, value, m, n, coalesced=True):
    """(index, value) of the n x m transpose. `coalesced=True` (default) also orders the result row-major
    and sums duplicate entries; `False` only swaps the two index rows."""
    new_row, new_col = index[1], index[0]
    if coalesced:
        new_row, new_col, value = ops.coalesce(new_row, new_col, value, n, m, "add")
    return torch.stack([new_row, new_col], dim=0), value
