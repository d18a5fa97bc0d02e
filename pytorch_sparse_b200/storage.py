"""SparseStorage — canonical row-major COO plus lazily derived CSR / CSC views.

API-compatible with `torch_sparse.storage.SparseStorage` (torch_sparse/storage.py:20-801): same
constructor arguments, the same private attribute names (`_row`, `_rowptr`, `_col`, `_value`,
`_rowcount`, `_colptr`, `_colcount`, `_csr2csc`, `_csc2csr`) that `matmul.py` and the reference's
tests read, and the same lazy-cache semantics. Differences, all on the GPU side:
  * sort-on-construct, csr2csc/colptr and coalesce run as fused libtsb200 kernels
    (one stable radix sort over only the significant key bits, no E-sized int64 temporaries),
  * `trust_data=True` (or CUDA tensors + explicit sizes) skips the host-synchronising bound checks.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
from torch import Tensor

from . import _host_index as host
from . import ops

_CACHE_KEYS = ("rowcount", "colptr", "colcount", "csr2csc", "csc2csr")


def _long_vector(t: Optional[Tensor], name: str, numel: Optional[int], device) -> Optional[Tensor]:
    if t is None:
        return None
    assert t.dtype == torch.long, f"{name} must be int64"
    assert t.dim() == 1, f"{name} must be 1-dimensional"
    assert t.device == device, f"{name} lives on {t.device}, expected {device}"
    if numel is not None:
        assert t.numel() == numel, f"{name} has {t.numel()} entries, expected {numel}"
    return t.contiguous()


class SparseStorage:
    __slots__ = ("_row", "_rowptr", "_col", "_value", "_sparse_sizes", "_rowcount", "_colptr", "_colcount",
                 "_csr2csc", "_csc2csr", "_row_csc")

    def __init__(self, row: Optional[Tensor] = None, rowptr: Optional[Tensor] = None,
                 col: Optional[Tensor] = None, value: Optional[Tensor] = None,
                 sparse_sizes: Optional[Tuple[Optional[int], Optional[int]]] = None,
                 rowcount: Optional[Tensor] = None, colptr: Optional[Tensor] = None,
                 colcount: Optional[Tensor] = None, csr2csc: Optional[Tensor] = None,
                 csc2csr: Optional[Tensor] = None, is_sorted: bool = False, trust_data: bool = False):
        assert row is not None or rowptr is not None
        assert col is not None
        dev = col.device
        col = _long_vector(col, "col", None, dev)
        E = col.numel()

        m_hint = None if sparse_sizes is None else sparse_sizes[0]
        n_hint = None if sparse_sizes is None else sparse_sizes[1]
        if m_hint is None:
            if rowptr is not None:
                M = rowptr.numel() - 1
            else:
                M = int(row.max()) + 1 if E > 0 else 0
        else:
            M = int(m_hint)
            if rowptr is not None:
                assert rowptr.numel() - 1 == M
            elif E > 0 and not trust_data:
                assert int(row.max()) < M
        if n_hint is None:
            N = int(col.max()) + 1 if E > 0 else 0
        else:
            N = int(n_hint)
            if E > 0 and not trust_data:
                assert int(col.max()) < N

        self._row = _long_vector(row, "row", E, dev)
        self._rowptr = _long_vector(rowptr, "rowptr", M + 1, dev)
        self._col = col
        if value is not None:
            assert value.device == dev
            assert value.size(0) == E
            value = value.contiguous()
        self._value = value
        self._sparse_sizes = (M, N)
        self._rowcount = _long_vector(rowcount, "rowcount", M, dev)
        self._colptr = _long_vector(colptr, "colptr", N + 1, dev)
        self._colcount = _long_vector(colcount, "colcount", N, dev)
        self._csr2csc = _long_vector(csr2csc, "csr2csc", E, dev)
        self._csc2csr = _long_vector(csc2csr, "csc2csr", E, dev)
        self._row_csc = None  # row[csr2csc]: derived view for the SpMM backward, not one of the reference's cache keys

        if not is_sorted and E > 1:
            self._sort_()

    # ------------------------------------------------------------------ construction helpers
    def _sort_(self) -> None:
        """Bring the entries into row-major order (torch_sparse/storage.py:149-162)."""
        M, N = self._sparse_sizes
        row = self.row()
        if self._col.is_cuda:
            perm = ops.sort_perm(row, self._col, M, N)
        else:
            perm = host.sort_perm(row, self._col, N)
        if perm is None:
            return
        self._row = row[perm]
        self._rowptr = None
        self._col = self._col[perm]
        if self._value is not None:
            self._value = self._value[perm]
        self._csr2csc = None
        self._csc2csr = None
        self._row_csc = None

    @classmethod
    def empty(cls) -> "SparseStorage":
        z = torch.empty(0, dtype=torch.long)
        return cls(row=z, col=z.clone(), sparse_sizes=(0, 0), is_sorted=True, trust_data=True)

    def _replace(self, **kw) -> "SparseStorage":
        """New storage sharing all tensors except the ones overridden in `kw`."""
        fields = dict(row=self._row, rowptr=self._rowptr, col=self._col, value=self._value,
                      sparse_sizes=self._sparse_sizes, rowcount=self._rowcount, colptr=self._colptr,
                      colcount=self._colcount, csr2csc=self._csr2csc, csc2csr=self._csc2csr)
        fields.update(kw)
        return SparseStorage(is_sorted=True, trust_data=True, **fields)

    # ------------------------------------------------------------------ COO / CSR views
    def has_row(self) -> bool:
        return self._row is not None

    def row(self) -> Tensor:
        if self._row is None:
            if self._rowptr is None:
                raise ValueError
            fn = ops.ptr2ind if self._rowptr.is_cuda else host.ptr2ind
            self._row = fn(self._rowptr, self._col.numel())
        return self._row

    def has_rowptr(self) -> bool:
        return self._rowptr is not None

    def rowptr(self) -> Tensor:
        if self._rowptr is None:
            if self._row is None:
                raise ValueError
            fn = ops.ind2ptr if self._row.is_cuda else host.ind2ptr
            self._rowptr = fn(self._row, self._sparse_sizes[0])
        return self._rowptr

    def col(self) -> Tensor:
        return self._col

    def has_value(self) -> bool:
        return self._value is not None

    def value(self) -> Optional[Tensor]:
        return self._value

    def _checked_value(self, value: Optional[Tensor], layout: Optional[str]) -> Optional[Tensor]:
        if value is None:
            return None
        if layout == "csc":
            value = value[self.csc2csr()]
        assert value.device == self._col.device
        assert value.size(0) == self._col.numel()
        return value.contiguous()

    def set_value_(self, value: Optional[Tensor], layout: Optional[str] = None) -> "SparseStorage":
        self._value = self._checked_value(value, layout)
        return self

    def set_value(self, value: Optional[Tensor], layout: Optional[str] = None) -> "SparseStorage":
        return self._replace(value=self._checked_value(value, layout))

    def sparse_sizes(self) -> Tuple[int, int]:
        return self._sparse_sizes

    def sparse_size(self, dim: int) -> int:
        return self._sparse_sizes[dim]

    def sparse_resize(self, sparse_sizes: Tuple[int, int]) -> "SparseStorage":
        assert len(sparse_sizes) == 2
        (M0, N0), nnz = self._sparse_sizes, self._col.numel()

        def _grow(ptr, count, old, new):
            d = new - old
            if d > 0:
                if ptr is not None:
                    ptr = torch.cat([ptr, ptr.new_full((d,), nnz)])
                if count is not None:
                    count = torch.cat([count, count.new_zeros(d)])
            elif d < 0:
                ptr = None if ptr is None else ptr[:d]
                count = None if count is None else count[:d]
            return ptr, count

        rowptr, rowcount = _grow(self._rowptr, self._rowcount, M0, sparse_sizes[0])
        colptr, colcount = _grow(self._colptr, self._colcount, N0, sparse_sizes[1])
        return self._replace(rowptr=rowptr, rowcount=rowcount, colptr=colptr, colcount=colcount,
                             sparse_sizes=tuple(sparse_sizes))

    def sparse_reshape(self, num_rows: int, num_cols: int) -> "SparseStorage":
        """Same entries, re-indexed for a (num_rows x num_cols) shape with the same element count; -1 infers one
        extent (torch_sparse/storage.py:316-346). Row-major order is preserved by construction."""
        assert num_rows > 0 or num_rows == -1
        assert num_cols > 0 or num_cols == -1
        assert num_rows > 0 or num_cols > 0
        total = self._sparse_sizes[0] * self._sparse_sizes[1]
        if num_rows == -1:
            num_rows = total // num_cols
        if num_cols == -1:
            num_cols = total // num_rows
        assert num_rows * num_cols == total
        lin = self._sparse_sizes[1] * self.row() + self._col
        return SparseStorage(row=torch.div(lin, num_cols, rounding_mode="floor"), col=lin % num_cols,
                             value=self._value, sparse_sizes=(num_rows, num_cols), is_sorted=True, trust_data=True)

    # ------------------------------------------------------------------ derived caches
    def has_rowcount(self) -> bool:
        return self._rowcount is not None

    def rowcount(self) -> Tensor:
        if self._rowcount is None:
            rowptr = self.rowptr()
            self._rowcount = rowptr[1:] - rowptr[:-1]
        return self._rowcount

    def has_colptr(self) -> bool:
        return self._colptr is not None

    def has_colcount(self) -> bool:
        return self._colcount is not None

    def has_csr2csc(self) -> bool:
        return self._csr2csc is not None

    def has_csc2csr(self) -> bool:
        return self._csc2csr is not None

    def _build_csc_(self) -> None:
        """csr2csc and colptr in one pass (torch_sparse/storage.py:369-385, 407-416)."""
        M, N = self._sparse_sizes
        if self._col.is_cuda:
            perm, colptr, row_csc = ops.csr2csc(self.row(), self._col, M, N, want_colptr=True, want_row_csc=True)
            self._csr2csc = perm
            self._row_csc = row_csc
            if self._colptr is None:
                self._colptr = colptr
        else:
            self._csr2csc = host.csr2csc(self.row(), self._col, M)

    def csr2csc(self) -> Tensor:
        if self._csr2csc is None:
            self._build_csc_()
        return self._csr2csc

    def row_csc(self) -> Tensor:
        """row[csr2csc] — the column index array of the CSC view, i.e. of A^T in CSR form. The reference gathers it
        on every backward (csrc/spmm.cpp:104); it only depends on the structure, so it is kept (the csr2csc kernel
        emits it for free)."""
        if self._row_csc is None:
            if self._csr2csc is None and self._col.is_cuda:
                self._build_csc_()
            if self._row_csc is None:
                self._row_csc = self.row()[self.csr2csc()]
        return self._row_csc

    def colptr(self) -> Tensor:
        if self._colptr is None:
            if self._col.is_cuda:
                if self._csr2csc is None:
                    self._build_csc_()
                else:
                    self._colptr = ops.ind2ptr(self._col[self._csr2csc], self._sparse_sizes[1])
            else:
                self._colptr = host.ind2ptr(self._col[self.csr2csc()], self._sparse_sizes[1])
        return self._colptr

    def colcount(self) -> Tensor:
        if self._colcount is None:
            colptr = self.colptr()
            self._colcount = colptr[1:] - colptr[:-1]
        return self._colcount

    def csc2csr(self) -> Tensor:
        if self._csc2csr is None:
            perm = self.csr2csc()
            inv = torch.empty_like(perm)
            inv[perm] = torch.arange(perm.numel(), dtype=perm.dtype, device=perm.device)
            self._csc2csr = inv
        return self._csc2csr

    # ------------------------------------------------------------------ coalesce
    def is_coalesced(self) -> bool:
        return host.is_coalesced(self.row(), self._col, self._sparse_sizes[1])

    def coalesce(self, reduce: str = "add") -> "SparseStorage":
        """Merge duplicate (row, col) entries (torch_sparse/storage.py:436-466)."""
        E = self._col.numel()
        if E < 2:
            return self
        M, N = self._sparse_sizes
        row, col, value = ops.coalesce(self.row(), self._col, self._value, M, N, reduce)
        if row.numel() == E:  # nothing merged; entries were already sorted by construction
            return self
        return SparseStorage(row=row, col=col, value=value, sparse_sizes=self._sparse_sizes, is_sorted=True,
                             trust_data=True)

    # ------------------------------------------------------------------ cache control
    def fill_cache_(self) -> "SparseStorage":
        self.row(); self.rowptr(); self.rowcount(); self.colptr(); self.colcount(); self.csr2csc(); self.csc2csr()
        return self

    def clear_cache_(self) -> "SparseStorage":
        for k in _CACHE_KEYS:
            setattr(self, "_" + k, None)
        self._row_csc = None
        return self

    def cached_keys(self) -> List[str]:
        return [k for k in _CACHE_KEYS if getattr(self, "_" + k) is not None]

    def num_cached_keys(self) -> int:
        return len(self.cached_keys())

    # ------------------------------------------------------------------ copies / moves
    def _map(self, fn, value_fn=None) -> "SparseStorage":
        def m(t):
            return None if t is None else fn(t)

        value = self._value
        if value is not None:
            value = (value_fn or fn)(value)
        return SparseStorage(row=m(self._row), rowptr=m(self._rowptr), col=fn(self._col), value=value,
                             sparse_sizes=self._sparse_sizes, rowcount=m(self._rowcount), colptr=m(self._colptr),
                             colcount=m(self._colcount), csr2csc=m(self._csr2csc), csc2csr=m(self._csc2csr),
                             is_sorted=True, trust_data=True)

    def copy(self) -> "SparseStorage":
        return self._replace()

    def clone(self) -> "SparseStorage":
        return self._map(lambda t: t.clone())

    def type(self, dtype: torch.dtype, non_blocking: bool = False) -> "SparseStorage":
        if self._value is None or self._value.dtype == dtype:
            return self
        return self.set_value(self._value.to(dtype=dtype, non_blocking=non_blocking), layout="coo")

    def type_as(self, tensor: Tensor, non_blocking: bool = False) -> "SparseStorage":
        return self.type(tensor.dtype, non_blocking)

    def to_device(self, device: torch.device, non_blocking: bool = False) -> "SparseStorage":
        device = torch.device(device)
        if device == self._col.device:
            return self
        return self._map(lambda t: t.to(device, non_blocking=non_blocking))

    def device_as(self, tensor: Tensor, non_blocking: bool = False) -> "SparseStorage":
        return self.to_device(tensor.device, non_blocking)

    def cuda(self) -> "SparseStorage":
        return self if self._col.is_cuda else self._map(lambda t: t.cuda())

    def cpu(self) -> "SparseStorage":
        return self._map(lambda t: t.cpu()) if self._col.is_cuda else self

    def is_cuda(self) -> bool:
        return self._col.is_cuda

    def pin_memory(self) -> "SparseStorage":
        return self._map(lambda t: t.pin_memory())

    def is_pinned(self) -> bool:
        tensors = [self._row, self._rowptr, self._col, self._value, self._rowcount, self._colptr, self._colcount,
                   self._csr2csc, self._csc2csr]
        return all(t.is_pinned() for t in tensors if t is not None)
