"""SparseStorage — canonical row-major COO plus lazily derived CSR / CSC views.

API-compatible with `torch_sparse.storage.SparseStorage` (torch_sparse/storage.py:20-801): same
constructor arguments, the same private attribute names (`_row`, `_rowptr`, `_col`, `_value`,
`_rowcount`, `_colptr`, `_colcount`, `_csr2csc`, `_csc2csr`) that `matmul.py` and the reference's
tests read, the same lazy-cache semantics, and — like the reference's — a TorchScript class
(`@torch.jit.script`, torch_sparse/storage.py:21): every native step is a registered operator
(`torch.ops.tsb200.*`, torch_ops.py), so scripted code can build and use storages. Differences, all on
the GPU side:
  * sort-on-construct, csr2csc/colptr and coalesce run as fused libtsb200 kernels
    (one stable radix sort over only the significant key bits, no E-sized int64 temporaries),
  * `trust_data=True` (or CUDA tensors + explicit sizes) skips the host-synchronising bound checks,
  * `_row_csc` (= row[csr2csc], the column array of the transposed CSR view) is kept as an extra cache.
CPU tensors are accepted for CONSTRUCTION only (index bookkeeping in _host_index.py); all arithmetic is CUDA.
"""
from typing import List, Optional, Tuple

import torch
from torch import Tensor

from . import _host_index as host
from . import torch_ops  # noqa: F401  (the operators below must be registered before this class is compiled)


def _vetted(t: Optional[Tensor], what: str, numel: int, like: Tensor) -> Optional[Tensor]:
    """An optional int64 index vector of `numel` entries (-1: any length) on the device of `like`."""
    if t is None:
        return None
    assert t.dtype == torch.long, what + " must be int64"
    assert t.dim() == 1, what + " must be 1-dimensional"
    assert t.device == like.device, what + " lives on another device than col"
    if numel >= 0:
        assert t.numel() == numel, what + " has the wrong number of entries"
    return t.contiguous()


def _moved(t: Optional[Tensor], how: int, device: torch.device, non_blocking: bool) -> Optional[Tensor]:
    """how: 0 = clone, 1 = move to `device`, 2 = pin."""
    if t is None:
        return None
    if how == 0:
        return t.clone()
    if how == 1:
        return t.to(device, non_blocking=non_blocking)
    return t.pin_memory()


def _resized(ptr: Optional[Tensor], count: Optional[Tensor], old: int, new: int,
             nnz: int) -> Tuple[Optional[Tensor], Optional[Tensor]]:
    """Pointer / count vectors of one sparse dimension after growing or shrinking it from `old` to `new`."""
    d = new - old
    if d > 0:
        if ptr is not None:
            ptr = torch.cat([ptr, ptr.new_full((d,), nnz)])
        if count is not None:
            count = torch.cat([count, count.new_zeros(d)])
    elif d < 0:
        if ptr is not None:
            ptr = ptr[:d]
        if count is not None:
            count = count[:d]
    return ptr, count


@torch.jit.script
class SparseStorage(object):
    _row: Optional[Tensor]
    _rowptr: Optional[Tensor]
    _col: Tensor
    _value: Optional[Tensor]
    _sparse_sizes: Tuple[int, int]
    _rowcount: Optional[Tensor]
    _colptr: Optional[Tensor]
    _colcount: Optional[Tensor]
    _csr2csc: Optional[Tensor]
    _csc2csr: Optional[Tensor]
    _row_csc: Optional[Tensor]

    def __init__(self, row: Optional[Tensor] = None, rowptr: Optional[Tensor] = None,
                 col: Optional[Tensor] = None, value: Optional[Tensor] = None,
                 sparse_sizes: Optional[Tuple[Optional[int], Optional[int]]] = None,
                 rowcount: Optional[Tensor] = None, colptr: Optional[Tensor] = None,
                 colcount: Optional[Tensor] = None, csr2csc: Optional[Tensor] = None,
                 csc2csr: Optional[Tensor] = None, is_sorted: bool = False, trust_data: bool = False):
        assert row is not None or rowptr is not None
        assert col is not None
        col_v = _vetted(col, "col", -1, col)
        assert col_v is not None
        E = col_v.numel()

        m_hint: Optional[int] = None
        n_hint: Optional[int] = None
        if sparse_sizes is not None:
            m_hint = sparse_sizes[0]
            n_hint = sparse_sizes[1]
        M = 0
        if m_hint is None:
            if rowptr is not None:
                M = rowptr.numel() - 1
            elif row is not None and E > 0:
                M = int(row.max()) + 1
        else:
            M = m_hint
            if rowptr is not None:
                assert rowptr.numel() - 1 == M
            elif row is not None and E > 0 and not trust_data:
                assert int(row.max()) < M
        N = 0
        if n_hint is None:
            if E > 0:
                N = int(col_v.max()) + 1
        else:
            N = n_hint
            if E > 0 and not trust_data:
                assert int(col_v.max()) < N

        if value is not None:
            assert value.device == col_v.device
            assert value.size(0) == E
            value = value.contiguous()
        self._row = _vetted(row, "row", E, col_v)
        self._rowptr = _vetted(rowptr, "rowptr", M + 1, col_v)
        self._col = col_v
        self._value = value
        self._sparse_sizes = (M, N)
        self._rowcount = _vetted(rowcount, "rowcount", M, col_v)
        self._colptr = _vetted(colptr, "colptr", N + 1, col_v)
        self._colcount = _vetted(colcount, "colcount", N, col_v)
        self._csr2csc = _vetted(csr2csc, "csr2csc", E, col_v)
        self._csc2csr = _vetted(csc2csr, "csc2csr", E, col_v)
        row_csc: Optional[Tensor] = None  # row[csr2csc]: derived view for the SpMM backward, not a reference cache key
        self._row_csc = row_csc

        if not is_sorted and E > 1:
            self._sort_()

    # ------------------------------------------------------------------ construction helpers
    def _sort_(self):
        """Bring the entries into row-major order (torch_sparse/storage.py:149-162)."""
        M, N = self._sparse_sizes
        row = self.row()
        perm: Optional[Tensor] = None
        if self._col.is_cuda:
            perm = torch.ops.tsb200.sort_perm(row, self._col, M, N)
        else:
            perm = host.sort_perm(row, self._col, N)
        if perm is not None:
            self._row = row[perm]
            self._rowptr = None
            self._col = self._col[perm]
            value = self._value
            if value is not None:
                self._value = value[perm]
            self._csr2csc = None
            self._csc2csr = None
            self._row_csc = None

    @classmethod
    def empty(self):
        z = torch.empty(0, dtype=torch.long)
        return SparseStorage(z, None, z.clone(), None, (0, 0), None, None, None, None, None, True, True)

    def _with(self, row: Optional[Tensor], rowptr: Optional[Tensor], col: Tensor, value: Optional[Tensor],
              sparse_sizes: Tuple[int, int], rowcount: Optional[Tensor], colptr: Optional[Tensor],
              colcount: Optional[Tensor], csr2csc: Optional[Tensor], csc2csr: Optional[Tensor]):
        """A storage over already ordered, already validated pieces."""
        return SparseStorage(row, rowptr, col, value, (sparse_sizes[0], sparse_sizes[1]), rowcount, colptr, colcount,
                             csr2csc, csc2csr, True, True)

    # ------------------------------------------------------------------ COO / CSR views
    def has_row(self) -> bool:
        return self._row is not None

    def row(self) -> Tensor:
        row = self._row
        if row is None:
            rowptr = self._rowptr
            if rowptr is None:
                raise ValueError
            if rowptr.is_cuda:
                row = torch.ops.tsb200.ptr2ind(rowptr, self._col.numel())
            else:
                row = host.ptr2ind(rowptr, self._col.numel())
            self._row = row
        return row

    def has_rowptr(self) -> bool:
        return self._rowptr is not None

    def rowptr(self) -> Tensor:
        rowptr = self._rowptr
        if rowptr is None:
            row = self._row
            if row is None:
                raise ValueError
            if row.is_cuda:
                rowptr = torch.ops.tsb200.ind2ptr(row, self._sparse_sizes[0])
            else:
                rowptr = host.ind2ptr(row, self._sparse_sizes[0])
            self._rowptr = rowptr
        return rowptr

    def col(self) -> Tensor:
        return self._col

    def has_value(self) -> bool:
        return self._value is not None

    def value(self) -> Optional[Tensor]:
        return self._value

    def _checked_value(self, value: Optional[Tensor], layout: Optional[str]) -> Optional[Tensor]:
        if value is None:
            return None
        if layout is not None and layout == "csc":
            value = value[self.csc2csr()]
        assert value.device == self._col.device
        assert value.size(0) == self._col.numel()
        return value.contiguous()

    def set_value_(self, value: Optional[Tensor], layout: Optional[str] = None):
        self._value = self._checked_value(value, layout)
        return self

    def set_value(self, value: Optional[Tensor], layout: Optional[str] = None):
        out = self._with(self._row, self._rowptr, self._col, self._checked_value(value, layout), self._sparse_sizes,
                         self._rowcount, self._colptr, self._colcount, self._csr2csc, self._csc2csr)
        out._row_csc = self._row_csc
        return out

    def sparse_sizes(self) -> Tuple[int, int]:
        return self._sparse_sizes

    def sparse_size(self, dim: int) -> int:
        return self._sparse_sizes[dim]

    def sparse_resize(self, sparse_sizes: Tuple[int, int]):
        assert len(sparse_sizes) == 2
        nnz = self._col.numel()
        rowptr, rowcount = _resized(self._rowptr, self._rowcount, self._sparse_sizes[0], sparse_sizes[0], nnz)
        colptr, colcount = _resized(self._colptr, self._colcount, self._sparse_sizes[1], sparse_sizes[1], nnz)
        return self._with(self._row, rowptr, self._col, self._value, sparse_sizes, rowcount, colptr, colcount,
                          self._csr2csc, self._csc2csr)

    def sparse_reshape(self, num_rows: int, num_cols: int):
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
        return self._with(torch.div(lin, num_cols, rounding_mode="floor"), None, lin % num_cols, self._value,
                          (num_rows, num_cols), None, None, None, None, None)

    # ------------------------------------------------------------------ derived caches
    def has_rowcount(self) -> bool:
        return self._rowcount is not None

    def rowcount(self) -> Tensor:
        rowcount = self._rowcount
        if rowcount is None:
            rowptr = self.rowptr()
            rowcount = rowptr[1:] - rowptr[:-1]
            self._rowcount = rowcount
        return rowcount

    def has_colptr(self) -> bool:
        return self._colptr is not None

    def has_colcount(self) -> bool:
        return self._colcount is not None

    def has_csr2csc(self) -> bool:
        return self._csr2csc is not None

    def has_csc2csr(self) -> bool:
        return self._csc2csr is not None

    def _build_csc_(self):
        """csr2csc and colptr in one pass (torch_sparse/storage.py:369-385, 407-416)."""
        M, N = self._sparse_sizes
        if self._col.is_cuda:
            perm, colptr, row_csc = torch.ops.tsb200.csr2csc(self.row(), self._col, M, N)
            self._csr2csc = perm
            self._row_csc = row_csc
            if self._colptr is None:
                self._colptr = colptr
        else:
            self._csr2csc = host.csr2csc(self.row(), self._col, M)

    def csr2csc(self) -> Tensor:
        if self._csr2csc is None:
            self._build_csc_()
        perm = self._csr2csc
        assert perm is not None
        return perm

    def row_csc(self) -> Tensor:
        """row[csr2csc] — the column index array of the CSC view, i.e. of A^T in CSR form. The reference gathers it
        on every backward (csrc/spmm.cpp:104); it only depends on the structure, so it is kept (the csr2csc kernel
        emits it for free)."""
        if self._row_csc is None and self._csr2csc is None and self._col.is_cuda:
            self._build_csc_()
        row_csc = self._row_csc
        if row_csc is None:
            row_csc = self.row()[self.csr2csc()]
            self._row_csc = row_csc
        return row_csc

    def colptr(self) -> Tensor:
        if self._colptr is None:
            if self._col.is_cuda:
                if self._csr2csc is None:
                    self._build_csc_()
                else:
                    self._colptr = torch.ops.tsb200.ind2ptr(self._col[self.csr2csc()], self._sparse_sizes[1])
            else:
                self._colptr = host.ind2ptr(self._col[self.csr2csc()], self._sparse_sizes[1])
        colptr = self._colptr
        assert colptr is not None
        return colptr

    def colcount(self) -> Tensor:
        colcount = self._colcount
        if colcount is None:
            colptr = self.colptr()
            colcount = colptr[1:] - colptr[:-1]
            self._colcount = colcount
        return colcount

    def csc2csr(self) -> Tensor:
        inv = self._csc2csr
        if inv is None:
            perm = self.csr2csc()
            inv = torch.empty_like(perm)
            inv[perm] = torch.arange(perm.numel(), dtype=perm.dtype, device=perm.device)
            self._csc2csr = inv
        return inv

    # ------------------------------------------------------------------ coalesce
    def is_coalesced(self) -> bool:
        return host.is_coalesced(self.row(), self._col, self._sparse_sizes[1])

    def coalesce(self, reduce: str = "add"):
        """Merge duplicate (row, col) entries (torch_sparse/storage.py:436-466)."""
        E = self._col.numel()
        if E < 2:
            return self
        M, N = self._sparse_sizes
        row, col, value = torch.ops.tsb200.coalesce(self.row(), self._col, self._value, M, N, reduce)
        if row.numel() == E:  # nothing merged; entries were already sorted by construction
            return self
        return self._with(row, None, col, value, self._sparse_sizes, None, None, None, None, None)

    # ------------------------------------------------------------------ cache control
    def fill_cache_(self):
        self.row()
        self.rowptr()
        self.rowcount()
        self.colptr()
        self.colcount()
        self.csr2csc()
        self.csc2csr()
        return self

    def clear_cache_(self):
        self._rowcount = None
        self._colptr = None
        self._colcount = None
        self._csr2csc = None
        self._csc2csr = None
        self._row_csc = None
        return self

    def cached_keys(self) -> List[str]:
        keys: List[str] = []
        if self._rowcount is not None:
            keys.append("rowcount")
        if self._colptr is not None:
            keys.append("colptr")
        if self._colcount is not None:
            keys.append("colcount")
        if self._csr2csc is not None:
            keys.append("csr2csc")
        if self._csc2csr is not None:
            keys.append("csc2csr")
        return keys

    def num_cached_keys(self) -> int:
        return len(self.cached_keys())

    # ------------------------------------------------------------------ copies / moves
    def _map(self, how: int, device: torch.device, non_blocking: bool):
        col = _moved(self._col, how, device, non_blocking)
        assert col is not None
        out = self._with(_moved(self._row, how, device, non_blocking), _moved(self._rowptr, how, device, non_blocking),
                         col, _moved(self._value, how, device, non_blocking), self._sparse_sizes,
                         _moved(self._rowcount, how, device, non_blocking),
                         _moved(self._colptr, how, device, non_blocking),
                         _moved(self._colcount, how, device, non_blocking),
                         _moved(self._csr2csc, how, device, non_blocking),
                         _moved(self._csc2csr, how, device, non_blocking))
        out._row_csc = _moved(self._row_csc, how, device, non_blocking)
        return out

    def copy(self):
        out = self._with(self._row, self._rowptr, self._col, self._value, self._sparse_sizes, self._rowcount,
                         self._colptr, self._colcount, self._csr2csc, self._csc2csr)
        out._row_csc = self._row_csc
        return out

    def clone(self):
        return self._map(0, self._col.device, False)

    def type(self, dtype: torch.dtype, non_blocking: bool = False):
        value = self._value
        if value is None or value.dtype == dtype:
            return self
        return self.set_value(value.to(dtype=dtype, non_blocking=non_blocking), layout="coo")

    def type_as(self, tensor: Tensor, non_blocking: bool = False):
        return self.type(tensor.dtype, non_blocking)

    def to_device(self, device: torch.device, non_blocking: bool = False):
        if device == self._col.device:
            return self
        return self._map(1, device, non_blocking)

    def device_as(self, tensor: Tensor, non_blocking: bool = False):
        return self.to_device(tensor.device, non_blocking)

    def cuda(self):
        if self._col.is_cuda:
            return self
        return self._map(1, torch.device("cuda"), False)

    def cpu(self):
        if not self._col.is_cuda:
            return self
        return self._map(1, torch.device("cpu"), False)

    def is_cuda(self) -> bool:
        return self._col.is_cuda

    def pin_memory(self):
        return self._map(2, self._col.device, False)

    def is_pinned(self) -> bool:
        ok = self._col.is_pinned()
        for t in [self._row, self._rowptr, self._value, self._rowcount, self._colptr, self._colcount, self._csr2csc,
                  self._csc2csr]:
            if t is not None:
                ok = ok and t.is_pinned()
        return ok
