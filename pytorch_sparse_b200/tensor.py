"""SparseTensor — the handle the hot-path operators are methods of.

API-compatible with the parts of `torch_sparse.tensor.SparseTensor` that lie on the sparse-matmul
path (torch_sparse/tensor.py:16-57 ctor/from_storage, :78-101 from_dense, :233-244 coo/csr/csc,
:282 coalesce, :530-563 to_dense / to_torch_sparse_coo_tensor) plus the small accessors its tests
use. `matmul / spmm / spspmm / __matmul__ / t` are bound in matmul.py / transpose.py exactly like
the reference does (torch_sparse/matmul.py:166-171, torch_sparse/transpose.py:34).
"""
from __future__ import annotations

from typing import Any, List, Optional, Tuple

import torch
from torch import Tensor

from .storage import SparseStorage


class SparseTensor:
    storage: SparseStorage

    def __init__(self, row: Optional[Tensor] = None, rowptr: Optional[Tensor] = None,
                 col: Optional[Tensor] = None, value: Optional[Tensor] = None,
                 sparse_sizes: Optional[Tuple[Optional[int], Optional[int]]] = None, is_sorted: bool = False,
                 trust_data: bool = False):
        self.storage = SparseStorage(row=row, rowptr=rowptr, col=col, value=value, sparse_sizes=sparse_sizes,
                                     is_sorted=is_sorted, trust_data=trust_data)

    # ------------------------------------------------------------------ constructors
    @classmethod
    def from_storage(cls, storage: SparseStorage) -> "SparseTensor":
        out = cls.__new__(cls)
        out.storage = storage
        return out

    @classmethod
    def from_edge_index(cls, edge_index: Tensor, edge_attr: Optional[Tensor] = None,
                        sparse_sizes: Optional[Tuple[Optional[int], Optional[int]]] = None,
                        is_sorted: bool = False, trust_data: bool = False) -> "SparseTensor":
        return cls(row=edge_index[0], col=edge_index[1], value=edge_attr, sparse_sizes=sparse_sizes,
                   is_sorted=is_sorted, trust_data=trust_data)

    @classmethod
    def from_dense(cls, mat: Tensor, has_value: bool = True) -> "SparseTensor":
        if mat.dim() > 2:
            index = mat.abs().sum([i for i in range(2, mat.dim())]).nonzero()
        else:
            index = mat.nonzero()
        index = index.t()
        row, col = index[0], index[1]
        value = mat[row, col] if has_value else None
        return cls(row=row, col=col, value=value, sparse_sizes=(mat.size(0), mat.size(1)), is_sorted=True,
                   trust_data=True)

    @classmethod
    def from_torch_sparse_coo_tensor(cls, mat: Tensor, has_value: bool = True) -> "SparseTensor":
        mat = mat.coalesce()
        index = mat._indices()
        value = mat._values() if has_value else None
        return cls(row=index[0], col=index[1], value=value, sparse_sizes=(mat.size(0), mat.size(1)),
                   is_sorted=True, trust_data=True)

    @classmethod
    def eye(cls, M: int, N: Optional[int] = None, has_value: bool = True, dtype: Optional[torch.dtype] = None,
            device: Optional[torch.device] = None) -> "SparseTensor":
        N = M if N is None else N
        idx = torch.arange(min(M, N), device=device)
        value = torch.ones(idx.numel(), dtype=dtype, device=device) if has_value else None
        return cls(row=idx, col=idx, value=value, sparse_sizes=(M, N), is_sorted=True, trust_data=True)

    # ------------------------------------------------------------------ views
    def coo(self) -> Tuple[Tensor, Tensor, Optional[Tensor]]:
        return self.storage.row(), self.storage.col(), self.storage.value()

    def csr(self) -> Tuple[Tensor, Tensor, Optional[Tensor]]:
        return self.storage.rowptr(), self.storage.col(), self.storage.value()

    def csc(self) -> Tuple[Tensor, Tensor, Optional[Tensor]]:
        perm = self.storage.csr2csc()
        value = self.storage.value()
        if value is not None:
            value = value[perm]
        return self.storage.colptr(), self.storage.row()[perm], value

    # ------------------------------------------------------------------ value handling
    def has_value(self) -> bool:
        return self.storage.has_value()

    def set_value_(self, value: Optional[Tensor], layout: Optional[str] = None) -> "SparseTensor":
        self.storage.set_value_(value, layout)
        return self

    def set_value(self, value: Optional[Tensor], layout: Optional[str] = None) -> "SparseTensor":
        return self.from_storage(self.storage.set_value(value, layout))

    def fill_value_(self, fill_value: float, dtype: Optional[torch.dtype] = None) -> "SparseTensor":
        value = torch.full((self.nnz(),), fill_value, dtype=dtype, device=self.device())
        return self.set_value_(value, layout="coo")

    def fill_value(self, fill_value: float, dtype: Optional[torch.dtype] = None) -> "SparseTensor":
        value = torch.full((self.nnz(),), fill_value, dtype=dtype, device=self.device())
        return self.set_value(value, layout="coo")

    # ------------------------------------------------------------------ sizes
    def sparse_sizes(self) -> Tuple[int, int]:
        return self.storage.sparse_sizes()

    def sparse_size(self, dim: int) -> int:
        return self.storage.sparse_sizes()[dim]

    def sparse_resize(self, sparse_sizes: Tuple[int, int]) -> "SparseTensor":
        return self.from_storage(self.storage.sparse_resize(sparse_sizes))

    def sparse_reshape(self, num_rows: int, num_cols: int) -> "SparseTensor":
        return self.from_storage(self.storage.sparse_reshape(num_rows, num_cols))

    def sizes(self) -> List[int]:
        sizes = list(self.sparse_sizes())
        value = self.storage.value()
        if value is not None:
            sizes += list(value.shape[1:])
        return sizes

    def size(self, dim: int) -> int:
        return self.sizes()[dim]

    def dim(self) -> int:
        return len(self.sizes())

    def nnz(self) -> int:
        return self.storage.col().numel()

    def numel(self) -> int:
        value = self.storage.value()
        return value.numel() if value is not None else self.nnz()

    def density(self) -> float:
        M, N = self.sparse_sizes()
        return self.nnz() / (M * N) if M * N > 0 else 0.0

    def sparsity(self) -> float:
        return 1 - self.density()

    def avg_row_length(self) -> float:
        return self.nnz() / self.sparse_size(0)

    def avg_col_length(self) -> float:
        return self.nnz() / self.sparse_size(1)

    def is_quadratic(self) -> bool:
        return self.sparse_size(0) == self.sparse_size(1)

    def is_symmetric(self) -> bool:
        """CSR view == CSC view, structure and values (torch_sparse/tensor.py:389-402)."""
        if not self.is_quadratic():
            return False
        rowptr, col, value1 = self.csr()
        colptr, row, value2 = self.csc()
        if bool((rowptr != colptr).any()) or bool((col != row).any()):
            return False
        if value1 is None or value2 is None:
            return True
        return bool((value1 == value2).all())

    def __eq__(self, other) -> bool:
        """Same sizes, same CSR structure, same values (torch_sparse/tensor.py:293-313)."""
        if not isinstance(other, self.__class__):
            return False
        if self.sizes() != other.sizes():
            return False
        rowptrA, colA, valueA = self.csr()
        rowptrB, colB, valueB = other.csr()
        if (valueA is None) != (valueB is None):
            return False
        if not torch.equal(rowptrA, rowptrB) or not torch.equal(colA, colB):
            return False
        return True if valueA is None else torch.equal(valueA, valueB)

    __hash__ = object.__hash__

    # ------------------------------------------------------------------ coalesce / caches
    def is_coalesced(self) -> bool:
        return self.storage.is_coalesced()

    def coalesce(self, reduce: str = "sum") -> "SparseTensor":
        return self.from_storage(self.storage.coalesce(reduce))

    def fill_cache_(self) -> "SparseTensor":
        self.storage.fill_cache_()
        return self

    def clear_cache_(self) -> "SparseTensor":
        self.storage.clear_cache_()
        return self

    def copy(self) -> "SparseTensor":
        return self.from_storage(self.storage)

    def clone(self) -> "SparseTensor":
        return self.from_storage(self.storage.clone())

    # ------------------------------------------------------------------ autograd plumbing
    def requires_grad(self) -> bool:
        value = self.storage.value()
        return value is not None and value.requires_grad

    def requires_grad_(self, requires_grad: bool = True, dtype: Optional[torch.dtype] = None) -> "SparseTensor":
        if requires_grad and not self.has_value():
            self.storage.set_value_(torch.ones(self.nnz(), dtype=dtype, device=self.device()), layout="coo")
        value = self.storage.value()
        if value is not None:
            value.requires_grad_(requires_grad)
        return self

    def detach_(self) -> "SparseTensor":
        value = self.storage.value()
        if value is not None:
            value.detach_()
        return self

    def detach(self) -> "SparseTensor":
        value = self.storage.value()
        return self if value is None else self.set_value(value.detach(), layout="coo")

    # ------------------------------------------------------------------ dtype / device
    def device(self) -> torch.device:
        return self.storage.col().device

    def is_cuda(self) -> bool:
        return self.storage.col().is_cuda

    def dtype(self) -> torch.dtype:
        value = self.storage.value()
        return value.dtype if value is not None else torch.float

    def is_floating_point(self) -> bool:
        value = self.storage.value()
        return torch.is_floating_point(value) if value is not None else True

    def type(self, dtype: torch.dtype, non_blocking: bool = False) -> "SparseTensor":
        storage = self.storage.type(dtype, non_blocking)
        return self if storage is self.storage else self.from_storage(storage)

    def type_as(self, tensor: Tensor, non_blocking: bool = False) -> "SparseTensor":
        return self.type(tensor.dtype, non_blocking)

    def to_device(self, device, non_blocking: bool = False) -> "SparseTensor":
        storage = self.storage.to_device(device, non_blocking)
        return self if storage is self.storage else self.from_storage(storage)

    def device_as(self, tensor: Tensor, non_blocking: bool = False) -> "SparseTensor":
        return self.to_device(tensor.device, non_blocking)

    def to(self, *args: Any, **kwargs: Any) -> "SparseTensor":
        out = self
        non_blocking = bool(kwargs.get("non_blocking", False))
        for a in list(args) + [kwargs.get("dtype"), kwargs.get("device")]:
            if a is None:
                continue
            if isinstance(a, torch.dtype):
                out = out.type(a, non_blocking)
            elif isinstance(a, (torch.device, str, int)):
                out = out.to_device(torch.device(a), non_blocking)
            elif isinstance(a, Tensor):
                out = out.type(a.dtype, non_blocking).to_device(a.device, non_blocking)
        return out

    def cpu(self) -> "SparseTensor":
        return self.to_device(torch.device("cpu"))

    def cuda(self, device=None, non_blocking: bool = False) -> "SparseTensor":
        return self.to_device(torch.device("cuda" if device is None else device), non_blocking)

    def pin_memory(self) -> "SparseTensor":
        return self.from_storage(self.storage.pin_memory())

    def is_pinned(self) -> bool:
        return self.storage.is_pinned()

    def bfloat16(self): return self.type(torch.bfloat16)
    def half(self): return self.type(torch.half)
    def float(self): return self.type(torch.float)
    def double(self): return self.type(torch.double)
    def int(self): return self.type(torch.int)
    def long(self): return self.type(torch.long)

    # ------------------------------------------------------------------ row slicing (row-block sharding)
    def narrow_rows(self, start: int, length: int) -> "SparseTensor":
        """Rows [start, start+length) as a new SparseTensor — the 1-D row-block partitioner
        (same result as torch_sparse.narrow(src, 0, start, length), torch_sparse/narrow.py:15-42)."""
        rowptr, col, value = self.csr()
        sub_ptr = rowptr[start:start + length + 1]
        lo, hi = int(sub_ptr[0]), int(sub_ptr[-1])
        return SparseTensor(rowptr=sub_ptr - lo, col=col[lo:hi], value=None if value is None else value[lo:hi],
                            sparse_sizes=(length, self.sparse_size(1)), is_sorted=True, trust_data=True)

    # ------------------------------------------------------------------ conversions
    def to_dense(self, dtype: Optional[torch.dtype] = None) -> Tensor:
        row, col, value = self.coo()
        if value is not None:
            mat = torch.zeros(self.sizes(), dtype=value.dtype, device=self.device())
            mat[row, col] = value
        else:
            mat = torch.zeros(self.sizes(), dtype=dtype, device=self.device())
            mat[row, col] = torch.ones(self.nnz(), dtype=mat.dtype, device=mat.device)
        return mat

    def to_torch_sparse_coo_tensor(self, dtype: Optional[torch.dtype] = None) -> Tensor:
        row, col, value = self.coo()
        index = torch.stack([row, col], dim=0)
        if value is None:
            value = torch.ones(self.nnz(), dtype=dtype, device=self.device())
        return torch.sparse_coo_tensor(index, value, self.sizes())

    def __repr__(self) -> str:
        row, col, value = self.coo()
        parts = [f"row={row.tolist() if row.numel() <= 16 else '[...]'}",
                 f"col={col.tolist() if col.numel() <= 16 else '[...]'}"]
        if value is not None:
            parts.append(f"val={value.tolist() if value.numel() <= 16 else '[...]'}")
        parts.append(f"size={tuple(self.sizes())}, nnz={self.nnz()}, density={100 * self.density():.2f}%")
        return "SparseTensor(" + ", ".join(parts) + ")"
