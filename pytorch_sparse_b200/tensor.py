"""SparseTensor — the handle the hot-path operators are methods of.

API-compatible with the parts of `torch_sparse.tensor.SparseTensor` that lie on the sparse-matmul
path (torch_sparse/tensor.py:16-57 ctor/from_storage, :78-101 from_dense, :233-244 coo/csr/csc,
:282 coalesce, :530-563 to_dense / to_torch_sparse_coo_tensor) plus the small accessors its tests
use, and like the reference's a TorchScript class (torch_sparse/tensor.py:12) over the scripted
SparseStorage, so `torch.jit.script` functions can take, build and return SparseTensors.
`matmul / spmm / spspmm / __matmul__ / t` are bound in matmul.py / transpose.py exactly like
the reference does (torch_sparse/matmul.py:166-171, torch_sparse/transpose.py:34); the Python-only
conveniences (`to`, `cpu`, `cuda`, `__eq__`, `__repr__`, `__getitem__`) are attached below the class.
"""
from typing import Any, List, Optional, Tuple

import torch
from torch import Tensor

from .storage import SparseStorage


@torch.jit.script
class SparseTensor(object):
    storage: SparseStorage

    def __init__(self, row: Optional[Tensor] = None, rowptr: Optional[Tensor] = None,
                 col: Optional[Tensor] = None, value: Optional[Tensor] = None,
                 sparse_sizes: Optional[Tuple[Optional[int], Optional[int]]] = None, is_sorted: bool = False,
                 trust_data: bool = False):
        self.storage = SparseStorage(row, rowptr, col, value, sparse_sizes, None, None, None, None, None, is_sorted,
                                     trust_data)

    # ------------------------------------------------------------------ constructors
    @classmethod
    def from_storage(self, storage: SparseStorage):
        # a handle around an existing storage: built over its (already ordered) arrays, then pointed at the storage
        # itself so that every cache it holds is shared
        out = SparseTensor(storage._row, storage._rowptr, storage._col, storage._value,
                           (storage._sparse_sizes[0], storage._sparse_sizes[1]), True, True)
        out.storage = storage
        return out

    @classmethod
    def from_edge_index(self, edge_index: Tensor, edge_attr: Optional[Tensor] = None,
                        sparse_sizes: Optional[Tuple[Optional[int], Optional[int]]] = None,
                        is_sorted: bool = False, trust_data: bool = False):
        return SparseTensor(edge_index[0], None, edge_index[1], edge_attr, sparse_sizes, is_sorted, trust_data)

    @classmethod
    def from_dense(self, mat: Tensor, has_value: bool = True):
        if mat.dim() > 2:
            index = mat.abs().sum([i for i in range(2, mat.dim())]).nonzero()
        else:
            index = mat.nonzero()
        index = index.t()
        row, col = index[0], index[1]
        value: Optional[Tensor] = None
        if has_value:
            value = mat[row, col]
        return SparseTensor(row, None, col, value, (mat.size(0), mat.size(1)), True, True)

    @classmethod
    def from_torch_sparse_coo_tensor(self, mat: Tensor, has_value: bool = True):
        mat = mat.coalesce()
        index = mat._indices()
        value: Optional[Tensor] = None
        if has_value:
            value = mat._values()
        return SparseTensor(index[0], None, index[1], value, (mat.size(0), mat.size(1)), True, True)

    @classmethod
    def eye(self, M: int, N: Optional[int] = None, has_value: bool = True, dtype: Optional[torch.dtype] = None,
            device: Optional[torch.device] = None):
        n = M if N is None else N
        idx = torch.arange(min(M, n), device=device)
        value: Optional[Tensor] = None
        if has_value:
            value = torch.ones(idx.numel(), dtype=dtype, device=device)
        return SparseTensor(idx, None, idx, value, (M, n), True, True)

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

    def set_value_(self, value: Optional[Tensor], layout: Optional[str] = None):
        self.storage.set_value_(value, layout)
        return self

    def set_value(self, value: Optional[Tensor], layout: Optional[str] = None):
        return self.from_storage(self.storage.set_value(value, layout))

    def fill_value_(self, fill_value: float, dtype: Optional[torch.dtype] = None):
        value = torch.full((self.nnz(),), fill_value, dtype=dtype, device=self.device())
        return self.set_value_(value, layout="coo")

    def fill_value(self, fill_value: float, dtype: Optional[torch.dtype] = None):
        value = torch.full((self.nnz(),), fill_value, dtype=dtype, device=self.device())
        return self.set_value(value, layout="coo")

    # ------------------------------------------------------------------ sizes
    def sparse_sizes(self) -> Tuple[int, int]:
        return self.storage.sparse_sizes()

    def sparse_size(self, dim: int) -> int:
        return self.storage.sparse_sizes()[dim]

    def sparse_resize(self, sparse_sizes: Tuple[int, int]):
        return self.from_storage(self.storage.sparse_resize(sparse_sizes))

    def sparse_reshape(self, num_rows: int, num_cols: int):
        return self.from_storage(self.storage.sparse_reshape(num_rows, num_cols))

    def sizes(self) -> List[int]:
        M, N = self.storage.sparse_sizes()
        sizes = [M, N]
        value = self.storage.value()
        if value is not None:
            sizes += value.size()[1:]
        return sizes

    def size(self, dim: int) -> int:
        return self.sizes()[dim]

    def dim(self) -> int:
        return len(self.sizes())

    def nnz(self) -> int:
        return self.storage.col().numel()

    def numel(self) -> int:
        value = self.storage.value()
        if value is not None:
            return value.numel()
        return self.nnz()

    def density(self) -> float:
        M, N = self.storage.sparse_sizes()
        if M * N == 0:
            return 0.0
        return self.nnz() / (M * N)

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

    # ------------------------------------------------------------------ coalesce / caches
    def is_coalesced(self) -> bool:
        return self.storage.is_coalesced()

    def coalesce(self, reduce: str = "sum"):
        return self.from_storage(self.storage.coalesce(reduce))

    def fill_cache_(self):
        self.storage.fill_cache_()
        return self

    def clear_cache_(self):
        self.storage.clear_cache_()
        return self

    def copy(self):
        return self.from_storage(self.storage)

    def clone(self):
        return self.from_storage(self.storage.clone())

    # ------------------------------------------------------------------ autograd plumbing
    def requires_grad(self) -> bool:
        value = self.storage.value()
        return value is not None and value.requires_grad

    def requires_grad_(self, requires_grad: bool = True, dtype: Optional[torch.dtype] = None):
        if requires_grad and not self.has_value():
            self.storage.set_value_(torch.ones(self.nnz(), dtype=dtype, device=self.device()), layout="coo")
        value = self.storage.value()
        if value is not None:
            value.requires_grad_(requires_grad)
        return self

    def detach_(self):
        value = self.storage.value()
        if value is not None:
            value.detach_()
        return self

    def detach(self):
        value = self.storage.value()
        if value is None:
            return self
        return self.set_value(value.detach(), layout="coo")

    # ------------------------------------------------------------------ dtype / device
    def device(self) -> torch.device:
        return self.storage.col().device

    def is_cuda(self) -> bool:
        return self.storage.col().is_cuda

    def dtype(self) -> torch.dtype:
        value = self.storage.value()
        if value is not None:
            return value.dtype
        return torch.float

    def is_floating_point(self) -> bool:
        value = self.storage.value()
        if value is not None:
            return torch.is_floating_point(value)
        return True

    def type(self, dtype: torch.dtype, non_blocking: bool = False):
        value = self.storage.value()
        if value is None or value.dtype == dtype:
            return self
        return self.from_storage(self.storage.type(dtype, non_blocking))

    def type_as(self, tensor: Tensor, non_blocking: bool = False):
        return self.type(tensor.dtype, non_blocking)

    def to_device(self, device: torch.device, non_blocking: bool = False):
        if device == self.device():
            return self
        return self.from_storage(self.storage.to_device(device, non_blocking))

    def device_as(self, tensor: Tensor, non_blocking: bool = False):
        return self.to_device(tensor.device, non_blocking)

    def pin_memory(self):
        return self.from_storage(self.storage.pin_memory())

    def is_pinned(self) -> bool:
        return self.storage.is_pinned()

    def bfloat16(self):
        return self.type(torch.bfloat16, False)

    def half(self):
        return self.type(torch.half, False)

    def float(self):
        return self.type(torch.float, False)

    def double(self):
        return self.type(torch.double, False)

    def int(self):
        return self.type(torch.int, False)

    def long(self):
        return self.type(torch.long, False)

    # ------------------------------------------------------------------ row slicing (row-block sharding)
    def narrow_rows(self, start: int, length: int):
        """Rows [start, start+length) as a new SparseTensor — the 1-D row-block partitioner
        (same result as torch_sparse.narrow(src, 0, start, length), torch_sparse/narrow.py:15-42)."""
        rowptr, col, value = self.csr()
        sub_ptr = rowptr[start:start + length + 1]
        lo, hi = int(sub_ptr[0]), int(sub_ptr[-1])
        if value is not None:
            value = value[lo:hi]
        return SparseTensor(None, sub_ptr - lo, col[lo:hi], value, (length, self.sparse_size(1)), True, True)

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


# ---------------------------------------------------------------------- Python-only conveniences
def _to(self: SparseTensor, *args: Any, **kwargs: Any) -> SparseTensor:
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


def _cpu(self: SparseTensor) -> SparseTensor:
    return self.to_device(torch.device("cpu"))


def _cuda(self: SparseTensor, device=None, non_blocking: bool = False) -> SparseTensor:
    return self.to_device(torch.device("cuda" if device is None else device), non_blocking)


def _eq(self: SparseTensor, other) -> bool:
    """Same sizes, same CSR structure, same values (torch_sparse/tensor.py:293-313)."""
    if not isinstance(other, SparseTensor):
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


def _repr(self: SparseTensor) -> str:
    row, col, value = self.coo()
    parts = [f"row={row.tolist() if row.numel() <= 16 else '[...]'}",
             f"col={col.tolist() if col.numel() <= 16 else '[...]'}"]
    if value is not None:
        parts.append(f"val={value.tolist() if value.numel() <= 16 else '[...]'}")
    parts.append(f"size={tuple(self.sizes())}, nnz={self.nnz()}, density={100 * self.density():.2f}%")
    return "SparseTensor(" + ", ".join(parts) + ")"


SparseTensor.to = _to
SparseTensor.cpu = _cpu
SparseTensor.cuda = _cuda
SparseTensor.__eq__ = _eq
SparseTensor.__ne__ = lambda self, other: not _eq(self, other)
SparseTensor.__hash__ = object.__hash__
SparseTensor.__repr__ = _repr
