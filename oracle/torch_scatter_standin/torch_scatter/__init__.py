"""Pure-PyTorch stand-in for the four `torch_scatter` functions the reference's Python layer imports
(torch_sparse/storage.py:5, tensor.py:7, spmm.py:2 ...). torch_scatter is a third-party dependency
that is neither vendored under /root/reference nor installed/pinned in this image; this file
restates its published semantics (sum/mean/min/max over an index or a CSR pointer along one
dimension, empty segments -> 0) so that the UNMODIFIED reference package can be imported to
generate golden vectors (oracle/gen_golden.py). TorchScript-compatible because
`SparseStorage` is a @torch.jit.script class. TEST INFRASTRUCTURE ONLY.
"""
from typing import Optional

import torch
from torch import Tensor

__version__ = "2.1.2"


def _expand(index: Tensor, src: Tensor, dim: int) -> Tensor:
    if dim < 0:
        dim = src.dim() + dim
    if index.dim() == 1:
        for _ in range(0, dim):
            index = index.unsqueeze(0)
    for _ in range(index.dim(), src.dim()):
        index = index.unsqueeze(-1)
    return index.expand(src.size())


def scatter(src: Tensor, index: Tensor, dim: int = -1, out: Optional[Tensor] = None,
            dim_size: Optional[int] = None, reduce: str = "sum") -> Tensor:
    if dim < 0:
        dim = src.dim() + dim
    idx = _expand(index, src, dim)
    size = list(src.size())
    if dim_size is not None:
        size[dim] = dim_size
    elif idx.numel() == 0:
        size[dim] = 0
    else:
        size[dim] = int(idx.max()) + 1
    res = torch.zeros(size, dtype=src.dtype, device=src.device)
    if reduce == "sum" or reduce == "add":
        return res.scatter_add_(dim, idx, src)
    if reduce == "mean":
        res = res.scatter_add_(dim, idx, src)
        cnt = torch.zeros(size, dtype=src.dtype, device=src.device).scatter_add_(dim, idx, torch.ones_like(src))
        cnt = cnt.clamp(min=1)
        if src.is_floating_point():
            return res / cnt
        return torch.div(res, cnt, rounding_mode="floor")
    # min/max come back as fresh tensors: the reference's test edits them in place before calling backward
    # (test/test_matmul.py:29-32), which torch_scatter's own autograd Function allows
    if reduce == "min":
        return res.scatter_reduce_(dim, idx, src, "amin", include_self=False).clone()
    if reduce == "max":
        return res.scatter_reduce_(dim, idx, src, "amax", include_self=False).clone()
    raise ValueError


def scatter_add(src: Tensor, index: Tensor, dim: int = -1, out: Optional[Tensor] = None,
                dim_size: Optional[int] = None) -> Tensor:
    return scatter(src, index, dim, out, dim_size, "sum")


def segment_csr(src: Tensor, indptr: Tensor, out: Optional[Tensor] = None, reduce: str = "sum") -> Tensor:
    n = indptr.numel() - 1
    counts = indptr[1:] - indptr[:-1]
    index = torch.repeat_interleave(torch.arange(n, device=src.device), counts)
    return scatter(src, index, 0, None, n, reduce)


def gather_csr(src: Tensor, indptr: Tensor, out: Optional[Tensor] = None) -> Tensor:
    counts = indptr[1:] - indptr[:-1]
    return torch.repeat_interleave(src, counts, dim=0)
