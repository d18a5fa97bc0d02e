"""Row-sharded SpMM across GPUs (one process per GPU, torch.distributed / NCCL over NVLink).

SpMM output rows are independent, so the path shards naturally by 1-D row blocks of A
(SURVEY §8e): rank r owns rows [r*M/P, (r+1)*M/P) of A (a `narrow_rows` slice) and the matching
row block of the dense operand X. X must be visible to every rank, so the blocks are all-gathered
once (NCCL all_gather over NVLink/NVSwitch); the product stays row-sharded — no reduction.
Backward: grad_value is local; grad_X = A^T grad_out is a full-height partial per rank, so it is
reduce-scattered back to row blocks.

There is no collective inside the SpMM step itself when X is already resident
(`local_spmm`), which is what the benchmark times ("broadcast once", BASELINE north_star).
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist
from torch import Tensor

from .matmul import matmul
from .tensor import SparseTensor


def _world(group) -> int:
    return dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1


class _AllGatherRows(torch.autograd.Function):
    """forward: concatenate every rank's row block (all_gather); backward: reduce_scatter."""

    @staticmethod
    def forward(ctx, x_local: Tensor, group) -> Tensor:
        ctx.group = group
        world = _world(group)
        ctx.rows = x_local.size(0)
        if world == 1:
            return x_local
        x_local = x_local.contiguous()
        out = x_local.new_empty((world * x_local.size(0),) + tuple(x_local.shape[1:]))
        dist.all_gather_into_tensor(out, x_local, group=group)
        return out

    @staticmethod
    def backward(ctx, grad_full: Tensor):
        world = _world(ctx.group)
        if world == 1:
            return grad_full, None
        grad_full = grad_full.contiguous()
        out = grad_full.new_empty((ctx.rows,) + tuple(grad_full.shape[1:]))
        dist.reduce_scatter_tensor(out, grad_full, op=dist.ReduceOp.SUM, group=ctx.group)
        return out, None


class RowShardedSpMM:
    """`a_local` is this rank's row block of A with FULL column extent (cols index the gathered X)."""

    def __init__(self, a_local: SparseTensor, reduce: str = "sum", group=None):
        self.a = a_local
        self.reduce = reduce
        self.group = group

    @staticmethod
    def partition(a: SparseTensor, rank: int, world: int) -> SparseTensor:
        """Contiguous row block `rank` of `world` (equal row counts, the last block takes the rest)."""
        M = a.sparse_size(0)
        per = (M + world - 1) // world
        start = min(rank * per, M)
        return a.narrow_rows(start, min(per, M - start))

    def gather_dense(self, x_local: Tensor) -> Tensor:
        return _AllGatherRows.apply(x_local, self.group)

    def local_spmm(self, x_full: Tensor) -> Tensor:
        return matmul(self.a, x_full, self.reduce)

    def __call__(self, x_local: Tensor) -> Tensor:
        return self.local_spmm(self.gather_dense(x_local))
