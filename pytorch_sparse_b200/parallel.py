"""Row-sharded SpMM across GPUs (one process per GPU, torch.distributed / NCCL over NVLink).

SpMM output rows are independent, so the path shards naturally by 1-D row blocks of A
(SURVEY §8e): rank r owns rows [r*M/P, (r+1)*M/P) of A (a `narrow_rows` slice, the reference's
torch_sparse/narrow.py:15-42) and the matching row block of the dense operand X. X must be visible
to every rank; the product stays row-sharded — no reduction. Backward: grad_value is local;
grad_X = A^T grad_out is a full-height partial per rank, so it is reduce-scattered back to row blocks.

Two ways to make X visible:

  * `RowShardedSpMM`: all-gather X once (NCCL), then any number of steps run `local_spmm` with no
    collective inside the step ("broadcast once", BASELINE north_star) — the steady state.
  * `PipelinedRowShardedSpMM`: X changes every step (chained layers), so the gather is part of the
    step. The gather is cut into C chunks (rows [c*Mb/C, (c+1)*Mb/C) of EVERY rank's block per chunk,
    so every NVLink stays busy in every chunk) issued back to back on NCCL's stream; A's columns are
    split the same way once at set-up, and the SpMM of column chunk c (tsb200_spmm_fw_acc: fp32
    partial shared by the chunks) is launched as soon as chunk c has landed — the multiply of chunk c
    overlaps the transfer of chunks c+1.. .
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
import torch.distributed as dist
from torch import Tensor

from . import ops
from .matmul import matmul
from .tensor import SparseTensor


def _world(group) -> int:
    return dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1


def _rank(group) -> int:
    return dist.get_rank(group) if dist.is_available() and dist.is_initialized() else 0


class _AllGatherRows(torch.autograd.Function):
    """forward: concatenate every rank's row block (all_gather); backward: reduce_scatter.
    Every rank must hold the same number of rows (pad the last block: `RowShardedSpMM.pad_rows`)."""

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
    """`a_local` is this rank's row block of A with FULL column extent (cols index the gathered X).

    The dense operand is sharded by the SAME block height on every rank: `block_rows(N, world)` =
    ceil(N / world); a rank whose share is shorter pads it with zero rows (`pad_rows`), so that the
    gathered operand has world * block_rows >= N rows, the first N of which are X."""

    def __init__(self, a_local: SparseTensor, reduce: str = "sum", group=None):
        self.a = a_local
        self.reduce = reduce
        self.group = group

    @staticmethod
    def block_rows(n: int, world: int) -> int:
        return (n + world - 1) // world

    @staticmethod
    def partition(a: SparseTensor, rank: int, world: int) -> SparseTensor:
        """Contiguous row block `rank` of `world` (ceil(M / world) rows, the last blocks may be shorter or empty)."""
        M = a.sparse_size(0)
        per = RowShardedSpMM.block_rows(M, world)
        start = min(rank * per, M)
        return a.narrow_rows(start, min(per, M - start))

    @staticmethod
    def pad_rows(x_local: Tensor, rows: int) -> Tensor:
        """Zero-pad a dense row block to `rows` rows (no copy when it already has them)."""
        if x_local.size(0) == rows:
            return x_local
        assert x_local.size(0) < rows
        pad = x_local.new_zeros((rows - x_local.size(0),) + tuple(x_local.shape[1:]))
        return torch.cat([x_local, pad], dim=0)

    def gather_dense(self, x_local: Tensor) -> Tensor:
        world = _world(self.group)
        if world > 1:
            n = self.a.sparse_size(1)
            per = self.block_rows(n, world)
            assert x_local.size(0) <= per, (
                f"dense row block has {x_local.size(0)} rows, expected at most ceil({n} / {world}) = {per}")
            x_local = self.pad_rows(x_local, per)   # equal block heights: what all_gather / reduce_scatter need
        return _AllGatherRows.apply(x_local, self.group)

    def local_spmm(self, x_full: Tensor) -> Tensor:
        n = self.a.sparse_size(1)
        if x_full.size(-2) > n:      # rows added by the padding of the last block
            x_full = x_full.narrow(-2, 0, n)
        return matmul(self.a, x_full, self.reduce)

    def __call__(self, x_local: Tensor) -> Tensor:
        return self.local_spmm(self.gather_dense(x_local))


def split_column_chunks(rowptr: Tensor, col: Tensor, value: Optional[Tensor], block: int, world: int,
                        chunks: int) -> Tuple[List[Tuple[Tensor, Tensor, Optional[Tensor]]], int]:
    """Split a CSR row block (columns over world * block dense rows) into `chunks` column chunks for the pipelined
    gather. Chunk c holds the entries whose column lies in rows [c*mc, (c+1)*mc) of SOME rank's block (mc = block /
    chunks), with the columns re-indexed to the chunk-major layout of the gathered operand:
        col = p * block + c * mc + i   ->   c * (world * mc) + p * mc + i.
    Returns ([(rowptr_c, col_c, value_c)], mc). Entry order inside a row is preserved. Structure-only set-up work."""
    assert block % chunks == 0, f"block height {block} is not divisible by {chunks} chunks"
    mc = block // chunks
    M = rowptr.numel() - 1
    counts = rowptr[1:] - rowptr[:-1]
    row = torch.repeat_interleave(torch.arange(M, device=col.device), counts)
    p = torch.div(col, block, rounding_mode="floor")
    within = col - p * block
    c = torch.div(within, mc, rounding_mode="floor")
    new_col = c * (world * mc) + p * mc + (within - c * mc)
    out = []
    for k in range(chunks):
        sel = (c == k).nonzero().view(-1)
        rp = torch.zeros(M + 1, dtype=torch.long, device=col.device)
        torch.cumsum(torch.bincount(row[sel], minlength=M), 0, out=rp[1:])
        out.append((rp, new_col[sel].contiguous(), None if value is None else value[sel].contiguous()))
    return out, mc


class PipelinedRowShardedSpMM:
    """Sum-SpMM of this rank's row block with a dense operand that is gathered INSIDE the step, chunk by chunk,
    while the chunks that have already landed are being multiplied (forward only; CUDA / NCCL).

    `block` = rows of the dense operand every rank holds (all equal); `chunks` = pipeline depth."""

    def __init__(self, a_local: SparseTensor, block: int, chunks: int = 4, group=None):
        self.group = group
        self.world = _world(group)
        self.block = block
        self.chunks = chunks
        rowptr, col, value = a_local.csr()
        assert a_local.sparse_size(1) == self.world * block
        self.M = a_local.sparse_size(0)
        self.parts, self.mc = split_column_chunks(rowptr, col, value, block, self.world, chunks)
        self._x = None
        self._partial = None

    def _buffers(self, x_local: Tensor):
        F = x_local.size(1)
        if self._x is None or self._x.dtype != x_local.dtype or self._x.size(-1) != F:
            self._x = x_local.new_empty((self.chunks, self.world * self.mc, F))
            self._partial = torch.empty((self.M, F), dtype=torch.float32, device=x_local.device)
        return self._x, self._partial

    def __call__(self, x_local: Tensor) -> Tensor:
        assert x_local.dim() == 2 and x_local.size(0) == self.block and x_local.is_contiguous()
        xg, partial = self._buffers(x_local)
        mc, C = self.mc, self.chunks
        works = []
        for c in range(C):
            src = x_local[c * mc:(c + 1) * mc]
            if self.world > 1:
                works.append(dist.all_gather_into_tensor(xg[c], src, group=self.group, async_op=True))
            else:
                xg[c].copy_(src)
        out = torch.empty((self.M, x_local.size(1)), dtype=x_local.dtype, device=x_local.device)
        x_flat = xg.view(C * self.world * mc, x_local.size(1))
        parts = [(rp, cl, None if v is None else v.to(x_local.dtype)) for rp, cl, v in self.parts]
        for c in range(C):
            if works:
                works[c].wait()          # the compute stream waits for chunk c only
            mode = 1 if c == 0 else (3 if c == C - 1 else 2)
            rp, cl, v = parts[c]
            if C == 1:
                return ops.spmm_fw(rp, cl, v, x_flat, "sum")[0]
            ops.spmm_fw_acc(rp, cl, v, x_flat, partial, out, mode)
        return out
