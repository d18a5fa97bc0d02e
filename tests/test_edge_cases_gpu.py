"""Edge cases of the operator boundary: empty / degenerate extents, batch dimensions on the vector
path, views (non-contiguous, sliced => 8-byte aligned only), dtype mismatches, very wide rows."""
import pytest
import torch

import pytorch_sparse_b200 as ts
from pytorch_sparse_b200 import ops
from util import random_csr

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_empty_and_degenerate(oracle):
    rowptr = torch.zeros(6, dtype=torch.long, device=DEV)          # 5 rows, no nnz
    col = torch.empty(0, dtype=torch.long, device=DEV)
    x = torch.randn(4, 16, device=DEV)
    for red in ("sum", "mean", "min", "max"):
        out, arg = ops.spmm_fw(rowptr, col, None, x, red)
        assert out.shape == (5, 16) and out.eq(0).all()
        if arg is not None:
            assert arg.eq(0).all()                                     # sentinel E == 0
    out, _ = ops.spmm_fw(torch.zeros(1, dtype=torch.long, device=DEV), col, None, x, "sum")   # M == 0
    assert out.shape == (0, 16)
    out, _ = ops.spmm_fw(rowptr, col, None, torch.randn(4, 0, device=DEV), "sum")               # K == 0
    assert out.shape == (5, 0)
    r, c, v = ops.coalesce(col, col, None, 3, 3)
    assert r.numel() == 0 and v is None
    a = ts.SparseTensor(row=col, col=col, sparse_sizes=(3, 3))
    assert (a @ a).nnz() == 0
    assert ops.ptr2ind(rowptr, 0).numel() == 0


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("reduce", ["sum", "mean", "max"])
def test_batch_dims_on_vector_path(oracle, dtype, reduce):
    row, rowptr, col = random_csr(70, 50, 5, seed=2, empty_rows=(3,), long_rows=[(9, 48)])
    g = torch.Generator().manual_seed(3)
    value = torch.randn(col.numel(), generator=g).to(dtype)
    mat = torch.randn(2, 3, 50, 64, generator=g).to(dtype)          # leading batch dims [2, 3]
    out, arg = ops.spmm_fw(rowptr.to(DEV), col.to(DEV), value.to(DEV), mat.to(DEV), reduce)
    ref, rarg = oracle.spmm(rowptr, col, value.double(), mat.double(), reduce)
    assert out.shape == (2, 3, 70, 64)
    tol = 1e-5 if dtype == torch.float32 else 2e-2
    assert torch.allclose(out.cpu().double(), ref, rtol=tol, atol=tol * 10)
    if arg is not None:
        _, a16 = oracle.spmm(rowptr, col, value, mat, reduce)
        assert torch.equal(arg.cpu(), a16)


def test_views_and_alignment(oracle):
    row, rowptr, col = random_csr(64, 80, 6, seed=5)
    g = torch.Generator().manual_seed(1)
    E = col.numel()
    big_val = torch.randn(E + 3, generator=g).bfloat16().to(DEV)
    value = big_val[1:E + 1]                      # 2-byte aligned view -> falls back to the generic kernel
    big = torch.randn(80, 200, generator=g).bfloat16().to(DEV)
    mat_nc = big[:, 8:136]                        # non-contiguous column slice (made contiguous inside)
    out, _ = ops.spmm_fw(rowptr.to(DEV), col.to(DEV), value, mat_nc, "sum")
    ref, _ = oracle.spmm(rowptr, col, value.cpu().double(), mat_nc.cpu().double(), "sum")
    assert torch.allclose(out.cpu().double(), ref, rtol=2e-2, atol=2e-1)
    col_view = torch.cat([col.new_zeros(1), col]).to(DEV)[1:]        # 8-byte (not 16-byte) aligned indices
    out2, _ = ops.spmm_fw(rowptr.to(DEV), col_view, value.contiguous(), mat_nc, "sum")
    assert torch.equal(out2, out) or torch.allclose(out2.float(), out.float(), rtol=2e-2, atol=2e-1)
    # transposed dense operand
    x = torch.randn(32, 80, generator=g).to(DEV).t()                 # [80, 32] non-contiguous
    out3, _ = ops.spmm_fw(rowptr.to(DEV), col.to(DEV), None, x, "sum")
    ref3, _ = oracle.spmm(rowptr, col, None, x.cpu().contiguous(), "sum")
    assert torch.allclose(out3.cpu(), ref3, rtol=1e-5, atol=1e-5)


def test_value_dtype_follows_dense_operand():
    """torch_sparse/matmul.py:16-17 casts `value.to(other.dtype)` before the op."""
    a = ts.SparseTensor.from_dense(torch.eye(4, dtype=torch.float64, device=DEV))
    x = torch.randn(4, 8, device=DEV).bfloat16()
    y = a @ x
    assert y.dtype == torch.bfloat16 and torch.equal(y, x)
    with pytest.raises(RuntimeError):
        ops.spmm_fw(*a.csr()[:2], a.storage.value(), x, "sum")     # op level: dtype mismatch is an error


@pytest.mark.parametrize("K", [1024 + 8, 2048])
def test_very_wide_rows_column_tiled(oracle, K):
    row, rowptr, col = random_csr(20, 30, 4, seed=8)
    g = torch.Generator().manual_seed(4)
    value = torch.randn(col.numel(), generator=g).bfloat16()
    mat = torch.randn(30, K, generator=g).bfloat16()
    for red in ("sum", "max"):
        out, arg = ops.spmm_fw(rowptr.to(DEV), col.to(DEV), value.to(DEV), mat.to(DEV), red)
        ref, _ = oracle.spmm(rowptr, col, value.double(), mat.double(), red)
        assert torch.allclose(out.cpu().double(), ref, rtol=2e-2, atol=5e-2)


def test_sum_has_no_nan_leak_from_unused_lanes():
    """inactive lanes / padded steps must contribute exactly +0 even if the dense operand holds inf/nan
    in rows that are not referenced."""
    rowptr = torch.tensor([0, 1, 3], device=DEV)
    col = torch.tensor([1, 1, 2], device=DEV)
    mat = torch.full((4, 128), float("nan"), device=DEV).bfloat16()
    mat[1] = 1.0
    mat[2] = 2.0
    out, _ = ops.spmm_fw(rowptr, col, torch.ones(3, device=DEV).bfloat16(), mat, "sum")
    assert out[0].eq(1).all() and out[1].eq(3).all()


def test_narrow_rows_group_per_row_kernel(oracle):
    """LPR <= 8 shapes take spmm_gpr_kernel: ragged rows inside a 32-row item, long rows deferred."""
    long_rows = [(3, 300), (40, 257)]
    row, rowptr, col = random_csr(101, 400, 7, seed=12, power_law=True, empty_rows=(0, 50, 100), long_rows=long_rows)
    g = torch.Generator().manual_seed(6)
    for dtype, K in ((torch.bfloat16, 32), (torch.bfloat16, 8), (torch.float32, 16), (torch.float16, 64), (torch.float32, 4)):
        value = torch.randn(col.numel(), generator=g).to(dtype)
        mat = torch.randn(400, K, generator=g).to(dtype)
        for red in ("sum", "mean"):
            for v in (value, None):
                out, _ = ops.spmm_fw(rowptr.to(DEV), col.to(DEV), None if v is None else v.to(DEV), mat.to(DEV), red)
                ref, _ = oracle.spmm(rowptr, col, None if v is None else v.double(), mat.double(), red)
                bound, _ = oracle.spmm(rowptr, col, None if v is None else v.double().abs(), mat.double().abs(), "sum")
                tol = 1e-5 if dtype == torch.float32 else 1e-2
                assert ((out.cpu().double() - ref).abs() <= tol * bound + 1e-30).all(), (dtype, K, red)
