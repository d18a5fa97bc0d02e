"""Drop-in surface: functional spmm (test/test_spmm.py), operator registry, argument rules and
error behaviour of the reference's operator layer (csrc/spmm.cpp:64-72)."""
import pytest
import torch

import pytorch_sparse_b200 as ts
from pytorch_sparse_b200 import SparseTensor, add, mul   # module level: TorchScript resolves annotations from globals
from pytorch_sparse_b200.matmul import matmul, spspmm

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("dtype", [torch.half, torch.float, torch.double, torch.int, torch.long, torch.bfloat16])
def test_functional_spmm_known_answer(dtype):
    """test/test_spmm.py:10-19"""
    row = torch.tensor([0, 0, 1, 2, 2], device=DEV)
    col = torch.tensor([0, 2, 1, 0, 1], device=DEV)
    index = torch.stack([row, col], dim=0)
    value = torch.tensor([1, 2, 4, 1, 3], dtype=dtype, device=DEV)
    x = torch.tensor([[1, 4], [2, 5], [3, 6]], dtype=dtype, device=DEV)
    out = ts.spmm(index, value, 3, 3, x)
    assert out.tolist() == [[7, 16], [8, 20], [7, 19]]


def test_functional_spmm_unsorted_duplicates_and_grad():
    g = torch.Generator().manual_seed(0)
    E, M, N, K = 500, 40, 30, 16
    row = torch.randint(M, (E,), generator=g).to(DEV)
    col = torch.randint(N, (E,), generator=g).to(DEV)
    value = torch.randn(E, generator=g, dtype=torch.float64).to(DEV).requires_grad_()
    x = torch.randn(N, K, generator=g, dtype=torch.float64).to(DEV).requires_grad_()
    out = ts.spmm(torch.stack([row, col]), value, M, N, x)
    dense = torch.zeros(M, N, dtype=torch.float64, device=DEV).index_put((row, col), value.detach(), accumulate=True)
    assert torch.allclose(out, dense @ x.detach(), atol=1e-10)
    go = torch.randn(M, K, generator=g, dtype=torch.float64).to(DEV)
    out.backward(go)
    assert torch.allclose(x.grad, dense.t() @ go, atol=1e-10)
    exp_gv = (go[row] * x.detach()[col]).sum(-1)
    assert torch.allclose(value.grad, exp_gv, atol=1e-10)


def test_registered_ops_and_optional_argument_rules():
    src = ts.SparseTensor.from_dense(torch.eye(4, device=DEV))
    rowptr, col, value = src.csr()
    x = torch.randn(4, 8, device=DEV)
    out = torch.ops.tsb200.spmm_sum(None, rowptr, col, value, None, None, x)
    assert torch.allclose(out, x)
    if ts.torch_ops.REGISTERED.get("torch_sparse"):
        out = torch.ops.torch_sparse.spmm_sum(None, rowptr, col, value, None, None, x)
        assert torch.allclose(out, x)
        o, a = torch.ops.torch_sparse.spmm_max(rowptr, col, value, x)
        assert torch.allclose(o, x) and a.dtype == torch.long
    xg = x.clone().requires_grad_()
    with pytest.raises(RuntimeError, match="Argument `row` is missing"):
        torch.ops.tsb200.spmm_sum(None, rowptr, col, value, None, None, xg)
    with pytest.raises(RuntimeError, match="must be CUDA tensor"):
        torch.ops.tsb200.spmm_sum(None, rowptr.cpu(), col.cpu(), value.cpu(), None, None, x.cpu())
    with pytest.raises(ValueError):
        ts.matmul(src, x, "prod")


def test_cpu_built_tensor_moves_to_gpu():
    row = torch.tensor([1, 0, 1, 0])
    col = torch.tensor([0, 1, 1, 0])
    val = torch.tensor([1., 2., 3., 4.])
    a = ts.SparseTensor(row=row, col=col, value=val, sparse_sizes=(2, 2)).cuda()
    x = torch.eye(2, device=DEV)
    assert (a @ x).tolist() == [[4., 2.], [1., 3.]]
    assert (a.t() @ x).tolist() == [[4., 1.], [2., 3.]]


def test_host_buffer_spmm(oracle):
    from pytorch_sparse_b200 import ops
    from util import random_csr
    row, rowptr, col = random_csr(5000, 3000, 9, seed=4, power_law=True)
    g = torch.Generator().manual_seed(1)
    value = torch.randn(col.numel(), generator=g)
    mat = torch.randn(3000, 32, generator=g)
    for reduce in ("sum", "max"):
        out, arg = ops.spmm_fw_host(rowptr, col, value, mat, reduce)
        dout, darg = ops.spmm_fw(rowptr.to(DEV), col.to(DEV), value.to(DEV), mat.to(DEV), reduce)
        assert torch.equal(out, dout.cpu())
        if arg is not None:
            assert torch.equal(arg, darg.cpu())


def test_host_buffer_spmm_chunked_pipeline():
    """Large enough for the pipelined host path (8 row chunks over two upload streams, compute and download
    streams): bit-identical to the device-resident call."""
    from pytorch_sparse_b200 import ops
    from util import fast_random_csr
    M, N, F = 120_000, 90_000, 16
    _, rowptr, col = fast_random_csr(M, N, 12, 7, DEV)
    assert col.numel() >= (1 << 20)
    g = torch.Generator(device=DEV).manual_seed(2)
    value = torch.randn(col.numel(), generator=g, device=DEV)
    mat = torch.randn(N, F, generator=g, device=DEV)
    rp_h, col_h, val_h, mat_h = rowptr.cpu(), col.cpu(), value.cpu(), mat.cpu()
    for red in ("sum", "max"):
        ref, ref_arg = ops.spmm_fw(rowptr, col, value, mat, red)
        out, arg = ops.spmm_fw_host(rp_h, col_h, val_h, mat_h, red)
        assert torch.equal(out, ref.cpu()), red
        if arg is not None:
            assert torch.equal(arg, ref_arg.cpu()), red


def test_torchscript_function_over_registered_ops():
    """A TorchScript function calling the registered ops computes the same as the Python API
    (the reference's scripted classes call torch.ops.torch_sparse.* the same way, storage.py:193,209,376)."""
    from typing import Optional

    @torch.jit.script
    def scripted(rowptr: torch.Tensor, col: torch.Tensor, value: Optional[torch.Tensor], x: torch.Tensor):
        row = torch.ops.torch_sparse.ptr2ind(rowptr, col.numel())
        out = torch.ops.torch_sparse.spmm_sum(row, rowptr, col, value, None, None, x)
        mx, arg = torch.ops.torch_sparse.spmm_max(rowptr, col, value, x)
        return row, out, mx, arg

    g = torch.Generator().manual_seed(3)
    M, N, K = 40, 30, 16
    dense = torch.randn(M, N, generator=g) * (torch.rand(M, N, generator=g) < 0.2)
    a = ts.SparseTensor.from_dense(dense.to("cuda:0"))
    x = torch.randn(N, K, generator=g).to("cuda:0")
    rowptr, col, value = a.csr()
    row, out, mx, arg = scripted(rowptr, col, value, x)
    assert torch.equal(row, a.storage.row())
    assert torch.allclose(out, dense.to("cuda:0") @ x, atol=1e-5)
    assert torch.equal(mx, ts.matmul(a, x, "max"))


def test_torchscript_functional_pipeline_over_every_native_op():
    """coalesce -> ind2ptr -> spspmm -> segment_reduce from ONE TorchScript function (torch.ops.tsb200.*): the
    reference's scripted Python layer is built the same way on its registered ops."""
    @torch.jit.script
    def scripted(index: torch.Tensor, value: torch.Tensor, m: int):
        row, col, val = torch.ops.tsb200.coalesce(index[0], index[1], value, m, m, "add")
        assert val is not None
        rowptr = torch.ops.tsb200.ind2ptr(row, m)
        perm, colptr, row_csc = torch.ops.tsb200.csr2csc(row, col, m, m)
        rp, r, c, v = torch.ops.tsb200.spspmm(rowptr, col, val, rowptr, col, val, m, m, m, True)
        deg = torch.ops.tsb200.segment_reduce(rowptr, val, "sum", None, None)
        return row, col, val, rp, r, c, v, deg, colptr

    g = torch.Generator().manual_seed(11)
    m, E = 50, 400
    index = torch.randint(m, (2, E), generator=g).to("cuda:0")
    value = torch.randn(E, generator=g, dtype=torch.float64).to("cuda:0")
    row, col, val, rp, r, c, v, deg, colptr = scripted(index, value, m)
    dense = torch.zeros(m, m, dtype=torch.float64, device="cuda:0").index_put((index[0], index[1]), value, accumulate=True)
    a = ts.SparseTensor(row=row, col=col, value=val, sparse_sizes=(m, m), is_sorted=True)
    assert torch.allclose(a.to_dense(), dense)
    prod = ts.SparseTensor(row=r, col=c, value=v, sparse_sizes=(m, m), is_sorted=True).to_dense()
    assert torch.allclose(prod, dense @ dense, atol=1e-10)
    assert torch.allclose(deg, dense.sum(1)) and torch.equal(colptr, a.storage.colptr())


def test_torchscript_functions_over_sparse_tensor_objects():
    """SparseStorage / SparseTensor are TorchScript classes like the reference's (torch_sparse/storage.py:21,
    tensor.py:12): scripted functions build, combine and multiply SparseTensors on the GPU and agree with eager."""
    torch.jit.script(spspmm)          # test/test_matmul.py:79

    @torch.jit.script
    def layer(edge_index: torch.Tensor, w: torch.Tensor, n: int, x: torch.Tensor):
        a = SparseTensor(row=edge_index[0], col=edge_index[1], value=w, sparse_sizes=(n, n))
        a = a.coalesce("sum")
        deg = a.storage.rowcount().to(x.dtype).clamp(min=1).pow(-1.0).view(-1, 1)
        y = matmul(mul(a, deg), x, "sum")
        sym = add(a, a.set_value(a.storage.value(), layout="coo"))
        return y, matmul(a, a, "sum"), sym

    g = torch.Generator().manual_seed(21)
    n, E = 60, 500
    ei = torch.randint(n, (2, E), generator=g).to(DEV)
    w = torch.randn(E, generator=g, dtype=torch.float64).to(DEV)
    x = torch.randn(n, 8, generator=g, dtype=torch.float64).to(DEV)
    y, sq, sym = layer(ei, w, n, x)
    dense = torch.zeros(n, n, dtype=torch.float64, device=DEV).index_put((ei[0], ei[1]), w, accumulate=True)
    cnt = (torch.zeros(n, n, device=DEV).index_put((ei[0], ei[1]), torch.ones(E, device=DEV), accumulate=True) > 0).sum(1)
    assert torch.allclose(y, (dense / cnt.clamp(min=1).view(-1, 1)) @ x, atol=1e-10)
    assert torch.allclose(sq.to_dense(), dense @ dense, atol=1e-10)
    assert torch.allclose(sym.to_dense(), 2 * dense, atol=1e-12)
    # eager objects go into scripted functions and come back as ordinary SparseTensors
    a = ts.SparseTensor(row=ei[0], col=ei[1], value=w, sparse_sizes=(n, n)).coalesce()

    @torch.jit.script
    def jit_add(A: SparseTensor, B: SparseTensor) -> SparseTensor:
        return add(A, B)

    assert jit_add(a, a) == a + a
