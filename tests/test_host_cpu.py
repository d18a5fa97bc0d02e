"""CPU-only checks: the C-ABI library loads and exports every symbol of include/tsb200.h, argument
validation that needs no GPU, the Python host logic (autograd wiring, optional-argument rules,
cache management, row sharding over gloo) with the oracle standing in for the CUDA kernels."""
import os
import re
import subprocess
import sys
from pathlib import Path

import pytest
import torch

import pytorch_sparse_b200 as ts
from pytorch_sparse_b200 import _lib, ops

ROOT = Path(__file__).resolve().parent.parent


def test_library_exports_every_declared_symbol():
    header = (ROOT / "include" / "tsb200.h").read_text()
    declared = set(re.findall(r"TSB200_API\s+[\w\s\*]+?\b(tsb200_\w+)\s*\(", header))
    assert len(declared) >= 19
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    for name in declared:
        assert hasattr(_lib.lib, name), name
    assert _lib.lib.tsb200_version() == 100
    assert _lib.strerror(0) == "success" and "workspace" in _lib.strerror(-3)


def test_argument_validation_without_gpu():
    lib = _lib.lib
    assert lib.tsb200_spmm_fw(None, None, None, None, None, None, 1, -1, 1, 1, 1, 0, 0, None, 0, None) == -1
    assert lib.tsb200_spmm_fw(None, None, None, None, None, None, 1, 4, 4, 4, 4, 0, 9, None, 0, None) == -1
    assert lib.tsb200_spmm_fw(None, None, None, None, None, None, 1, 4, 4, 4, 4, 99, 0, None, 0, None) == -1
    assert lib.tsb200_spmm_fw(None, None, None, None, None, None, 1, 0, 4, 4, 0, 0, 0, None, 0, None) == 0  # empty
    assert lib.tsb200_spmm_fw_workspace_bytes(1, 1000, 128, 16000, 3, 0) > 0
    assert lib.tsb200_spmm_fw_workspace_bytes(1, 1000, 128, 16000, 5, 0) == 0      # int64: generic kernel
    assert lib.tsb200_coalesce_workspace_bytes(1000, 10, 10) > 16 * 1000
    assert lib.tsb200_spmm_value_bw(None, None, None, None, None, None, 1, 4, 4, 4, 4, 0, 2, None, 0, None) == -1  # reduce=min


def test_ops_reject_cpu_tensors():
    rowptr, col = torch.tensor([0, 1]), torch.tensor([0])
    with pytest.raises(RuntimeError, match="must be CUDA tensor"):
        ops.spmm_fw(rowptr, col, None, torch.ones(1, 4), "sum")
    with pytest.raises(RuntimeError, match="must be CUDA tensor"):
        ops.ind2ptr(col, 1)
    with pytest.raises(RuntimeError, match="must be CUDA tensor"):
        ops.coalesce(col, col, None, 1, 1)
    a = ts.SparseTensor(row=col, col=col, value=torch.ones(1), sparse_sizes=(1, 1))
    with pytest.raises(RuntimeError, match="must be CUDA tensor"):
        a @ torch.ones(1, 4)
    with pytest.raises(RuntimeError, match="must be CUDA tensor"):
        a @ a                      # SpSpMM
    with pytest.raises(RuntimeError, match="must be CUDA tensor"):
        a.to_symmetric()           # cat + coalesce: arithmetic, no CPU path
    with pytest.raises(RuntimeError, match="must be CUDA tensor"):
        ts.spadd(torch.stack([col, col]), torch.ones(1), torch.stack([col, col]), torch.ones(1), 1, 1)


def test_storage_host_bookkeeping_matches_reference_known_answers():
    """test/test_storage.py:27-92 on CPU-resident tensors (construction convenience path)."""
    st = ts.SparseStorage(row=torch.tensor([0, 0, 1, 1]), col=torch.tensor([1, 0, 1, 0]),
                          value=torch.tensor([2., 1., 4., 3.]))
    assert st.row().tolist() == [0, 0, 1, 1] and st.col().tolist() == [0, 1, 0, 1]
    assert st.value().tolist() == [1, 2, 3, 4] and st.sparse_sizes() == (2, 2)
    assert st.num_cached_keys() == 0
    st.fill_cache_()
    assert st._rowptr.tolist() == [0, 2, 4] and st._colptr.tolist() == [0, 2, 4]
    assert st._csr2csc.tolist() == [0, 2, 1, 3] and st._csc2csr.tolist() == [0, 2, 1, 3]
    assert st.cached_keys() == ["rowcount", "colptr", "colcount", "csr2csc", "csc2csr"]
    t = ts.SparseTensor.from_storage(st).t()
    assert t.storage.row().tolist() == [0, 0, 1, 1] and t.storage.value().tolist() == [1, 3, 2, 4]
    assert st.clear_cache_().num_cached_keys() == 0


@pytest.fixture
def oracle_kernels(monkeypatch, oracle):
    """Route the tensor-level kernels of ops.py through the oracle so the HOST logic (autograd
    Functions, argument rules, matmul dispatch) can be exercised on a machine without a GPU."""
    monkeypatch.setattr(ops, "spmm_fw", lambda rowptr, col, value, mat, reduce: oracle.spmm(rowptr, col, value, mat, reduce))
    monkeypatch.setattr(ops, "spmm_value_bw", oracle.spmm_value_bw)
    monkeypatch.setattr(ops, "spmm_minmax_bw",
                        lambda col, value, mat, go, arg, nv, nm: oracle.spmm_minmax_bw(col, value, mat, go, arg, nv, nm))
    return oracle


@pytest.mark.parametrize("reduce", ["sum", "add", "mean", "min", "max"])
def test_autograd_wiring_like_reference_test(oracle_kernels, reduce):
    """test/test_matmul.py:12-51 end to end through SparseTensor / matmul / autograd Functions."""
    torch.manual_seed(0)
    dense = torch.randn(10, 8, dtype=torch.float64)
    dense[2:4, :] = 0
    dense[:, 2:4] = 0
    src = ts.SparseTensor.from_dense(dense).requires_grad_()
    row, col, value = src.coo()
    other = torch.randn(2, 8, 2, dtype=torch.float64, requires_grad=True)
    v2, o2 = value.detach().clone().requires_grad_(), other.detach().clone().requires_grad_()
    src_col = o2.index_select(-2, col) * v2.unsqueeze(-1)
    if reduce in ("sum", "add", "mean"):
        exp = torch.zeros(2, 10, 2, dtype=torch.float64).index_add(-2, row, src_col)
        if reduce == "mean":
            exp = exp / torch.bincount(row, minlength=10).clamp(min=1).view(1, 10, 1)
    else:
        fill = float("inf") if reduce == "min" else float("-inf")
        idx = row.view(1, -1, 1).expand_as(src_col)
        exp = torch.full((2, 10, 2), fill, dtype=torch.float64).scatter_reduce(
            -2, idx, src_col, reduce="amin" if reduce == "min" else "amax", include_self=True)
        exp = torch.where(torch.isinf(exp), torch.zeros_like(exp), exp)
    go = torch.randn_like(exp)
    exp.backward(go)
    out = ts.matmul(src, other, reduce)
    out.backward(go)
    assert torch.allclose(exp, out, atol=1e-10)
    assert torch.allclose(v2.grad, value.grad, atol=1e-10)
    assert torch.allclose(o2.grad, other.grad, atol=1e-10)


def test_optional_argument_rules(oracle_kernels):
    """csrc/spmm.cpp:64-72: row / colptr / csr2csc are required exactly when a gradient needs them."""
    src = ts.SparseTensor.from_dense(torch.eye(3, dtype=torch.float64))
    rowptr, col, value = src.csr()
    x = torch.randn(3, 2, dtype=torch.float64)
    assert torch.allclose(ops.spmm_sum(None, rowptr, col, value, None, None, x), x)
    with pytest.raises(RuntimeError, match="Argument `row` is missing"):
        ops.spmm_sum(None, rowptr, col, value, None, None, x.clone().requires_grad_())
    with pytest.raises(RuntimeError, match="Argument `row` is missing"):
        ops.spmm_sum(None, rowptr, col, value.clone().requires_grad_(), None, None, x)
    with pytest.raises(RuntimeError, match="Argument `colptr` is missing"):
        ops.spmm_sum(src.storage.row(), rowptr, col, value, None, None, x.clone().requires_grad_())
    with pytest.raises(RuntimeError, match="Argument `rowcount` is missing"):
        ops.spmm_mean(src.storage.row(), rowptr, col, value, None, src.storage.colptr(), src.storage.csr2csc(),
                      x.clone().requires_grad_())
    with pytest.raises(ValueError):
        ts.matmul(src, x, "prod")
    with pytest.raises(NotImplementedError):
        ts.matmul(src, src, "max")
    o, a = ops.spmm_max(rowptr, col, None, x)
    assert a.dtype == torch.long and not a.requires_grad


def test_matmul_materialises_only_needed_caches(oracle_kernels):
    """torch_sparse/matmul.py:19-25: csr2csc/colptr are built only when `other` needs a gradient."""
    src = ts.SparseTensor.from_dense(torch.rand(6, 5, dtype=torch.float64).round())
    x = torch.randn(5, 3, dtype=torch.float64)
    src @ x
    assert src.storage.num_cached_keys() == 0
    (src @ x.requires_grad_()).sum().backward()
    assert set(src.storage.cached_keys()) >= {"colptr", "csr2csc"}


_GLOO = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, "%(root)s"); sys.path.insert(0, "%(root)s/tests")
import oracle
import pytorch_sparse_b200 as ts
from pytorch_sparse_b200 import ops
from pytorch_sparse_b200.parallel import RowShardedSpMM, PipelinedRowShardedSpMM
ops.spmm_fw = lambda rowptr, col, value, mat, reduce: oracle.spmm(rowptr, col, value, mat, reduce)
ops.spmm_value_bw = oracle.spmm_value_bw
def _acc(rowptr, col, value, mat, partial, out, mode):      # CPU stand-in of tsb200_spmm_fw_acc (oracle kernels)
    res = oracle.spmm(rowptr, col, value, mat, "sum")[0].float()
    if mode == 1: partial.copy_(res)
    elif mode == 2: partial.add_(res)
    else: out.copy_((partial + res).to(out.dtype))
ops.spmm_fw_acc = _acc
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo")
g = torch.Generator().manual_seed(0)
K = 8
for M, N in ((64, 64), (65, 67)):        # 65 / 67: the last row blocks are shorter -> padded gather
    dense = (torch.rand(M, N, generator=g) < 0.1).double() * torch.randn(M, N, generator=g, dtype=torch.float64)
    x = torch.randn(N, K, generator=g, dtype=torch.float64)
    full = ts.SparseTensor.from_dense(dense)
    a_local = RowShardedSpMM.partition(full, rank, world)
    pm, pn = RowShardedSpMM.block_rows(M, world), RowShardedSpMM.block_rows(N, world)
    x_local = x[rank * pn:(rank + 1) * pn].clone().requires_grad_()
    op = RowShardedSpMM(a_local.requires_grad_(), "sum")
    y_local = op(x_local)
    ref = dense @ x
    assert torch.allclose(y_local, ref[rank * pm:(rank + 1) * pm], atol=1e-12), "forward shard mismatch"
    go = torch.randn(M, K, generator=g, dtype=torch.float64)
    y_local.backward(go[rank * pm:(rank + 1) * pm])
    gx = dense.t() @ go                      # needs the reduce-scatter of the per-rank partials
    assert x_local.grad.shape == x_local.shape
    assert torch.allclose(x_local.grad, gx[rank * pn:(rank + 1) * pn], atol=1e-12), "grad_X shard mismatch"
    assert a_local.storage.value().grad is not None
# pipelined gather + column-chunk SpMM (the gather is part of the step)
M, block, C = 48, 24, 4
N = world * block
dense = ((torch.rand(world * M, N, generator=g) < 0.15).float() * torch.randn(world * M, N, generator=g))
x = torch.randn(N, K, generator=g)
a_local = ts.SparseTensor.from_dense(dense[rank * M:(rank + 1) * M])
for split in ("feature", "column"):
    pipe = PipelinedRowShardedSpMM(a_local, block=block, chunks=C, split=split)
    for _ in range(2):                           # buffers are reused across steps
        y = pipe(x[rank * block:(rank + 1) * block].contiguous())
        assert torch.allclose(y, (dense @ x)[rank * M:(rank + 1) * M], atol=1e-4), f"pipelined ({split}) shard mismatch"
xs = pipe.to_sliced(x[rank * block:(rank + 1) * block].contiguous())
assert torch.equal(pipe.from_sliced(xs), x[rank * block:(rank + 1) * block])
dist.barrier(); dist.destroy_process_group()
print("rank", rank, "ok")
'''


def test_row_sharded_spmm_gloo_world2(oracle, tmp_path):
    """N>1 path on CPU: 2 ranks over gloo, row-block partition (equal and ragged block heights), all-gather of
    X, local SpMM (oracle kernels), reduce-scatter of grad_X — forward and backward match the dense product; and
    the pipelined variant (chunked gather + column-chunk SpMM with an fp32 partial)."""
    script = tmp_path / "gloo_worker.py"
    script.write_text(_GLOO % {"root": str(ROOT)})
    env = dict(os.environ, TSB200_REGISTER_TORCH_SPARSE="0")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29531", str(script)]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env)
    assert out.returncode == 0, (out.stdout + out.stderr)[-3000:]
    assert out.stdout.count("ok") == 2


def test_pinned_output_pool_never_aliases_live_results():
    """spmm_fw_host's output pool: a buffer is reused only after every tensor derived from an earlier
    result is gone (pin=False here: same logic, no CUDA needed)."""
    a = ops._pinned_empty((4, 8), torch.bfloat16, pin=False)
    b = ops._pinned_empty((4, 8), torch.bfloat16, pin=False)
    assert a.data_ptr() != b.data_ptr()
    pa, view = a.data_ptr(), a[:2]
    del a
    c = ops._pinned_empty((4, 8), torch.bfloat16, pin=False)
    assert c.data_ptr() != pa          # a derived view is still alive
    del view
    d = ops._pinned_empty((4, 8), torch.bfloat16, pin=False)
    assert d.data_ptr() == pa          # now it may be reused


def test_index_select_host_bookkeeping_matches_reference():
    """index_select / index_select_nnz (torch_sparse/index_select.py:9-99) are pure index bookkeeping: on CPU
    tensors (storage construction path, no arithmetic) they reproduce the unmodified reference's outputs
    (tests/golden/next_rows2.pt) bit for bit."""
    d = torch.load(Path(__file__).resolve().parent / "golden" / "next_rows2.pt", weights_only=False)
    for name, case in d["cases"].items():
        i = case["in"]
        for tag in ("v", "v2", "nv"):
            v = {"v": i["v"], "v2": i["v2"], "nv": None}[tag]
            a = ts.SparseTensor(row=i["row"], col=i["col"], value=v, sparse_sizes=(i["M"], i["N"]))
            s0, ref = a.index_select(0, i["idx0"]), case[f"sel0_{tag}"]
            assert s0.sparse_sizes() == tuple(ref["sizes"])
            assert torch.equal(s0.storage.rowptr(), ref["rowptr"]) and torch.equal(s0.storage.col(), ref["col"])
            s1, ref = a.index_select(1, i["idx1"]), case[f"sel1_{tag}"]
            assert torch.equal(s1.storage.row(), ref["row"]) and torch.equal(s1.storage.col(), ref["col"])
            assert torch.equal(s1.storage.colptr(), ref["colptr"])
            if v is not None:
                assert torch.equal(s0.storage.value(), case[f"sel0_{tag}"]["value"])
                assert torch.equal(s1.storage.value(), ref["value"])
            for lay in ("coo", "csc"):
                sn, ref = a.index_select_nnz(i["idxe"], lay), case[f"selnnz_{lay}_{tag}"]
                assert torch.equal(sn.storage.row(), ref["row"]) and torch.equal(sn.storage.col(), ref["col"])


def test_torchscript_functions_can_call_the_registered_ops():
    """The reference's Python layer is TorchScript that calls torch.ops.torch_sparse.* (storage.py:193,209,376);
    the same works on the operators this package registers: scripting succeeds and the call reaches the op
    (which rejects CPU tensors — there is no CPU path)."""
    from typing import Optional

    import torch
    import pytorch_sparse_b200  # noqa: F401

    @torch.jit.script
    def scripted(rowptr: torch.Tensor, col: torch.Tensor, value: Optional[torch.Tensor], x: torch.Tensor):
        row = torch.ops.tsb200.ptr2ind(rowptr, col.numel())
        return torch.ops.tsb200.spmm_sum(row, rowptr, col, value, None, None, x)

    assert "tsb200::spmm_sum" in str(scripted.graph) and "tsb200::ptr2ind" in str(scripted.graph)
    with pytest.raises(RuntimeError, match="must be CUDA tensor"):
        scripted(torch.tensor([0, 1]), torch.tensor([0]), None, torch.ones(1, 2))


def test_cuda_version_op_matches_reference_encoding():
    import torch
    import pytorch_sparse_b200  # noqa: F401
    v = torch.ops.tsb200.cuda_version()
    # CUDA_VERSION encoding (csrc/version.cpp:27-41); the reference's import check (torch_sparse/__init__.py:23-37)
    # derives major from the first two digits and must agree with torch's CUDA major
    assert v >= 12000 and int(str(v)[0:2]) == int(torch.version.cuda.split(".")[0])


@torch.jit.script
def _scripted_views(row: torch.Tensor, col: torch.Tensor, value: torch.Tensor, M: int, N: int):
    st = ts.SparseStorage(row=row, col=col, value=value, sparse_sizes=(M, N))
    st.fill_cache_()
    value = st.value()
    assert value is not None
    return (st.row(), st.col(), value, st.rowptr(), st.rowcount(), st.colptr(), st.colcount(), st.csr2csc(),
            st.csc2csr(), st.num_cached_keys())


def test_scripted_storage_matches_reference_views():
    """SparseStorage is a TorchScript class (torch_sparse/storage.py:21): built and queried from COMPILED code on CPU
    tensors (index bookkeeping only), it reproduces the unmodified reference's nine views bit for bit
    (tests/golden/storage_views.pt)."""
    d = torch.load(Path(__file__).resolve().parent / "golden" / "storage_views.pt", weights_only=False)
    i, o = d["in"], d["out"]
    got = _scripted_views(i["row"], i["col"], i["value"], i["M"], i["N"])
    names = ("row", "col", "value", "rowptr", "rowcount", "colptr", "colcount", "csr2csc", "csc2csr")
    for name, t in zip(names, got[:9]):
        assert torch.equal(t, o[name]), name
    assert got[9] == 5


def test_scripted_sparse_tensor_api_on_cpu():
    """SparseTensor as a TorchScript class: scripted code builds tensors, reshapes, resizes, converts and compares —
    everything that is index bookkeeping runs compiled on CPU tensors."""
    from pytorch_sparse_b200 import SparseTensor

    @torch.jit.script
    def f(dense: torch.Tensor):
        idx = dense.nonzero().t()     # (class-level constructors such as from_dense are eager-only conveniences)
        a = SparseTensor(row=idx[0], col=idx[1], value=dense[idx[0], idx[1]], sparse_sizes=(dense.size(0), dense.size(1)))
        b = a.set_value(a.storage.value(), layout="coo").sparse_resize((dense.size(0) + 2, dense.size(1)))
        c = a.sparse_reshape(dense.size(1), dense.size(0))
        return a.to_dense(), b.sizes(), b.storage.rowptr(), c.to_dense(), a.is_symmetric(), a.density(), a.nnz()

    g = torch.Generator().manual_seed(4)
    dense = torch.randn(6, 4, generator=g, dtype=torch.float64) * (torch.rand(6, 4, generator=g) < 0.5)
    back, sizes, rowptr, reshaped, sym, density, nnz = f(dense)
    assert torch.equal(back, dense) and sizes == [8, 4] and rowptr.numel() == 9 and int(rowptr[-1]) == nnz
    assert torch.equal(reshaped, dense.reshape(4, 6)) and sym is False
    assert abs(density - (dense != 0).sum().item() / 24) < 1e-12
    # eager and scripted objects are the same Python class
    a = SparseTensor.from_dense(dense)
    assert isinstance(a, SparseTensor) and a == SparseTensor.from_dense(back)
