"""GPU parity of SpSpMM against the oracle (Gustavson restatement of torch.sparse.mm's contract) and
the reference's known answers (test/test_spspmm.py, test/test_matmul.py:54-79)."""
import pytest
import torch

import pytorch_sparse_b200 as ts
from pytorch_sparse_b200 import ops
from util import random_csr

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(autouse=True, params=["fused", "two_phase"])
def spspmm_mode(request, monkeypatch):
    """Every test runs through both native paths: the single-pass kernel (row placement by decoupled look-back)
    and the symbolic + numeric pair."""
    monkeypatch.setenv("TSB200_SPSPMM", request.param)
    return request.param


@pytest.mark.parametrize("dtype", [torch.float, torch.double])
def test_functional_known_answer(dtype):
    """test/test_spspmm.py:10-22"""
    indexA = torch.tensor([[0, 0, 1, 2, 2], [1, 2, 0, 0, 1]], device=DEV)
    valueA = torch.tensor([1, 2, 3, 4, 5], dtype=dtype, device=DEV)
    indexB = torch.tensor([[0, 2], [1, 0]], device=DEV)
    valueB = torch.tensor([2, 4], dtype=dtype, device=DEV)
    indexC, valueC = ts.spspmm(indexA, valueA, indexB, valueB, 3, 3, 2)
    assert indexC.tolist() == [[0, 1, 2], [0, 1, 1]]
    assert valueC.tolist() == [8, 6, 8]


@pytest.mark.parametrize("dtype", [torch.float, torch.double])
def test_identity_product(dtype):
    """test/test_matmul.py:54-77"""
    src = torch.tensor([[1, 0, 0], [0, 1, 0], [0, 0, 1]], dtype=dtype, device=DEV)
    src = ts.SparseTensor.from_dense(src)
    out = ts.matmul(src, src)
    assert out.sizes() == [3, 3]
    assert out.has_value()
    rowptr, col, value = out.csr()
    assert rowptr.tolist() == [0, 1, 2, 3]
    assert col.tolist() == [0, 1, 2]
    assert value.tolist() == [1, 1, 1]
    src.set_value_(None)
    out = ts.matmul(src, src)
    assert out.sizes() == [3, 3]
    assert not out.has_value()
    rowptr, col, value = out.csr()
    assert rowptr.tolist() == [0, 1, 2, 3]
    assert col.tolist() == [0, 1, 2]


@pytest.mark.parametrize("dtype", [torch.float, torch.double])
def test_orthonormal_rows(dtype):
    """test/test_spspmm.py:25-51: x @ x.t() == I for a 10x16 matrix with orthonormal rows."""
    x = ts.SparseTensor(
        row=torch.tensor([0, 1, 1, 1, 2, 3, 4, 5, 5, 6, 6, 7, 7, 7, 8, 8, 9, 9], device=DEV),
        col=torch.tensor([0, 5, 10, 15, 1, 2, 3, 7, 13, 6, 9, 5, 10, 15, 11, 14, 5, 15], device=DEV),
        value=torch.tensor([1, 3**-0.5, 3**-0.5, 3**-0.5, 1, 1, 1, -2**-0.5, -2**-0.5, -2**-0.5, -2**-0.5, 6**-0.5,
                            -6**0.5 / 3, 6**-0.5, -2**-0.5, -2**-0.5, 2**-0.5, -2**-0.5], dtype=dtype, device=DEV))
    expected = torch.eye(10, device=DEV).to(dtype)
    out = x @ x.to_dense().t()
    assert torch.allclose(out, expected, atol=1e-2)
    out = (x @ x.t()).to_dense()
    assert torch.allclose(out, expected, atol=1e-2)


@pytest.mark.parametrize("M,Kd,N,da,db,dtype", [
    (200, 150, 180, 6, 5, torch.float32),
    (64, 3000, 70000, 12, 40, torch.float32),     # N > 65536
    (50, 400, 700_000, 10, 300, torch.float64),   # N > one 2^18-column window: multi-window path
    (30, 40, 5000, 30, 1000, torch.float32),      # rows wider than the 1024-entry shared staging area
    (300, 64, 96, 28, 28, torch.float32),         # simple-row path, >= 3 products per output column (duplicates)
    (300, 64, 96, 28, 28, torch.float64),
    (257, 500, 262_144, 30, 30, torch.float32),   # simple and general rows mixed (Poisson(30) straddles 32), full window
    (40, 3000, 20_000, 400, 3, torch.float32),    # A rows longer than one 256-entry batch, short B rows
    (64, 200, 1500, 20, 60, torch.float32),       # B rows longer than 32 (general path), output staged in smem
    (37, 50, 33, 5, 4, torch.float64),            # tiny window (1024 bits, 4-word scan chunks)
    (60, 64, 9000, 45, 45, torch.float32),        # products per row straddle the 2048-entry flat/staging limit
    (50, 300, 262_145, 10, 40, torch.float32),    # one column past a full 2^18 window: second window nearly empty
    (20, 500, 3000, 140, 6, torch.float64),       # A rows straddle the 128-entry batch (flat vs multi-batch general)
])
def test_random_vs_oracle(oracle, M, Kd, N, da, db, dtype):
    _, rpa, ca = random_csr(M, Kd, da, seed=1, empty_rows=(0,))
    _, rpb, cb = random_csr(Kd, N, db, seed=2, empty_rows=(1,))
    g = torch.Generator().manual_seed(3)
    va = torch.randn(ca.numel(), generator=g).to(dtype)
    vb = torch.randn(cb.numel(), generator=g).to(dtype)
    rp, r, c, v = ops.spspmm(rpa.to(DEV), ca.to(DEV), va.to(DEV), rpb.to(DEV), cb.to(DEV), vb.to(DEV), M, Kd, N, True)
    orp, orow, oc, ov = oracle.spspmm(rpa, ca, va, rpb, cb, vb, M, Kd, N)
    assert torch.equal(rp.cpu(), orp) and torch.equal(r.cpu(), orow) and torch.equal(c.cpu(), oc)  # bit-exact
    tol = 1e-5 if dtype == torch.float32 else 1e-12
    _, _, _, bound = oracle.spspmm(rpa, ca, va.abs(), rpb, cb, vb.abs(), M, Kd, N)
    assert ((v.cpu() - ov).abs() <= tol * bound + 1e-30).all()
    # structure-only and one-sided values
    rp2, r2, c2, v2 = ops.spspmm(rpa.to(DEV), ca.to(DEV), None, rpb.to(DEV), cb.to(DEV), None, M, Kd, N, False)
    assert v2 is None and torch.equal(c2.cpu(), oc)
    rp3, r3, c3, v3 = ops.spspmm(rpa.to(DEV), ca.to(DEV), va.to(DEV), rpb.to(DEV), cb.to(DEV), None, M, Kd, N, True)
    _, _, _, ov3 = oracle.spspmm(rpa, ca, va, rpb, cb, None, M, Kd, N)
    assert ((v3.cpu() - ov3).abs() <= tol * bound + 1e-30).all()


def test_structural_zeros_kept(oracle):
    """cancelling products stay as explicit zeros (probed torch.sparse.mm behaviour)."""
    A = ts.SparseTensor(row=torch.tensor([0, 0], device=DEV), col=torch.tensor([0, 1], device=DEV),
                        value=torch.tensor([1., -1.], device=DEV), sparse_sizes=(1, 2))
    B = ts.SparseTensor(row=torch.tensor([0, 1], device=DEV), col=torch.tensor([0, 0], device=DEV),
                        value=torch.tensor([2., 2.], device=DEV), sparse_sizes=(2, 1))
    C = A @ B
    assert C.nnz() == 1 and C.storage.value().tolist() == [0.0]


def test_spspmm_reduce_errors():
    src = ts.SparseTensor.eye(3, device=DEV)
    for red in ("mean", "min", "max"):
        with pytest.raises(NotImplementedError):
            ts.matmul(src, src, red)
    with pytest.raises(RuntimeError):
        ts.matmul(src.half(), src.half())   # "sparse_matmul" not implemented for Half


def test_c4_scale_slice_vs_oracle(oracle):
    """BASELINE configs[3] operands (262 144 x 262 144, ~32 nnz/row): the first 16 384 rows of A times the
    FULL B (~16.7M output nnz) against the oracle — structure bit-exact, values within 1e-5 of |A||B|;
    plus whole-problem invariants of the full product (sorted unique columns per row, rowptr consistency)."""
    from util import fast_random_csr
    M = 262_144
    _, rpa, ca = fast_random_csr(M, M, 32, 3, DEV)
    _, rpb, cb = fast_random_csr(M, M, 32, 4, DEV)
    g = torch.Generator(device=DEV).manual_seed(9)
    va = torch.randn(ca.numel(), generator=g, device=DEV)
    vb = torch.randn(cb.numel(), generator=g, device=DEV)
    Ms = 16_384
    ea = int(rpa[Ms])
    rp, r, c, v = ops.spspmm(rpa[:Ms + 1], ca[:ea], va[:ea], rpb, cb, vb, Ms, M, M, True)
    orp, orow, oc, ov = oracle.spspmm(rpa[:Ms + 1], ca[:ea], va[:ea], rpb, cb, vb, Ms, M, M)
    assert torch.equal(rp.cpu(), orp) and torch.equal(r.cpu(), orow) and torch.equal(c.cpu(), oc)
    _, _, _, bound = oracle.spspmm(rpa[:Ms + 1], ca[:ea], va[:ea].abs(), rpb, cb, vb.abs(), Ms, M, M)
    assert ((v.cpu() - ov).abs() <= 1e-5 * bound + 1e-30).all()
    # full product: invariants only (268M nnz)
    rp, r, c, v = ops.spspmm(rpa, ca, va, rpb, cb, vb, M, M, M, True)
    assert int(rp[-1]) == c.numel() == r.numel() == v.numel()
    key = r * M + c
    assert bool((key[1:] > key[:-1]).all())                      # strictly sorted by (row, col) => unique
    assert torch.equal(torch.bincount(r, minlength=M), rp[1:] - rp[:-1])
    del key
