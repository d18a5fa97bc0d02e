"""GPU parity of ind2ptr / ptr2ind / csr2csc / sort-on-construct / coalesce against the oracle and
the reference's known answers (test/test_storage.py, test/test_coalesce.py, test/test_transpose.py)."""
import pytest
import torch

import pytorch_sparse_b200 as ts
from pytorch_sparse_b200 import ops
from pytorch_sparse_b200.storage import SparseStorage
from util import random_csr

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_ind2ptr_known_answer():
    """test/test_storage.py:10-24"""
    row = torch.tensor([2, 2, 4, 5, 5, 6], device=DEV)
    rowptr = torch.ops.tsb200.ind2ptr(row, 8)
    assert rowptr.tolist() == [0, 0, 0, 2, 2, 3, 5, 6, 6]
    assert torch.ops.tsb200.ptr2ind(rowptr, 6).tolist() == [2, 2, 4, 5, 5, 6]
    row = torch.tensor([], dtype=torch.long, device=DEV)
    rowptr = torch.ops.tsb200.ind2ptr(row, 8)
    assert rowptr.tolist() == [0] * 9
    assert torch.ops.tsb200.ptr2ind(rowptr, 0).tolist() == []


@pytest.mark.parametrize("M,N,deg,pl", [(1000, 800, 7, False), (333, 5000, 20, True), (5, 3, 1, False)])
def test_convert_random(oracle, M, N, deg, pl):
    row, rowptr, col = random_csr(M, N, deg, seed=M, power_law=pl, empty_rows=(0, M - 1),
                                  long_rows=[(2, min(N, 700))] if M > 10 else [])
    E = col.numel()
    assert torch.equal(ops.ind2ptr(row.to(DEV), M).cpu(), oracle.ind2ptr(row, M))
    assert torch.equal(ops.ind2ptr(row.to(DEV), M).cpu(), rowptr)
    assert torch.equal(ops.ptr2ind(rowptr.to(DEV), E).cpu(), oracle.ptr2ind(rowptr, E))
    perm, colptr, row_csc = ops.csr2csc(row.to(DEV), col.to(DEV), M, N, True, True)
    ref = oracle.csr2csc(row, col, M)
    assert torch.equal(perm.cpu(), ref)
    assert torch.equal(colptr.cpu(), oracle.ind2ptr(col[ref], N))
    assert torch.equal(row_csc.cpu(), row[ref])


def test_storage_sort_and_caches():
    """test/test_storage.py:27-43 (sort on construct) and :46-92 (cache fill)."""
    row = torch.tensor([0, 0, 1, 1], device=DEV)
    col = torch.tensor([1, 0, 1, 0], device=DEV)
    value = torch.tensor([2., 1., 4., 3.], device=DEV)
    st = SparseStorage(row=row, col=col, value=value)
    assert st.row().tolist() == [0, 0, 1, 1]
    assert st.col().tolist() == [0, 1, 0, 1]
    assert st.value().tolist() == [1, 2, 3, 4]
    assert st.sparse_sizes() == (2, 2)
    assert st.num_cached_keys() == 0
    st.fill_cache_()
    assert st._rowcount.tolist() == [2, 2]
    assert st._rowptr.tolist() == [0, 2, 4]
    assert st._colcount.tolist() == [2, 2]
    assert st._colptr.tolist() == [0, 2, 4]
    assert st._csr2csc.tolist() == [0, 2, 1, 3]
    assert st._csc2csr.tolist() == [0, 2, 1, 3]
    assert st.num_cached_keys() == 5
    st.clear_cache_()
    assert st.num_cached_keys() == 0


def test_storage_coalesce_known_answer():
    """test/test_storage.py:125-141"""
    row = torch.tensor([0, 0, 0, 1, 1], device=DEV)
    col = torch.tensor([0, 1, 1, 0, 1], device=DEV)
    for dtype in (torch.half, torch.float, torch.double, torch.int, torch.long, torch.bfloat16):
        value = torch.tensor([1, 1, 1, 3, 4], dtype=dtype, device=DEV)
        st = SparseStorage(row=row, col=col, value=value)
        assert not st.is_coalesced()
        st = st.coalesce()
        assert st.is_coalesced()
        assert st.row().tolist() == [0, 0, 1, 1]
        assert st.col().tolist() == [0, 1, 0, 1]
        assert st.value().tolist() == [1, 2, 3, 4]


def test_coalesce_known_answers():
    """test/test_coalesce.py:5-33"""
    row = torch.tensor([1, 0, 1, 0, 2, 1], device=DEV)
    col = torch.tensor([0, 1, 1, 1, 0, 0], device=DEV)
    index = torch.stack([row, col], dim=0)
    out, _ = ts.coalesce(index, None, m=3, n=2)
    assert out.tolist() == [[0, 1, 1, 2], [1, 0, 1, 0]]
    value = torch.tensor([[1, 2], [2, 3], [3, 4], [4, 5], [5, 6], [6, 7]], device=DEV)
    out, v = ts.coalesce(index, value, m=3, n=2)
    assert out.tolist() == [[0, 1, 1, 2], [1, 0, 1, 0]]
    assert v.tolist() == [[6, 8], [7, 9], [3, 4], [5, 6]]
    out, v = ts.coalesce(index, value, m=3, n=2, op="max")
    assert out.tolist() == [[0, 1, 1, 2], [1, 0, 1, 0]]
    assert v.tolist() == [[4, 5], [6, 7], [3, 4], [5, 6]]


def test_transpose_known_answers():
    """test/test_transpose.py:10-32"""
    for dtype in (torch.half, torch.float, torch.double, torch.int, torch.long):
        row = torch.tensor([1, 0, 1, 2], device=DEV)
        col = torch.tensor([0, 1, 1, 0], device=DEV)
        index = torch.stack([row, col], dim=0)
        value = torch.tensor([1, 2, 3, 4], dtype=dtype, device=DEV)
        index, value = ts.transpose(index, value, m=3, n=2)
        assert index.tolist() == [[0, 0, 1, 1], [1, 2, 0, 1]]
        assert value.tolist() == [1, 4, 2, 3]
        row = torch.tensor([1, 0, 1, 0, 2, 1], device=DEV)
        col = torch.tensor([0, 1, 1, 1, 0, 0], device=DEV)
        index = torch.stack([row, col], dim=0)
        value = torch.tensor([[1, 2], [2, 3], [3, 4], [4, 5], [5, 6], [6, 7]], dtype=dtype, device=DEV)
        index, value = ts.transpose(index, value, m=3, n=2)
        assert index.tolist() == [[0, 0, 1, 1], [1, 2, 0, 1]]
        assert value.tolist() == [[7, 9], [5, 6], [6, 8], [3, 4]]


@pytest.mark.parametrize("op", ["add", "mean", "min", "max"])
@pytest.mark.parametrize("dtype", [torch.float32, torch.float64, torch.int64])
def test_coalesce_random_vs_oracle(oracle, op, dtype):
    """every key duplicated 1-3 times, shuffled (SURVEY §8d G4 coalesce-correctness input, small)."""
    g = torch.Generator().manual_seed(5)
    M, N, E0 = 400, 300, 5000
    row = torch.randint(M, (E0,), generator=g)
    col = torch.randint(N, (E0,), generator=g)
    reps = torch.randint(1, 4, (E0,), generator=g)
    row, col = row.repeat_interleave(reps), col.repeat_interleave(reps)
    perm = torch.randperm(row.numel(), generator=g)
    row, col = row[perm], col[perm]
    value = torch.randn(row.numel(), 3, generator=g)
    value = (value * 8).round().to(dtype) if dtype == torch.int64 else value.to(dtype)
    r, c, v = ops.coalesce(row.to(DEV), col.to(DEV), value.to(DEV), M, N, op)
    rr, rc, rv = oracle.coalesce(row, col, value, M, N, op)
    assert torch.equal(r.cpu(), rr) and torch.equal(c.cpu(), rc)          # bit-exact structure
    if dtype == torch.int64 or op in ("min", "max"):
        assert torch.equal(v.cpu(), rv)
    else:
        # same (stable, input-order) summation order as the oracle => bit-exact as well
        assert torch.equal(v.cpu(), rv)
    # value=None path
    r2, c2, v2 = ops.coalesce(row.to(DEV), col.to(DEV), None, M, N, op)
    assert v2 is None and torch.equal(r2.cpu(), rr) and torch.equal(c2.cpu(), rc)


def test_coalesce_large_roundtrip():
    """E = 8.4M (config-4 scale): idempotence, sortedness, checksum of values preserved."""
    g = torch.Generator(device=DEV).manual_seed(5)
    M = N = 262_144
    E0 = 4_194_304
    row = torch.randint(M, (E0,), generator=g, device=DEV)
    col = torch.randint(N, (E0,), generator=g, device=DEV)
    perm = torch.randperm(2 * E0, generator=g, device=DEV)
    row2, col2 = torch.cat([row, row])[perm], torch.cat([col, col])[perm]   # every key at least twice
    value = torch.randint(-8, 9, (2 * E0,), generator=g, device=DEV).double()
    r, c, v = ops.coalesce(row2, col2, value, M, N, "add")
    key = r * N + c
    assert bool((key[1:] > key[:-1]).all())                  # strictly sorted => unique
    assert torch.equal(key, torch.unique(row * N + col))     # exactly the distinct keys
    assert v.sum().item() == value.sum().item()              # integer-valued doubles: exact checksum
    r3, c3, v3 = ops.coalesce(r, c, v, M, N, "add")          # idempotent
    assert torch.equal(r3, r) and torch.equal(c3, c) and torch.equal(v3, v)
