"""Gradients w.r.t. the stored values through every op that rides on coalesce / segment reduce, against the
autograd results of the UNMODIFIED reference (tests/golden/grads.pt, written by oracle/gen_golden.py::gen_grads):
coalesce (torch_sparse/coalesce.py:5-25 -> storage.py:436-466), transpose (transpose.py:39-62), spadd (spadd.py:5-18),
add (add.py:38-56), to_symmetric (tensor.py:404-438), sum/mean/min/max over dim 0/1 (reduce.py:36-54) and a
GCN-normalisation flow. float64, tolerance 1e-12 (sums of at most a handful of terms)."""
from pathlib import Path

import pytest
import torch

import pytorch_sparse_b200 as ts
from pytorch_sparse_b200 import ops

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
G = torch.load(Path(__file__).resolve().parent / "golden" / "grads.pt", weights_only=False)
TOL = dict(rtol=1e-12, atol=1e-12)


def _leaf(t):
    return t.to(DEV).clone().requires_grad_()


@pytest.mark.parametrize("tag", ["v1", "v2"])
@pytest.mark.parametrize("op", ["add", "mean", "min", "max"])
def test_coalesce_grad_vs_reference(tag, op):
    i, ref = G["coalesce"]["in"], G["coalesce"][f"{tag}_{op}"]
    v = _leaf(i[tag])
    idx, val = ts.coalesce(i["index"].to(DEV), v, i["M"], i["N"], op=op)
    assert val.grad_fn is not None
    assert torch.equal(idx.cpu(), ref["index"])
    assert torch.allclose(val.detach().cpu(), ref["value"], **TOL)
    val.backward(ref["grad_out"].to(DEV))
    assert torch.allclose(v.grad.cpu(), ref["grad_value"], **TOL)


def test_transpose_grad_vs_reference():
    i, ref = G["coalesce"]["in"], G["coalesce"]["transpose"]
    v = _leaf(i["v1"])
    idx, val = ts.transpose(i["index"].to(DEV), v, i["M"], i["N"])
    assert torch.equal(idx.cpu(), ref["index"])
    val.backward(ref["grad_out"].to(DEV))
    assert torch.allclose(v.grad.cpu(), ref["grad_value"], **TOL)


def _ab():
    i = G["ops"]["in"]
    return i, _leaf(i["va"]), _leaf(i["vb"])


def test_spadd_and_add_grad_vs_reference():
    i, a, b = _ab()
    ref = G["ops"]["spadd"]
    idx, val = ts.spadd(torch.stack([i["ra"], i["ca"]]).to(DEV), a, torch.stack([i["rb"], i["cb"]]).to(DEV), b,
                        i["M"], i["N"])
    assert torch.equal(idx.cpu(), ref["index"])
    val.backward(ref["grad_out"].to(DEV))
    assert torch.allclose(a.grad.cpu(), ref["grad_a"], **TOL) and torch.allclose(b.grad.cpu(), ref["grad_b"], **TOL)

    i, a, b = _ab()
    ref = G["ops"]["add"]
    A = ts.SparseTensor(row=i["ra"].to(DEV), col=i["ca"].to(DEV), value=a, sparse_sizes=(i["M"], i["N"]))
    B = ts.SparseTensor(row=i["rb"].to(DEV), col=i["cb"].to(DEV), value=b, sparse_sizes=(i["M"], i["N"]))
    C = A + B
    assert torch.equal(C.storage.row().cpu(), ref["row"]) and torch.equal(C.storage.col().cpu(), ref["col"])
    C.storage.value().backward(ref["grad_out"].to(DEV))
    assert torch.allclose(a.grad.cpu(), ref["grad_a"], **TOL) and torch.allclose(b.grad.cpu(), ref["grad_b"], **TOL)


@pytest.mark.parametrize("tag", ["v", "v2"])
@pytest.mark.parametrize("red", ["sum", "mean", "min", "max"])
def test_to_symmetric_and_reduce_grad_vs_reference(tag, red):
    i = G["ops"]["in"]
    src = i["va"] if tag == "v" else i["va2"]

    def mk():
        v = _leaf(src)
        return v, ts.SparseTensor(row=i["ra"].to(DEV), col=i["ca"].to(DEV), value=v, sparse_sizes=(i["M"], i["N"]))

    ref = G["ops"][f"sym_{red}_{tag}"]
    v, A = mk()
    S = A.to_symmetric(red)
    assert torch.equal(S.storage.row().cpu(), ref["row"]) and torch.equal(S.storage.col().cpu(), ref["col"])
    assert torch.allclose(S.storage.value().detach().cpu(), ref["value"], **TOL)
    S.storage.value().backward(ref["grad_out"].to(DEV))
    assert torch.allclose(v.grad.cpu(), ref["grad_value"], **TOL)

    for dim in (0, 1):
        ref = G["ops"][f"reduce_{red}_{dim}_{tag}"]
        v, A = mk()
        r = getattr(ts, red)(A, dim)
        assert r.grad_fn is not None
        assert torch.allclose(r.detach().cpu(), ref["out"], **TOL)
        r.backward(ref["grad_out"].to(DEV))
        assert torch.allclose(v.grad.cpu(), ref["grad_value"], **TOL), (red, dim)


def test_gcn_normalisation_flow_vs_reference_and_dense():
    i, ref = G["ops"]["in"], G["gcn"]
    a, x = _leaf(ref["value"]), _leaf(ref["x"])
    A = ts.SparseTensor(row=i["ra"].to(DEV), col=i["ca"].to(DEV), value=a, sparse_sizes=(i["M"], i["N"]))
    deg = ts.sum(A, dim=1)
    dis = deg.pow(-0.5)
    dis = dis.masked_fill(dis == float("inf"), 0.0)
    y = ts.mul(ts.mul(A, dis.view(-1, 1)), dis.view(1, -1)) @ x
    assert torch.allclose(y.detach().cpu(), ref["y"], rtol=1e-11, atol=1e-11)
    y.backward(ref["grad_out"].to(DEV))
    assert torch.allclose(a.grad.cpu(), ref["grad_value"], rtol=1e-10, atol=1e-10)
    assert torch.allclose(x.grad.cpu(), ref["grad_x"], rtol=1e-10, atol=1e-10)
    # and against fp64 dense autograd (independent of the reference)
    a2, x2 = ref["value"].clone().requires_grad_(), ref["x"].clone().requires_grad_()
    dense = torch.zeros(i["M"], i["N"], dtype=torch.float64).index_put((i["ra"], i["ca"]), a2)
    d = dense.sum(1).pow(-0.5)
    d = d.masked_fill(d == float("inf"), 0.0)
    (d.view(-1, 1) * dense * d.view(1, -1) @ x2).backward(ref["grad_out"])
    assert torch.allclose(a.grad.cpu(), a2.grad, rtol=1e-10, atol=1e-10)
    assert torch.allclose(x.grad.cpu(), x2.grad, rtol=1e-10, atol=1e-10)


def test_ties_route_to_first_entry_and_empty_segments():
    # duplicates with equal values: the whole gradient goes to the earliest input entry (torch_scatter's
    # strict-compare arg rule); empty segments produce 0 and receive nothing
    row = torch.tensor([1, 1, 1, 0], device=DEV)
    col = torch.tensor([2, 2, 2, 0], device=DEV)
    v = torch.tensor([3.0, 5.0, 5.0, 1.0], device=DEV, requires_grad=True)
    _, _, out = ops.coalesce(row, col, v, 3, 3, "max")
    out.backward(torch.tensor([10.0, 20.0], device=DEV))
    assert v.grad.tolist() == [0.0, 20.0, 0.0, 10.0]
    ptr = torch.tensor([0, 0, 3, 3, 4], device=DEV)
    w = torch.tensor([2.0, 2.0, 7.0, 4.0], device=DEV, requires_grad=True)
    r = ops.segment_reduce(ptr, w, "min")
    assert r.tolist() == [0.0, 2.0, 0.0, 4.0]
    r.backward(torch.tensor([1.0, 2.0, 3.0, 4.0], device=DEV))
    assert w.grad.tolist() == [2.0, 0.0, 0.0, 4.0]
    w.grad = None
    ops.segment_reduce(ptr, w, "mean").backward(torch.tensor([1.0, 3.0, 5.0, 4.0], device=DEV))
    assert w.grad.tolist() == [1.0, 1.0, 1.0, 4.0]


def test_no_grad_path_untouched():
    row = torch.tensor([1, 1, 0], device=DEV)
    col = torch.tensor([2, 2, 0], device=DEV)
    v = torch.tensor([3.0, 5.0, 1.0], device=DEV)
    _, _, out = ops.coalesce(row, col, v, 3, 3, "add")
    assert out.grad_fn is None and out.tolist() == [1.0, 8.0]
    with torch.no_grad():
        _, _, out = ops.coalesce(row, col, v.clone().requires_grad_(), 3, 3, "add")
    assert out.grad_fn is None
