"""SpSpMM single-pass kernel at C4: sweep of the schedule slot at which a CTA takes the NEXT row's ticket
(TSB200_SPSPMM_CLAIM) with look-back statistics (development helper)."""
import os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tests"))
import torch
from pytorch_sparse_b200._lib import lib
from pytorch_sparse_b200.ops import _p, _stream, _workspace
from util import fast_random_csr

dev = "cuda:0"
M = 262_144
_, rpa, ca = fast_random_csr(M, M, 32, 3, dev)
_, rpb, cb = fast_random_csr(M, M, 32, 4, dev)
va = torch.randn(ca.numel(), device=dev); vb = torch.randn(cb.numel(), device=dev)
products = int(((rpb[1:] - rpb[:-1])[ca]).sum())
nws = lib.tsb200_spspmm_workspace_bytes(M, M, M, ca.numel(), cb.numel())
ws = _workspace(nws, torch.device(dev)); st = _stream(torch.device(dev))
rp = torch.empty(M + 1, dtype=torch.long, device=dev)
r2 = torch.empty(products, dtype=torch.long, device=dev); c2 = torch.empty_like(r2); v2 = torch.empty(products, device=dev)


def run():
    return lib.tsb200_spspmm_fused(_p(rpa), _p(ca), _p(va), _p(rpb), _p(cb), _p(vb), M, M, M, ca.numel(), cb.numel(),
                                   _p(rp), _p(r2), _p(c2), _p(v2), products, 0, _p(ws), nws, None, st)


def timeit(fn, steps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


ref = None
for slot in (sys.argv[1:] or ["0", "1", "2", "3", "4", "5", "6"]):
    os.environ["TSB200_SPSPMM_CLAIM"] = slot
    os.environ.pop("TSB200_SPSPMM_DEBUG", None)
    ms = timeit(run)
    os.environ["TSB200_SPSPMM_DEBUG"] = "1"
    run(); torch.cuda.synchronize()
    steps, polls, n = ws[64:88].view(torch.int64).tolist()
    nnz = int(rp[-1])
    if ref is None:
        ref = (rp.clone(), c2[:nnz].clone())
    assert torch.equal(rp, ref[0]) and torch.equal(c2[:nnz], ref[1])
    print(f"claim_slot={slot}: fused {ms:.3f} ms  look-backs {n}  steps/look-back {steps / max(n, 1):.2f}  "
          f"empty polls/look-back {polls / max(n, 1):.1f}", flush=True)
