"""One un-timed invocation of a secondary op between cudaProfilerStart/Stop, for
`ncu --profile-from-start off --metrics gpu__time_duration.sum --csv` launch lists (profiles/*_launches_*.csv).

    python scripts/launch_list.py coalesce|csr2csc|spspmm|c3bwd
"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tests"))
import torch
import pytorch_sparse_b200 as ts
from pytorch_sparse_b200 import ops
from util import fast_random_csr

dev = "cuda:0"
what = sys.argv[1]
if what == "coalesce":
    M = N = 262_144
    g = torch.Generator(device=dev).manual_seed(5)
    E0 = 4_194_304
    row = torch.randint(M, (E0,), generator=g, device=dev); col = torch.randint(N, (E0,), generator=g, device=dev)
    perm = torch.randperm(2 * E0, generator=g, device=dev)
    row2, col2 = torch.cat([row, row])[perm], torch.cat([col, col])[perm]
    val = torch.randn(2 * E0, device=dev)
    fn = lambda: ops.coalesce(row2, col2, val, M, N, "add")
elif what == "csr2csc":
    M = 1_000_000
    row, rowptr, col = fast_random_csr(M, M, 16, 1, dev)
    fn = lambda: ops.csr2csc(row, col, M, M, want_colptr=True, want_row_csc=True)
elif what == "spspmm":
    M = 262_144
    _, rpa, ca = fast_random_csr(M, M, 32, 3, dev)
    _, rpb, cb = fast_random_csr(M, M, 32, 4, dev)
    va = torch.randn(ca.numel(), device=dev); vb = torch.randn(cb.numel(), device=dev)
    fn = lambda: ops.spspmm(rpa, ca, va, rpb, cb, vb, M, M, M, True)
else:
    raise SystemExit("unknown op")
for _ in range(2):
    fn()
torch.cuda.synchronize()
torch.cuda.profiler.start()
fn()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
