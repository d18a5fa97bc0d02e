"""Quick device-resident timing of one SpMM workload (development helper, not the driver bench)."""
import argparse, os, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tests"))
import torch
from pytorch_sparse_b200 import ops
from util import fast_random_csr

ap = argparse.ArgumentParser()
ap.add_argument("--M", type=int, default=1_000_000)
ap.add_argument("--deg", type=int, default=16)
ap.add_argument("--N", type=int, default=0, help="columns (default: = M)")
ap.add_argument("--band", type=int, default=0, help="if > 0: columns drawn within +-band of the diagonal (locality)")
ap.add_argument("--F", type=int, default=128)
ap.add_argument("--dtype", default="bf16")
ap.add_argument("--reduce", default="sum")
ap.add_argument("--steps", type=int, default=30)
a = ap.parse_args()
dev = "cuda:0"
dt = {"bf16": torch.bfloat16, "f32": torch.float32, "f16": torch.float16}[a.dtype]
N = a.N or a.M
if a.band:
    g0 = torch.Generator(device=dev).manual_seed(1)
    row = torch.arange(a.M, device=dev).repeat_interleave(a.deg)
    col = (row * N // a.M + torch.randint(-a.band, a.band + 1, (row.numel(),), generator=g0, device=dev)).clamp_(0, N - 1)
    key = torch.unique(row * N + col)
    row, col = key // N, key % N
    rowptr = torch.zeros(a.M + 1, dtype=torch.long, device=dev)
    rowptr[1:] = torch.cumsum(torch.bincount(row, minlength=a.M), 0)
else:
    row, rowptr, col = fast_random_csr(a.M, N, a.deg, 1, dev)
g = torch.Generator(device=dev).manual_seed(2)
val = (torch.rand(col.numel(), generator=g, device=dev) + 0.5).to(dt)
x = torch.randn(N, a.F, generator=g, device=dev).to(dt)
for _ in range(5):
    ops.spmm_fw(rowptr, col, val, x, a.reduce)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.steps):
    ops.spmm_fw(rowptr, col, val, x, a.reduce)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.steps
E = col.numel()
s_ = x.element_size()
alg = 8 * (a.M + 1) + 8 * E + s_ * E + s_ * N * a.F + s_ * a.M * a.F + (8 * a.M * a.F if a.reduce in ("min", "max") else 0)
print(f"M={a.M} N={N} band={a.band} F={a.F} {a.dtype} {a.reduce}: {ms:.4f} ms  {2*E*a.F/ms/1e6:.0f} GFLOP/s  "
      f"algorithmic {alg/1e6:.0f} MB -> {alg/ms/1e6:.0f} GB/s (frac {alg/ms/1e6/6572.2:.3f})  gather-counted {(E*a.F*s_)/ms/1e6:.0f} GB/s")
