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
ap.add_argument("--F", type=int, default=128)
ap.add_argument("--dtype", default="bf16")
ap.add_argument("--reduce", default="sum")
ap.add_argument("--steps", type=int, default=30)
a = ap.parse_args()
dev = "cuda:0"
dt = {"bf16": torch.bfloat16, "f32": torch.float32, "f16": torch.float16}[a.dtype]
row, rowptr, col = fast_random_csr(a.M, a.M, a.deg, 1, dev)
g = torch.Generator(device=dev).manual_seed(2)
val = (torch.rand(col.numel(), generator=g, device=dev) + 0.5).to(dt)
x = torch.randn(a.M, a.F, generator=g, device=dev).to(dt)
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
print(f"variant={os.environ.get('TSB200_SPMM_VARIANT','0')} M={a.M} F={a.F} {a.dtype} {a.reduce}: {ms:.4f} ms  "
      f"{2*E*a.F/ms/1e6:.0f} GFLOP/s  gather {(E*a.F*x.element_size())/ms/1e6:.0f} GB/s")
