"""Sweep the feature-slice width of the fused min/max backward at BASELINE configs[2] scale
(each configuration in a fresh process: the knobs are read from the environment by libtsb200)."""
import json, os, subprocess, sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
CHILD = r'''
import json, os, sys
sys.path.insert(0, %r); sys.path.insert(0, %r + "/tests")
import torch, bench
import pytorch_sparse_b200 as ts
from pytorch_sparse_b200 import ops
dev = "cuda:0"
w = bench.WORKLOADS["c3"]
rowptr, col, value, N = bench.gen_matrix(w, 0, 1)
M, F = w["M"], w["F"]
rowptr, col, value = rowptr.to(dev), col.to(dev), value.to(dev)
x = torch.randn(N, F, device=dev); go = torch.randn(M, F, device=dev)
out, arg = ops.spmm_fw(rowptr, col, value, x, "max")
def t(fn, steps=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps
both = t(lambda: ops.spmm_minmax_bw(col, value, x, go, arg, True, True))
only_mat = t(lambda: ops.spmm_minmax_bw(col, value, x, go, arg, False, True))
only_val = t(lambda: ops.spmm_minmax_bw(col, value, x, go, arg, True, False))
zero = t(lambda: (torch.zeros(col.numel(), device=dev), torch.zeros_like(x)))
print(json.dumps(dict(lsl=os.environ.get("TSB200_MMBW_LSL", "auto"), split=os.environ.get("TSB200_MMBW_SPLIT", "auto"),
                      both_ms=both, only_mat_ms=only_mat, only_val_ms=only_val, zero_fill_ms=zero)), flush=True)
''' % (str(ROOT), str(ROOT))

configs = [("auto", "auto")] + [(str(l), "0") for l in (8, 7, 6, 5, 4)]
for lsl, split in configs:
    env = dict(os.environ)
    if lsl != "auto":
        env["TSB200_MMBW_LSL"] = lsl
    r = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True, timeout=300)
    print(r.stdout.strip().splitlines()[-1] if r.stdout.strip() else "ERR " + r.stderr[-400:], flush=True)
