"""Planned vs unplanned SpMM forward at C1 / C2 / C3 (development helper): tsb200_spmm_plan once, then
tsb200_spmm_fw_planned = one memset + one kernel."""
import sys, json
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import bench
from pytorch_sparse_b200 import ops
dev = "cuda:0"


def T(fn, steps, warm=5):
    return bench._time_cuda(fn, steps, warm)


for name, steps in (("c1", 300), ("c2", 50), ("c3", 20), ("c2_f32", 50)):
    w = bench.WORKLOADS[name]
    rowptr, col, value, N = bench.gen_matrix(w, 0, 1)
    dt = bench.torch_dtype(w["dtype"])
    x = bench.gen_dense(w, 0, N).to(dt).to(dev)
    rp, c, v = rowptr.to(dev), col.to(dev), value.to(dt).to(dev)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); plan = ops.spmm_plan(rp, c.numel()); e1.record(); torch.cuda.synchronize()
    plan_ms = e0.elapsed_time(e1)
    red = w["reduce"]
    ms_u = T(lambda: ops.spmm_fw(rp, c, v, x, red), steps)
    ms_p = T(lambda: ops.spmm_fw(rp, c, v, x, red, plan=plan), steps)
    ou, au = ops.spmm_fw(rp, c, v, x, red)
    op, ap = ops.spmm_fw(rp, c, v, x, red, plan=plan)
    same = bool(torch.allclose(ou.float(), op.float(), rtol=1e-2, atol=1e-2)) and (au is None or bool(torch.equal(au, ap)))
    print(json.dumps({"workload": name, "unplanned_ms": ms_u, "planned_ms": ms_p, "plan_build_ms": plan_ms,
                      "segments": plan.n_seg, "multi_segment_rows": plan.n_long, "same_result": same}), flush=True)
