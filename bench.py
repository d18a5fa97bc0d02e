#!/usr/bin/env python
"""bench.py — the hot-path benchmark (driver contract).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|...]

One "step" = one CSR SpMM (sum) pass over the synthetic matrix of BASELINE.json configs[1]:
1M x 1M (per GPU), ~16 nnz/row, dense operand F=128 bf16 (SURVEY §8d generator G2 / G5).
Rank 0 prints ONE JSON line. `value` = whole-job GFLOP/s (2*E*F per SpMM) with inputs resident in HBM;
`e2e` = the same metric through the C-ABI host-buffer call (pinned host tensors in, host tensors out,
H2D/D2H inside the timed region); `roofline` = algorithmic HBM bytes / step time vs the measured copy
peak; `cpu_baseline` = the reference's own CPU spmm (oracle/_ref) on this box's host cores.

`--impl reference` times the reference's CPU operator itself on the same config (rank 0 only).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

WORKLOADS = {
    # name: (rows per GPU, avg nnz/row, F, dtype, reduce, generator)
    "c2": dict(M=1_000_000, deg=16, F=128, dtype="bf16", reduce="sum", gen="uniform",
               desc="SpMM_sum CSR 1Mx1M, avg 16 nnz/row, F=128 bf16 (BASELINE configs[1])"),
    "c2_f32": dict(M=1_000_000, deg=16, F=32, dtype="bf16", reduce="sum", gen="uniform",
                   desc="SpMM_sum CSR 1Mx1M, avg 16 nnz/row, F=32 bf16"),
    "c2_f256": dict(M=1_000_000, deg=16, F=256, dtype="bf16", reduce="sum", gen="uniform",
                    desc="SpMM_sum CSR 1Mx1M, avg 16 nnz/row, F=256 bf16"),
    "c2_fp32": dict(M=1_000_000, deg=16, F=128, dtype="f32", reduce="sum", gen="uniform",
                    desc="SpMM_sum CSR 1Mx1M, avg 16 nnz/row, F=128 fp32"),
    "c1": dict(M=10_000, deg=5, F=32, dtype="f32", reduce="sum", gen="uniform",
               desc="spmm_sum random COO 10kx10k, 50k nnz, F=32 fp32 (BASELINE configs[0])"),
    "c3": dict(M=500_000, deg=16, F=256, dtype="f32", reduce="max", gen="powerlaw",
               desc="SpMM_max CSR 500kx500k power-law degree, F=256 fp32 (BASELINE configs[2], forward)"),
}


def torch_dtype(name):
    import torch
    return {"bf16": torch.bfloat16, "f32": torch.float32, "f16": torch.float16}[name]


# ---------------------------------------------------------------------------------------------------
# synthetic inputs (identical on both arms: generated on the CPU from fixed seeds)
# ---------------------------------------------------------------------------------------------------
def gen_matrix(w, rank, world):
    """Row block `rank` of the (world*M) x (world*M) matrix: returns rowptr, col, value(float32), Ncols."""
    import torch
    M, N = w["M"], w["M"] * world
    g = torch.Generator().manual_seed(1 if world == 1 else 10 + rank)
    if w["gen"] == "uniform":
        E0 = w["deg"] * M
        row = torch.randint(M, (E0,), generator=g)
        col = torch.randint(N, (E0,), generator=g)
        key = torch.unique(row * N + col)
        row, col = key // N, key % N
    else:  # power-law degrees: deg_i = min(floor(d0 * u^(-1/alpha)), N/10), alpha = 1.5
        u = torch.rand(M, generator=g).clamp_(min=1e-9)
        deg = (w["deg"] / 3.0 * u.pow(-1.0 / 1.5)).floor().long().clamp_(max=N // 10)
        deg[torch.rand(M, generator=g) < 0.02] = 0          # some empty rows by design
        row = torch.repeat_interleave(torch.arange(M), deg)
        col = torch.randint(N, (row.numel(),), generator=g)
        key = torch.unique(row * N + col)
        row, col = key // N, key % N
    rowptr = torch.zeros(M + 1, dtype=torch.long)
    rowptr[1:] = torch.cumsum(torch.bincount(row, minlength=M), 0)
    value = torch.rand(col.numel(), generator=g) + 0.5     # U(0.5, 1.5)
    return rowptr, col, value, N


def gen_dense(w, rank, rows):
    import torch
    g = torch.Generator().manual_seed(1000 + rank)
    return torch.randn(rows, w["F"], generator=g)


def algorithmic_bytes(M, N, E, F, s, arg):
    """compulsory model, SURVEY §8d: 8(M+1) + 8E + sE + s*N*F + s*M*F (+ 8*M*F arg_out)."""
    return 8 * (M + 1) + 8 * E + s * E + s * N * F + s * M * F + (8 * M * F if arg else 0)


# ---------------------------------------------------------------------------------------------------
# clocks sampling (NVML, in-process thread)
# ---------------------------------------------------------------------------------------------------
class ClockSampler:
    def __init__(self, index):
        self.samples, self.reasons, self.power = [], set(), []
        self.stop_flag = False
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.max_mhz = None

    def _loop(self):
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
                self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
            except Exception:
                pass
            time.sleep(0.002)

    def start(self):
        if self.ok:
            self.t = threading.Thread(target=self._loop, daemon=True)
            self.t.start()

    def stop(self):
        self.stop_flag = True
        if self.ok:
            self.t.join(timeout=1)
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s),
                "power_w_max": max(self.power) if self.power else None}


# ---------------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU operator (oracle/_ref), all host threads
# ---------------------------------------------------------------------------------------------------
def run_reference(args, w):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    torch.set_num_threads(os.cpu_count() or 1)
    cores = torch.get_num_threads()
    from oracle import build_ref
    kind = "reference"
    if build_ref.available():
        build_ref.load()

        def cpu_spmm(rowptr, col, value, x):
            if w["reduce"] == "max":
                return torch.ops.torch_sparse.spmm_max(rowptr, col, value, x)[0]
            return torch.ops.torch_sparse.spmm_sum(None, rowptr, col, value, None, None, x)
    else:  # the reference could not be compiled here: time the oracle port instead
        import oracle
        kind = "port"
        os.environ.setdefault("OMP_NUM_THREADS", str(cores))

        def cpu_spmm(rowptr, col, value, x):
            return oracle.spmm(rowptr, col, value, x, w["reduce"])[0]

    dt = torch_dtype(w["dtype"])
    rowptr, col, value, N = gen_matrix(w, 0, 1)
    x = gen_dense(w, 0, N).to(dt)
    value = value.to(dt)
    M, F = w["M"], w["F"]

    def sample(rows):
        rp = rowptr[:rows + 1]
        e = int(rp[-1])
        return rp, col[:e], value[:e]

    # calibrate a bounded sample: whole run (warmup + steps) within ~150 s of CPU time
    rp, c, v = sample(M)
    t0 = time.perf_counter()
    cpu_spmm(rp, c, v, x)
    t_full = time.perf_counter() - t0
    total = args.steps + args.warmup
    frac = min(1.0, 150.0 / max(t_full * total, 1e-9))
    rows = M if frac >= 1.0 else max(1024, int(M * frac))
    rp, c, v = sample(rows)
    E = c.numel()
    for _ in range(args.warmup):
        cpu_spmm(rp, c, v, x)
    times = []
    for _ in range(args.steps):
        t0 = time.perf_counter()
        cpu_spmm(rp, c, v, x)
        times.append(time.perf_counter() - t0)
    t = sum(times) / len(times)
    gflops = 2.0 * E * F / t / 1e9
    sample_desc = f"rows [0,{rows}) of the {M}-row workload ({E} nnz), all {N} dense rows, mean of {args.steps} steps"
    line = {
        "impl": "reference", "metric": "spmm_gflops", "value": gflops, "unit": "GFLOP/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": w["dtype"], "data": "synthetic",
        "config": {"workload": w["desc"], "rows": rows, "nnz": E, "F": F, "reduce": w["reduce"],
                   "device": "host CPU"},
        "cpu_baseline": {"value": gflops, "unit": "GFLOP/s", "cores": cores, "kind": kind, "sample": sample_desc,
                         "best_ms": min(times) * 1e3},
        "e2e": {"value": gflops, "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "hbm_gbs_equiv": algorithmic_bytes(rows, N, E, F, x.element_size(), w["reduce"] == "max") / t / 1e9,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------
def run_ours(args, w):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        args.gpus = world
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    import pytorch_sparse_b200 as ts
    from pytorch_sparse_b200 import ops
    from pytorch_sparse_b200.parallel import RowShardedSpMM

    dt = torch_dtype(w["dtype"])
    M, F, reduce = w["M"], w["F"], w["reduce"]
    rowptr_h, col_h, value_h, N = gen_matrix(w, rank, world)
    value_h = value_h.to(dt)
    x_local_h = gen_dense(w, rank, M).to(dt)   # this rank's row block of the dense operand
    E = col_h.numel()
    s = x_local_h.element_size()

    a_local = ts.SparseTensor(rowptr=rowptr_h.to(dev), col=col_h.to(dev), value=value_h.to(dev),
                              sparse_sizes=(M, N), is_sorted=True, trust_data=True)
    sharded = RowShardedSpMM(a_local, reduce=reduce)
    # dense operand made resident on every GPU ONCE over NVLink (north_star: "broadcast once")
    x_local = x_local_h.to(dev)
    if world > 1:  # NCCL communicator setup is not part of the gather time
        dist.all_reduce(torch.zeros(1, device=dev))
        sharded.gather_dense(x_local)
    torch.cuda.synchronize()
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    x_full = sharded.gather_dense(x_local)
    t1.record(); torch.cuda.synchronize()
    gather_ms = t0.elapsed_time(t1)

    def step():
        return sharded.local_spmm(x_full)

    for _ in range(max(args.warmup, 3)):
        out = step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    torch.cuda.synchronize()
    ev0 = torch.cuda.Event(enable_timing=True); ev1 = torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        out = step()
    ev1.record()
    torch.cuda.synchronize()
    clocks = sampler.stop()
    if world > 1:
        dist.barrier()
    ms_total = ev0.elapsed_time(ev1)
    tmax = torch.tensor([ms_total], device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms_step = tmax.item() / args.steps

    # ---- end to end through the C-ABI host-buffer call (pinned host in, host out) ------------------
    pin = lambda t: t.pin_memory()
    rp_p, col_p, val_p = pin(rowptr_h), pin(col_h), pin(value_h)
    x_full_p = pin(x_full.cpu())
    e2e_steps = max(3, min(args.steps, 10))
    # warm-up: device staging buffers + BOTH pinned output buffers the steady-state loop alternates between
    w1 = ops.spmm_fw_host(rp_p, col_p, val_p, x_full_p, reduce)
    w2 = ops.spmm_fw_host(rp_p, col_p, val_p, x_full_p, reduce)
    del w1, w2
    ops.spmm_fw_host(rp_p, col_p, val_p, x_full_p, reduce)
    if world > 1:
        dist.barrier()
    te = time.perf_counter()
    for _ in range(e2e_steps):
        out_h, _ = ops.spmm_fw_host(rp_p, col_p, val_p, x_full_p, reduce)
    e2e_ms = (time.perf_counter() - te) * 1e3 / e2e_steps
    te_t = torch.tensor([e2e_ms], device=dev)
    if world > 1:
        dist.all_reduce(te_t, op=dist.ReduceOp.MAX)
    e2e_ms = te_t.item()
    arg = reduce in ("min", "max")
    h2d = 8 * (M + 1) + 8 * E + s * E + s * N * F
    d2h = s * M * F + (8 * M * F if arg else 0)

    # parity spot check of the timed result vs the end-to-end result (same kernel, two paths)
    assert torch.equal(out_h, out.cpu()), "device-resident and host-buffer results differ"

    flops = 2.0 * E * F
    tot = torch.tensor([flops], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tot)
    gflops = tot.item() / (ms_step * 1e-3) / 1e9
    e2e_gflops = tot.item() / (e2e_ms * 1e-3) / 1e9

    if rank == 0:
        peaks = {}
        pk = ROOT / "MEASURED_PEAKS.json"
        if pk.exists():
            peaks = json.loads(pk.read_text())
        peak = float(peaks.get("hbm_gbs", 6650.0))
        abytes = algorithmic_bytes(M, N, E, F, s, arg)
        achieved = abytes / (ms_step * 1e-3) / 1e9
        traffic = None
        tf = ROOT / "profiles" / "ncu_traffic.json"
        if tf.exists() and world == 1:  # ncu capture of the N=1 launch (profiles/r01_ncu_spmm_c2.md)
            traffic = json.loads(tf.read_text()).get(args.workload)
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            cpu = cpu_baseline_subprocess(args)
        line = {
            "metric": "spmm_gflops", "value": gflops, "unit": "GFLOP/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": w["dtype"], "data": "synthetic",
            "config": {"workload": w["desc"], "rows_per_gpu": M, "cols": N, "nnz_per_gpu": E, "F": F,
                       "reduce": reduce, "parallelism": f"row-block x{world}; dense operand all-gathered once "
                       f"over NCCL ({gather_ms:.2f} ms, not in the step)" if world > 1 else "single GPU",
                       "l2": "inputs larger than L2 (dense operand %d MB + indices %d MB vs 126 MB L2); no flush"
                             % (s * N * F >> 20, (8 * E + s * E) >> 20),
                       "accumulate": "fp32"},
            "hbm_gbs": achieved, "gather_counted_gbs": (abytes - s * N * F + s * E * F) / (ms_step * 1e-3) / 1e9,
            "clocks": clocks,
            "e2e": {"value": e2e_gflops, "unit": "GFLOP/s", "ms_per_step": e2e_ms, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "path": "tsb200_spmm_fw_host (pinned host buffers)"},
            "gpu_launches": 3 * args.steps,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": "measured" if peaks else "fallback",
                         "algorithmic_bytes": abytes},
            "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def cpu_baseline_subprocess(args):
    """Time the reference's CPU operator in a clean process (its torch_sparse op namespace must not
    collide with ours); bounded sample chosen by run_reference()."""
    cmd = [sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "3", "--warmup", "1",
           "--workload", args.workload]
    env = dict(os.environ)
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE"):
        env.pop(k, None)
    try:
        res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
        line = [l for l in res.stdout.splitlines() if l.startswith("{")][-1]
        return json.loads(line)["cpu_baseline"]
    except Exception as e:  # pragma: no cover
        return {"value": None, "error": repr(e)[:200]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    w = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, w)
    else:
        run_ours(args, w)


if __name__ == "__main__":
    main()
