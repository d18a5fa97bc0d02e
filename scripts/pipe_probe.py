"""Multi-GPU probe: gather of the dense operand (NCCL all_gather vs copy-engine peer pulls) and the pipelined
including-gather SpMM step for a few pipeline settings. torchrun; prints one line per setting on rank 0."""
import os, sys, json
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch, torch.distributed as dist
import bench
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
import pytorch_sparse_b200 as ts
from pytorch_sparse_b200.parallel import PipelinedRowShardedSpMM, RowShardedSpMM
w = bench.WORKLOADS["c2"]; M, F = w["M"], w["F"]
rowptr, col, value, N = bench.gen_matrix(w, rank, world)
a = ts.SparseTensor(rowptr=rowptr.to(dev), col=col.to(dev), value=value.bfloat16().to(dev), sparse_sizes=(M, N), is_sorted=True, trust_data=True)
x = bench.gen_dense(w, rank, M).bfloat16().to(dev)
sh = RowShardedSpMM(a, "sum")
dist.all_reduce(torch.zeros(1, device=dev))
def mx(v):
    t = torch.tensor([v], device=dev, dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX); return t.item()
def T(fn, steps=10, warm=2):
    dist.barrier(); return mx(bench._time_cuda(fn, steps, warm))
xf = sh.gather_dense(x)
res = {"world": world, "nccl_gather_ms": T(lambda: sh.gather_dense(x)), "steady_ms": T(lambda: sh.local_spmm(xf), 20, 3),
       "serial_ms": T(lambda: sh.local_spmm(sh.gather_dense(x)))}
ref = sh.local_spmm(xf)
if rank == 0: print(json.dumps(res), flush=True)
settings = [(t, s, c) for t in ("peer", "nccl") for s in ((1, 2, 7) if t == "peer" else (0,)) for c in (1, 2, 4)]
if os.environ.get("PROBE_SHORT"):
    settings = [("peer", 1, 1), ("peer", 2, 1), ("peer", 7, 1), ("peer", 2, 2), ("peer", 1, 2), ("nccl", 0, 2)]
for transport, streams, chunks in settings:
    os.environ["TSB200_PEER_STREAMS"] = str(streams)
    pipe = PipelinedRowShardedSpMM(a, block=M, chunks=chunks, transport=transport)
    xs = pipe.to_sliced(x)
    if transport == "peer":
        for _ in range(2):
            pipe.input_buffer(xs, xs.size(-1)).copy_(xs); pipe._step += 1
        pipe._step = 0
        run = lambda: pipe.forward_sliced(pipe.input_buffer(xs, xs.size(-1)))
    else:
        run = lambda: pipe.forward_sliced(xs)
    ms = T(run, 6, 2)
    out = pipe.from_sliced(run())
    ok = bool(torch.allclose(out.float(), ref.float(), rtol=2e-2, atol=2e-2))
    okt = torch.tensor([1.0 if ok else 0.0], device=dev); dist.all_reduce(okt, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(json.dumps({"transport": transport, "peer_streams": streams, "chunks": chunks, "incl_gather_ms": ms, "matches_steady_state": bool(okt.item() > 0.5)}), flush=True)
    del pipe
dist.barrier(); dist.destroy_process_group()
