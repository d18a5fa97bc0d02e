"""Probe: torch.distributed._symmetric_memory on this image — peer views + copy-engine pulls over NVLink."""
import os, time, torch, torch.distributed as dist
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
import torch.distributed._symmetric_memory as symm
rows, F = 1_000_000, 128
try:
    t = symm.empty((rows, F), dtype=torch.bfloat16, device=dev)
    hdl = symm.rendezvous(t, dist.group.WORLD)
    t.fill_(float(rank + 1))
    hdl.barrier()
    full = torch.empty(world, rows, F, dtype=torch.bfloat16, device=dev)
    peers = [hdl.get_buffer(p, (rows, F), torch.bfloat16) for p in range(world)]
    streams = [torch.cuda.Stream() for _ in range(world)]
    def pull():
        ev = torch.cuda.Event(); ev.record()
        for p in range(world):
            streams[p].wait_event(ev)
            with torch.cuda.stream(streams[p]):
                full[p].copy_(peers[p], non_blocking=True)
        for s in streams:
            torch.cuda.current_stream().wait_stream(s)
    for _ in range(3): pull()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): pull()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    ok = all(float(full[p, 0, 0]) == p + 1 and float(full[p, -1, -1]) == p + 1 for p in range(world))
    nbytes = (world - 1) * rows * F * 2
    print(f"rank {rank}: symm pull ok={ok} {ms:.3f} ms  {nbytes / ms / 1e6:.0f} GB/s in from peers", flush=True)
    # overlap with a persistent compute kernel stand-in: a big matmul loop on the current stream
    a = torch.randn(8192, 8192, device=dev, dtype=torch.bfloat16)
    def both():
        ev = torch.cuda.Event(); ev.record()
        for p in range(world):
            streams[p].wait_event(ev)
            with torch.cuda.stream(streams[p]):
                full[p].copy_(peers[p], non_blocking=True)
        for _ in range(4): a @ a
        for s in streams:
            torch.cuda.current_stream().wait_stream(s)
    both(); torch.cuda.synchronize(); dist.barrier()
    e0.record(); 
    for _ in range(5): both()
    e1.record(); torch.cuda.synchronize()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    for _ in range(5):
        for _ in range(4): a @ a
    e3.record(); torch.cuda.synchronize()
    print(f"rank {rank}: copy+matmul {e0.elapsed_time(e1) / 5:.3f} ms vs matmul alone {e2.elapsed_time(e3) / 5:.3f} ms", flush=True)
except Exception as e:
    import traceback; traceback.print_exc()
    print(f"rank {rank}: symm FAILED {e!r}", flush=True)
dist.barrier(); dist.destroy_process_group()
