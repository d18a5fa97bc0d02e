#!/bin/bash
# multi-GPU bench lines (steady state, including-gather serial and pipelined) for N = $1
N=${1:-2}
mkdir -p gpurun_out
for C in ${CHUNKS:-2 4}; do
  TSB200_PIPE_CHUNKS=$C python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 30 --warmup 5 > gpurun_out/bench_n${N}_c${C}.json 2> gpurun_out/bench_n${N}_c${C}.err
  tail -c 600 gpurun_out/bench_n${N}_c${C}.err
  python - <<EOF
import json
try:
    l=[x for x in open("gpurun_out/bench_n${N}_c${C}.json") if x.startswith("{")][-1]
    d=json.loads(l)
    print("N=$N chunks=$C", "steady ms", round(d["ms_per_step"],4), d.get("multi_gpu"), d.get("parity"))
except Exception as e:
    print("no line", e)
EOF
done
