#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_spmm_gpu.py tests/test_golden_gpu.py tests/test_edge_cases_gpu.py tests/test_api_gpu.py tests/test_grads_gpu.py tests/test_next_rows_gpu.py -x -q 2>&1 | tail -3 > gpurun_out/plan2_tests.log
python bench.py --steps 50 --warmup 5 --no-secondary > gpurun_out/bench_n1_planned.json 2> gpurun_out/bench_n1_planned.err
ncu --set full --clock-control none -k regex:spmm_vec_kernel -s 6 -c 1 -f -o /tmp/pl python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-secondary > /dev/null 2> gpurun_out/ncu_pl.err
ncu -i /tmp/pl.ncu-rep --page raw --csv > gpurun_out/r02_spmm_c2_planned_raw.csv 2>> gpurun_out/ncu_pl.err; rm -f /tmp/pl.ncu-rep
cat gpurun_out/plan2_tests.log; tail -c 300 gpurun_out/bench_n1_planned.err; head -c 900 gpurun_out/bench_n1_planned.json
