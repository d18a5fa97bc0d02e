#!/bin/bash
# round-2 GPU experiment batch A (1 GPU): SpSpMM modes, L2 policy / panel sweep, SDDMM predicated variant, bench line
mkdir -p gpurun_out
python scripts/bench_extra.py c4 > gpurun_out/c4.log 2>&1
python scripts/sweep_l2_policy.py > gpurun_out/l2_sweep.log 2>&1
python scripts/sweep_l2_policy.py --F 256 --pins 0,48,64,96 --panels 2 >> gpurun_out/l2_sweep.log 2>&1
python scripts/sweep_l2_policy.py --F 32 --pins 0,32 --panels "" >> gpurun_out/l2_sweep.log 2>&1
python scripts/bench_extra.py c2bw > gpurun_out/c2bw_default.log 2>&1
TSB200_LIB=$PWD/pytorch_sparse_b200/libtsb200_sddmm_pred.so python scripts/bench_extra.py c2bw > gpurun_out/c2bw_pred.log 2>&1
TSB200_LIB=$PWD/pytorch_sparse_b200/libtsb200_sddmm_pred.so python -m pytest tests/test_spmm_gpu.py tests/test_golden_gpu.py -q -x 2>&1 | tail -5 > gpurun_out/pred_tests.log
python bench.py --steps 50 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
tail -c 3000 gpurun_out/c4.log; cat gpurun_out/l2_sweep.log; cat gpurun_out/c2bw_default.log gpurun_out/c2bw_pred.log gpurun_out/pred_tests.log; tail -c 1500 gpurun_out/bench_n1.err; head -c 6000 gpurun_out/bench_n1.json
