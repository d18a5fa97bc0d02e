#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/plan_tests.log
timeout 300 python scripts/time_planned.py > gpurun_out/planned.log 2>&1
cap() { local name=$1 k=$2 s=$3 c=$4; shift 4
  ncu --set full --clock-control none -k regex:$k -s $s -c $c -f -o /tmp/$name "$@" > /dev/null 2> gpurun_out/ncu_$name.err
  ncu -i /tmp/$name.ncu-rep --page raw --csv > gpurun_out/${name}_raw.csv 2>> gpurun_out/ncu_$name.err; rm -f /tmp/$name.ncu-rep; }
cap r02_spmm_c2 spmm_vec_kernel 4 1 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-secondary
cat gpurun_out/plan_tests.log gpurun_out/planned.log
