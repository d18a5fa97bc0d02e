#!/bin/bash
# round-2 final evidence batch (1 GPU): the driver bench line, ncu launch list + full captures exported to csv on the box
# (the .ncu-rep files are too large to travel back: only the raw-page csv of each capture is kept)
mkdir -p gpurun_out
python bench.py --steps 100 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r02_launches_c2.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-secondary > /dev/null 2> gpurun_out/ncu_launches.err
cap() {  # name, kernel regex, skip, count, command...
  local name=$1 k=$2 s=$3 c=$4; shift 4
  ncu --set full --clock-control none -k regex:$k -s $s -c $c -f -o /tmp/$name "$@" > /dev/null 2> gpurun_out/ncu_$name.err
  ncu -i /tmp/$name.ncu-rep --page raw --csv > gpurun_out/${name}_raw.csv 2>> gpurun_out/ncu_$name.err
  rm -f /tmp/$name.ncu-rep
}
cap r02_spmm_c2 spmm_vec_kernel 4 1 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-secondary
TSB200_PIN_MB=64 cap r02_spmm_c2_pin64 spmm_vec_kernel 4 1 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-secondary
cap r02_sddmm_c2 sddmm_vec_kernel 2 1 python scripts/bench_extra.py c2bw
cap r02_spspmm_c4 spspmm_kernel 2 2 python scripts/bench_extra.py c4
TSB200_SPSPMM=fused cap r02_spspmm_c4_fused spspmm_kernel 1 1 python scripts/bench_extra.py c4
tail -c 400 gpurun_out/bench_n1.err; ls -la gpurun_out/; du -sh gpurun_out
