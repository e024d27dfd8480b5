#!/bin/bash
# ncu --set full of named bench_one cases: usage gpu_r2_ncu2.sh case:kernel_regex ...
mkdir -p gpurun_out
bash tools/gpu_ncu.sh "$@"
for spec in "$@"; do
  c=${spec%%:*}
  ncu -i gpurun_out/prof_$c.ncu-rep --page details > gpurun_out/ncu_${c}_details.txt 2>&1
  ncu -i gpurun_out/prof_$c.ncu-rep --page raw --csv > gpurun_out/ncu_${c}_raw.csv 2>&1
  ncu -i gpurun_out/prof_$c.ncu-rep --page source --csv --print-source sass > gpurun_out/ncu_${c}_sass.csv 2>&1
done
ls -la gpurun_out/ | head -40
