#!/bin/bash
# ncu --set full captures of the top kernels (one launch each, after two warm-up launches).  usage: gpu_ncu.sh case:kernel_regex ...
mkdir -p gpurun_out
for spec in "$@"; do
  c=${spec%%:*}; k=${spec##*:}
  python tools/bench_one.py $c > gpurun_out/plain_$c.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f -o gpurun_out/prof_$c python tools/bench_one.py $c > gpurun_out/ncu_$c.log 2>&1
  tail -2 gpurun_out/ncu_$c.log
done
ls -la gpurun_out/*.ncu-rep
