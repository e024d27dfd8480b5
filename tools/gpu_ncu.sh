#!/bin/bash
mkdir -p gpurun_out
python tools/bench_one.py conv256 > gpurun_out/plain1.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 1 -c 2 -f -o gpurun_out/prof_conv256 python tools/bench_one.py conv256 > gpurun_out/ncu1.log 2>&1
python tools/bench_one.py conv128 > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 1 -c 1 -f -o gpurun_out/prof_conv128 python tools/bench_one.py conv128 > gpurun_out/ncu2.log 2>&1
python tools/bench_one.py gn > gpurun_out/plain3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gn_apply -s 1 -c 1 -f -o gpurun_out/prof_gn python tools/bench_one.py gn > gpurun_out/ncu3.log 2>&1
tail -3 gpurun_out/ncu1.log gpurun_out/ncu2.log gpurun_out/ncu3.log
ls -la gpurun_out/*.ncu-rep
