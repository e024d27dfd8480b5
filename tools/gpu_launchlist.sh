#!/bin/bash
# ncu per-launch device times of the bench command (first N launches: cold-cache, serialised -> compare SHARES)
mkdir -p gpurun_out
timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/plain_ncu_cmd.log 2>&1 && \
timeout 1300 ncu --metrics gpu__time_duration.sum --clock-control none -c ${NLAUNCH:-9000} --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launches.log 2>&1
tail -2 gpurun_out/ncu_launches.log; wc -l gpurun_out/launches.csv; tail -1 gpurun_out/plain_ncu_cmd.log | cut -c1-200
