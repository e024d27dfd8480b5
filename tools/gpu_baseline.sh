#!/bin/bash
# GPU session: full gpu test suite, bench, and the ncu per-launch time list of one bench step.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -8 | tee gpurun_out/tests_gpu.log
HYVAE_PROFILE_DUMP=gpurun_out/profile_dump.csv timeout 600 python bench.py --steps 2 --warmup 3 2>&1 | tail -3 | tee gpurun_out/bench.log
python tools/summarize_profile.py gpurun_out/profile_dump.csv 60 > gpurun_out/profile_summary.txt
timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/plain_ncu_cmd.log 2>&1 && \
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 80000 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launches.log 2>&1
tail -2 gpurun_out/ncu_launches.log
wc -l gpurun_out/launches.csv
