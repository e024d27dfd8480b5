#!/bin/bash
# GPU session: gpu test suite + bench with the per-layer CUDA-event dump (no ncu).
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -8 | tee gpurun_out/tests_gpu.log
HYVAE_PROFILE_DUMP=gpurun_out/profile_dump.csv timeout 600 python bench.py --steps ${STEPS:-2} --warmup ${WARMUP:-3} 2>&1 | tail -3 | tee gpurun_out/bench.log
python tools/summarize_profile.py gpurun_out/profile_dump.csv 60 > gpurun_out/profile_summary.txt
python tools/profile_families.py gpurun_out/profile_dump.csv > gpurun_out/profile_families.txt 2>&1
head -30 gpurun_out/profile_families.txt
