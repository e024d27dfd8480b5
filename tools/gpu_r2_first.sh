#!/bin/bash
# Round-2 first GPU session: existing suite, the production-shape parity suite (-s: prints the measured errors), smoke,
# bench with the PyTorch/cuDNN leg.  No -x on the production suite: every case should report.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/nproc.txt; free -g >> gpurun_out/nproc.txt
timeout 900 python -m pytest tests -q -m gpu --ignore=tests/test_gpu_production_shapes.py 2>&1 | tail -25 | tee gpurun_out/tests_gpu.log
timeout 1500 python -m pytest tests/test_gpu_production_shapes.py -q -m gpu -s --durations=0 2>&1 | tail -60 | tee gpurun_out/tests_production.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/smoke.log
HYVAE_PROFILE_DUMP=gpurun_out/profile_dump.csv timeout 1200 python bench.py --steps ${STEPS:-2} --warmup ${WARMUP:-3} > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -c 6000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
python tools/profile_families.py gpurun_out/profile_dump.csv > gpurun_out/profile_families.txt 2>&1
head -30 gpurun_out/profile_families.txt
