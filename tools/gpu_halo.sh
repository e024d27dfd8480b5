#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -4 | tee gpurun_out/tests_gpu.log
HYVAE_PROFILE_DUMP=gpurun_out/profile_dump.csv timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_halo.json
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_halo.json"))
print("value", round(d["value"], 3), "e2e", round(d["e2e"]["value"], 3), "ms", round(d["ms_per_step"], 1), "clk", d["clocks"]["sm_mhz"], d["kernel_ms_per_step_rank0"], "launches", d["gpu_launches"])
PY
python tools/profile_families.py gpurun_out/profile_dump.csv > gpurun_out/profile_families.txt; grep -E "pad_upsample|total" gpurun_out/profile_families.txt
