#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "groupnorm or model or resnet or reproducible" 2>&1 | tail -3 | tee gpurun_out/gn_tests.log
timeout 200 python tools/bench_gn.py 2>&1 | tail -5 | tee gpurun_out/bench_gn.log
HYVAE_PROFILE_DUMP=gpurun_out/profile_dump.csv timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1 > gpurun_out/bench_gn.json
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_gn.json"))
print("value", round(d["value"], 3), "ms", round(d["ms_per_step"], 1), "clk", d["clocks"]["sm_mhz"], d["kernel_ms_per_step_rank0"])
PY
python tools/profile_families.py gpurun_out/profile_dump.csv | grep -E "gn_apply|total"
