#!/bin/bash
# Round-2 quick session: kernel/model suite (without the 3-minute production-shape suite unless PROD=1), smoke, bench A/B.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x --ignore=tests/test_gpu_production_shapes.py 2>&1 | tail -12 | tee gpurun_out/tests_gpu.log
if [ "${PROD:-0}" = "1" ]; then
  timeout 1500 python -m pytest tests/test_gpu_production_shapes.py -q -m gpu -s 2>&1 | tail -30 | tee gpurun_out/tests_production.log
fi
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke.log
HYVAE_PROFILE_DUMP=gpurun_out/profile_dump.csv timeout 900 python bench.py --steps ${STEPS:-2} --warmup ${WARMUP:-3} --no-cpu-baseline --no-torch-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -3 gpurun_out/bench.err; python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench.json").read().strip().splitlines()[-1])
print("value", d["value"], "e2e", d["e2e"]["value"], "ms", d["ms_per_step"], "clocks", d["clocks"])
print("roofline achieved", d["roofline"]["achieved"], "executed", d["roofline"]["executed_tflops"], d["kernel_ms_per_step_rank0"])
PY
python tools/profile_families.py gpurun_out/profile_dump.csv > gpurun_out/profile_families.txt 2>&1
head -32 gpurun_out/profile_families.txt
if [ -n "${AB:-}" ]; then
  env ${AB} timeout 900 python bench.py --steps ${STEPS:-2} --warmup ${WARMUP:-3} --no-cpu-baseline --no-torch-baseline --no-e2e > gpurun_out/bench_ab.json 2> gpurun_out/bench_ab.err
  python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_ab.json").read().strip().splitlines()[-1])
print("A/B value", d["value"], "ms", d["ms_per_step"], "clocks", d["clocks"], d["kernel_ms_per_step_rank0"])
PY
fi
