#!/bin/bash
# End-of-session check on one GPU: full gpu test suite, smoke(), the default bench line, and BASELINE configs 2 and 3.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -x 2>&1 | tail -5 | tee gpurun_out/tests_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke.log
HYVAE_PROFILE_DUMP=gpurun_out/profile_dump.csv timeout 900 python bench.py 2>&1 | tail -1 > gpurun_out/bench_default.json
python tools/profile_families.py gpurun_out/profile_dump.csv > gpurun_out/profile_families.txt 2>&1
timeout 600 python bench.py --workload config2 --steps 2 --warmup 2 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_config2.json
timeout 600 python bench.py --workload config3 --steps 3 --warmup 2 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_config3.json
python - <<'PY'
import json
for n in ("default", "config2", "config3"):
    try:
        d = json.load(open(f"gpurun_out/bench_{n}.json"))
        print(n, d["metric"], "value", round(d["value"], 2), "e2e", d["e2e"] and round(d["e2e"]["value"], 2), "ms", round(d["ms_per_step"], 1), "clk", d["clocks"]["sm_mhz"], d["clocks"]["reasons"], "cpu", d.get("cpu_baseline", {}).get("value"))
    except Exception as e:
        print(n, "FAILED", e)
PY
