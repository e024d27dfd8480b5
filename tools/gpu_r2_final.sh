#!/bin/bash
# End-of-round check on one GPU: the FULL gpu suite (incl. the production-shape parity tests), smoke(), the default bench line
# (with the CPU and PyTorch/cuDNN legs), BASELINE configs 2 and 3, and the reference arm.
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu -x 2>&1 | tail -6 | tee gpurun_out/tests_gpu_full.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke.log
HYVAE_PROFILE_DUMP=gpurun_out/profile_dump.csv timeout 1200 python bench.py 2> gpurun_out/bench_default.err | tail -1 > gpurun_out/bench_default.json
python tools/profile_families.py gpurun_out/profile_dump.csv > gpurun_out/profile_families.txt 2>&1
timeout 600 python bench.py --workload config2 --steps 2 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/bench_config2.json
timeout 600 python bench.py --workload config3 --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/bench_config3.json
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 2>/dev/null | tail -1 > gpurun_out/bench_reference.json
python - <<'PY'
import json
for n in ("default", "config2", "config3", "reference"):
    try:
        d = json.load(open(f"gpurun_out/bench_{n}.json"))
        print(n, d["metric"], "value", round(d["value"], 3), "e2e", d["e2e"] and round(d["e2e"]["value"], 3), "ms", round(d["ms_per_step"], 1),
              "clk", d.get("clocks", {}).get("sm_mhz"), "cpu", d.get("cpu_baseline", {}).get("value"), "torch", d.get("torch_gpu_baseline", {}).get("value"))
    except Exception as e:
        print(n, "FAILED", e)
PY
