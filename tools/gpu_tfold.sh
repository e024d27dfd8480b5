#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "temporal_fold or model or reproducible or tile_streams" 2>&1 | tail -5 | tee gpurun_out/tfold_tests.log
for f in ${ORDER:-1 0}; do
  HYVAE_TFOLD=$f HYVAE_PROFILE_DUMP=gpurun_out/profile_dump_tfold_$f.csv timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1 > gpurun_out/bench_tfold_$f.json
  python - <<PY
import json
d = json.load(open("gpurun_out/bench_tfold_$f.json"))
print("tfold=$f value", round(d["value"], 3), "ms", round(d["ms_per_step"], 1), "clk", d["clocks"]["sm_mhz"], "conv ms", round(d["roofline"]["ms_per_step"],1), "exec", round(d["roofline"]["executed_tflops"],1), "alg", round(d["roofline"]["achieved"],1))
PY
  python tools/profile_families.py gpurun_out/profile_dump_tfold_$f.csv > gpurun_out/families_tfold_$f.txt
done
