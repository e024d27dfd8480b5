#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "kw_packed or thin or model_16bit or golden or smoke" 2>&1 | tail -5 | tee gpurun_out/kwpack_tests.log
timeout 300 python tools/bench_tfold.py 2>&1 | tail -8 | tee gpurun_out/bench_tfold_micro.log
for f in 1 0; do
  HYVAE_KWPACK=$f HYVAE_PROFILE_DUMP=gpurun_out/profile_dump_kw_$f.csv timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1 > gpurun_out/bench_kw_$f.json
  python - <<PY
import json
d = json.load(open("gpurun_out/bench_kw_$f.json"))
print("kwpack=$f value", round(d["value"], 3), "ms", round(d["ms_per_step"], 1), "clk", d["clocks"]["sm_mhz"], "conv ms", round(d["roofline"]["ms_per_step"],1))
PY
  grep -h "thin" gpurun_out/profile_dump_kw_$f.csv | awk -F, '{s+=$4; n+=1} END {print "  thin conv_in launches", n, "total ms", s, "avg us", 1000*s/n}'
done
