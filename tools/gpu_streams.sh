#!/bin/bash
# A/B of the tile-stream count on the headline workload (device-timed value only; per-kernel events overlap).
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "tile_streams or reproducible" 2>&1 | tail -5 | tee gpurun_out/streams_tests.log
for n in ${STREAMS:-1 2 3}; do
  HYVAE_TILE_STREAMS=$n timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline ${EXTRA} 2>&1 | tail -1 > gpurun_out/bench_streams_$n.json
  python - <<PY
import json
d = json.load(open("gpurun_out/bench_streams_$n.json"))
print("streams=$n value", round(d["value"], 3), "e2e", d["e2e"] and round(d["e2e"]["value"], 3), "ms", round(d["ms_per_step"], 1), "clk", d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
PY
done
