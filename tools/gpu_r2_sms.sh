#!/bin/bash
# A/B of HYVAE_CONV_SMS (SMs the persistent conv kernels occupy; the rest serve the other tile stream's HBM-bound passes)
mkdir -p gpurun_out
for s in ${SMS:-148 132 124}; do
  for ts in ${TS:-2}; do
    HYVAE_CONV_SMS=$s timeout 600 python bench.py --steps 2 --warmup 2 --tile-streams $ts --no-cpu-baseline --no-torch-baseline --no-e2e --no-profile > gpurun_out/bench_sms${s}_ts${ts}.json 2> gpurun_out/bench_sms${s}_ts${ts}.err
    python - <<PY
import json
d = json.loads(open("gpurun_out/bench_sms${s}_ts${ts}.json").read().strip().splitlines()[-1])
print("conv_sms $s streams $ts: value", round(d["value"], 2), "ms", round(d["ms_per_step"], 1), d["clocks"])
PY
  done
done
