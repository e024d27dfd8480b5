#!/bin/bash
# Scaling session (gpurun --gpus N): headline bench (config 4, tiles over ranks; NCCL gather) and, if PUSH=1, the same with the
# opt-in peer-memory tile push; then BASELINE config 3 (clip per rank).  usage: N=8 bash tools/gpu_r2_scale.sh
N=${N:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
if [ "${CFG4:-1}" = "1" ]; then
  timeout 600 $TR --master-port 29512 bench.py --gpus $N --steps ${STEPS:-4} --warmup ${WARMUP:-3} > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err
  tail -2 gpurun_out/bench_${N}gpu.err
  if [ "${PUSH:-0}" = "1" ]; then   # the opt-in peer-memory tile push (default exchange: NCCL gather)
    HYVAE_TILE_PUSH=1 timeout 600 $TR --master-port 29513 bench.py --gpus $N --steps ${STEPS:-4} --warmup ${WARMUP:-3} --no-profile > gpurun_out/bench_${N}gpu_push.json 2> gpurun_out/bench_${N}gpu_push.err
  fi
fi
timeout 600 $TR --master-port 29514 bench.py --gpus $N --workload config3 --steps ${STEPS:-4} --warmup ${WARMUP:-3} > gpurun_out/bench_config3_${N}gpu.json 2> gpurun_out/bench_config3_${N}gpu.err
tail -2 gpurun_out/bench_config3_${N}gpu.err
python - <<PY
import json
for f in ["gpurun_out/bench_${N}gpu.json", "gpurun_out/bench_${N}gpu_push.json", "gpurun_out/bench_config3_${N}gpu.json"]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["metric"], "value", round(d["value"], 2), "ms", round(d["ms_per_step"], 1), "e2e", d["e2e"] and round(d["e2e"]["value"], 2), d["config"].get("tile_exchange"), d["clocks"])
    except Exception as e:
        print(f, "unreadable:", e)
PY
