#!/bin/bash
# Multi-GPU session (gpurun --gpus N): bit-identity of the tile-parallel round trip (peer-memory push and NCCL gather),
# then the headline bench with both exchanges.  usage: N=2 bash tools/gpu_r2_multi.sh
N=${N:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "blend" 2>&1 | tail -3 | tee gpurun_out/tests_blend.log
HYVAE_TILE_PUSH_DEBUG=1 timeout 300 $TR --master-port 29511 tools/check_tile_parallel.py 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM" | tail -12 | tee gpurun_out/check_tile_parallel_${N}gpu.log
timeout 600 $TR --master-port 29512 bench.py --gpus $N --steps ${STEPS:-3} --warmup ${WARMUP:-2} > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err
tail -2 gpurun_out/bench_${N}gpu.err
if [ "${GATHER:-1}" = "1" ]; then
  HYVAE_TILE_PUSH=0 timeout 600 $TR --master-port 29513 bench.py --gpus $N --steps ${STEPS:-3} --warmup ${WARMUP:-2} --no-e2e > gpurun_out/bench_${N}gpu_gather.json 2> gpurun_out/bench_${N}gpu_gather.err
fi
python - <<PY
import json
for f in ["gpurun_out/bench_${N}gpu.json", "gpurun_out/bench_${N}gpu_gather.json"]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value", round(d["value"], 2), "ms", round(d["ms_per_step"], 1), "e2e", d["e2e"] and round(d["e2e"]["value"], 2), d["config"].get("tile_exchange"), d["clocks"])
    except Exception as e:
        print(f, "unreadable:", e)
PY
