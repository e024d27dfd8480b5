#!/bin/bash
# production-shape parity with the Winograd path on, then ncu --set full of the Winograd conv and its GroupNorm producer
mkdir -p gpurun_out
if [ "${PROD:-1}" = "1" ]; then
  timeout 1500 python -m pytest tests/test_gpu_production_shapes.py -q -m gpu -s 2>&1 | grep -v "^$" | tail -20 | tee gpurun_out/tests_production.log
fi
bash tools/gpu_ncu.sh wino128:conv_wino wino256:conv_wino winogn128:gn_apply_wino
for c in wino128 wino256 winogn128; do
  ncu -i gpurun_out/prof_$c.ncu-rep --page details > gpurun_out/ncu_${c}_details.txt 2>&1
  ncu -i gpurun_out/prof_$c.ncu-rep --page raw --csv > gpurun_out/ncu_${c}_raw.csv 2>&1
done
ls -la gpurun_out/
