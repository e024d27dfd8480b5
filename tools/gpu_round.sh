#!/bin/bash
# One GPU session: staged so that a fault in the tcgen05 kernel cannot hide the results of the rest.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
echo "== stage 1: kernels + blocks (no tcgen05)"; 
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "not tc_ and not model and not full_size" 2>&1 | tail -15 | tee gpurun_out/stage1.log
echo "== stage 2: whole model on the CUDA-core path"
HYVAE_FORCE_DIRECT=1 timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "model_fp32 or model_16bit or pure_bf16" 2>&1 | tail -15 | tee gpurun_out/stage2.log
echo "== stage 3: tcgen05 conv"
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "tc_conv or tc_gemm or tc_big or tc_thin or tc_cta" 2>&1 | tail -30 | tee gpurun_out/stage3.log
echo "== stage 4: model with tcgen05"
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "model_16bit or tc_model or full_size or reproducible or pure_bf16" 2>&1 | tail -30 | tee gpurun_out/stage4.log
echo "== stage 5: microbench"
timeout 600 python tools/bench_conv.py 2>&1 | tail -20 | tee gpurun_out/bench_conv.log
echo "== stage 6: smoke + bench"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/smoke.log
HYVAE_PROFILE_DUMP=gpurun_out/profile_dump.csv timeout 900 python bench.py --steps 1 --warmup 1 2>&1 | tail -5 | tee gpurun_out/bench.log
python tools/summarize_profile.py gpurun_out/profile_dump.csv 45 | tee gpurun_out/profile_summary.txt
