#!/bin/bash
# one-shot check of the experimental multicast-pair attention variant (HYVAE_ATTN_MULTICAST=1)
mkdir -p gpurun_out
HYVAE_ATTN_MULTICAST=1 timeout 100 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "attn" 2>&1 | tail -6 | tee gpurun_out/attn_mc_tests.log
HYVAE_ATTN_MULTICAST=1 timeout 60 python tools/bench_attn.py 17 1024 512 2>&1 | tail -5 | tee gpurun_out/attn_mc_bench.log
timeout 60 python tools/bench_attn.py 17 1024 512 fused 2>&1 | tail -3 | tee gpurun_out/attn_plain_bench.log
