#!/bin/bash
# GPU session for the fused attention kernel: small parity tests first (bounded), then the canonical shape.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "attn or midblock" 2>&1 | tail -15 | tee gpurun_out/attn_tests.log
timeout 120 python tools/bench_attn.py 9 288 512 2>&1 | tail -6 | tee gpurun_out/attn_bench_small.log
timeout 180 python tools/bench_attn.py 2>&1 | tail -6 | tee gpurun_out/attn_bench.log
