#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "quantisation or psnr_ssim or sweep_config or tile_streams" 2>&1 | tail -15 | tee gpurun_out/metrics_tests.log
timeout 600 python bench.py --steps 2 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_streams2.json | cut -c1-600
