#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/test_wino.py ${1:-} 2>&1 | tail -40 | tee gpurun_out/wino.log
