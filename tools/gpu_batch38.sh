#!/usr/bin/env bash
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
timeout 150 python -m pytest tests/test_cg_global_gpu.py -m gpu -q -s -x -p no:cacheprovider -k two_ranks 2>&1 | grep -vE "Warning|warn" | tail -25 | cut -c1-220 > gpurun_out/r2b38.log
cat gpurun_out/r2b38.log
