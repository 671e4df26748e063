#!/usr/bin/env bash
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_reference_gpu.py tests/test_parity_gpu.py -m gpu -q -s -p no:cacheprovider 2>&1 | grep -E "parity\]|passed|failed|Error|error" | cut -c1-200 > gpurun_out/r2b36_tests.log
cat gpurun_out/r2b36_tests.log | tail -60
