#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_conv_halo_gpu.py -m gpu -q -p no:cacheprovider -x -k "bf16_padded_output and 5x7" 2>&1 | grep -v "^$" | head -60 > gpurun_out/r2b11_sanitizer.log
cat gpurun_out/r2b11_sanitizer.log | cut -c1-220
