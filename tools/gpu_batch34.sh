#!/usr/bin/env bash
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
echo "== default" > gpurun_out/r2b34.log
timeout 300 python tools/parity_vs_fp64.py "" 1 >> gpurun_out/r2b34.log 2>&1
timeout 200 python tools/parity_vs_fp64.py lenet 4 >> gpurun_out/r2b34.log 2>&1
echo "== BB200_CHANSUM_V1" >> gpurun_out/r2b34.log
BB200_CHANSUM_V1=1 timeout 200 python tools/parity_vs_fp64.py lenet 4 >> gpurun_out/r2b34.log 2>&1
echo "== all first-generation" >> gpurun_out/r2b34.log
BB200_CHANSUM_V1=1 BB200_CONV_SMALL_V1=1 timeout 200 python tools/parity_vs_fp64.py lenet 4 >> gpurun_out/r2b34.log 2>&1
grep -v Warning gpurun_out/r2b34.log | cut -c1-200
