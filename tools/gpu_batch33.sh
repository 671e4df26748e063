#!/usr/bin/env bash
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
for i in 1 2 3 4; do
  timeout 200 python -m pytest tests/test_parity_gpu.py -m gpu -q -s -p no:cacheprovider -k "lenet_cg and native" 2>&1 | grep -E "parity\]|rel|passed|failed|Error|assert" | cut -c1-260
done > gpurun_out/r2b33_new.log 2>&1
for i in 1 2 3 4; do
  BB200_CHANSUM_V1=1 timeout 200 python -m pytest tests/test_parity_gpu.py -m gpu -q -s -p no:cacheprovider -k "lenet_cg and native" 2>&1 | grep -E "parity\]|rel|passed|failed|Error|assert" | cut -c1-260
done > gpurun_out/r2b33_v1.log 2>&1
for i in 1 2; do
  BB200_CHANSUM_V1=1 BB200_CONV_SMALL_V1=1 timeout 200 python -m pytest tests/test_parity_gpu.py -m gpu -q -s -p no:cacheprovider -k "lenet_cg and native" 2>&1 | grep -E "parity\]|rel|passed|failed|Error|assert" | cut -c1-260
done > gpurun_out/r2b33_allv1.log 2>&1
echo NEW; cat gpurun_out/r2b33_new.log; echo CHANSUM_V1; cat gpurun_out/r2b33_v1.log; echo ALLV1; cat gpurun_out/r2b33_allv1.log
