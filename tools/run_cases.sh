#!/bin/bash
# Run every native-plan GPU case in its own process (one sticky CUDA error must not poison the rest);
# on failure re-run under compute-sanitizer to name the kernel.  Output: gpurun_out/cases.log
mkdir -p gpurun_out
out=gpurun_out/cases.log
: > $out
for c in logistic mlp lenet lenet_b300 fourconv fourconv_mini roberta fourconv_bf16 roberta_bf16; do
  echo "=== $c" >> $out
  CUDA_LAUNCH_BLOCKING=1 timeout 300 python -m pytest tests/test_plan_gpu.py -q --tb=short -p no:cacheprovider -k "test_plan_matches and [$c]" 2>&1 | grep -v "^$" | tail -25 >> $out
  if grep -q "illegal memory\|AcceleratorError\|status 7" <(tail -30 $out); then
    echo "--- sanitizer $c" >> $out
    timeout 600 compute-sanitizer --tool memcheck --print-limit 3 python -m pytest tests/test_plan_gpu.py -q --tb=no -p no:cacheprovider -k "test_plan_matches and [$c]" 2>&1 | grep -A22 "Invalid\|Error:" | head -60 >> $out
  fi
done
echo "=== graph" >> $out
CUDA_LAUNCH_BLOCKING=0 timeout 300 python -m pytest tests/test_plan_gpu.py -q --tb=short -p no:cacheprovider -k "graph_replay" 2>&1 | tail -15 >> $out
echo "=== parity(native)" >> $out
timeout 600 python -m pytest tests/test_parity_gpu.py -q --tb=line -p no:cacheprovider -k "native" 2>&1 | tail -25 >> $out
echo "=== bench lenet native" >> $out
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_native.json 2> gpurun_out/bench_native.err
tail -c 3000 gpurun_out/bench_native.json >> $out; tail -5 gpurun_out/bench_native.err >> $out
