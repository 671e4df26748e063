#!/usr/bin/env bash
mkdir -p gpurun_out
python -m pytest tests/test_bf16_parity_gpu.py -m gpu -q -s -p no:cacheprovider 2>&1 | grep -E "\[bf16 parity\]|passed|failed" > gpurun_out/r2b2_bf16.log
python -m pytest tests/test_plan_gpu.py tests/test_parity_gpu.py tests/test_reference_gpu.py -m gpu -q -p no:cacheprovider -k "fourconv or graph_replay" 2>&1 | tail -60 > gpurun_out/r2b2_convblock_tests.log
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_plan_gpu.py -m gpu -q -p no:cacheprovider -k "fourconv_mini and not bf16" 2>&1 | tail -25 > gpurun_out/r2b2_sanitizer.log
python bench.py --workload implicit_maml --steps 5 --no-cpu-baseline > gpurun_out/r2b2_bench_maml.json 2> gpurun_out/r2b2_bench_maml.err
cat gpurun_out/r2b2_bf16.log; tail -15 gpurun_out/r2b2_convblock_tests.log; tail -8 gpurun_out/r2b2_sanitizer.log; cut -c1-400 gpurun_out/r2b2_bench_maml.json; tail -3 gpurun_out/r2b2_bench_maml.err
