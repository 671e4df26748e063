#!/usr/bin/env bash
mkdir -p gpurun_out
python -m pytest tests/test_bf16_parity_gpu.py -m gpu -q -s -p no:cacheprovider 2>&1 | grep -E "^\[bf16 parity\]|^F?\[bf16|passed|failed|Error" > gpurun_out/r2b3_bf16.log
python -m pytest tests/test_plan_cache_gpu.py tests/test_plan_gpu.py tests/test_parity_gpu.py tests/test_reference_gpu.py -m gpu -q -p no:cacheprovider 2>&1 | tail -80 > gpurun_out/r2b3_tests.log
python bench.py --workload implicit_maml --steps 5 --no-cpu-baseline > gpurun_out/r2b3_bench_maml.json 2> gpurun_out/r2b3_bench_maml.err
cat gpurun_out/r2b3_bf16.log | cut -c1-250; tail -30 gpurun_out/r2b3_tests.log; cut -c1-300 gpurun_out/r2b3_bench_maml.json; tail -3 gpurun_out/r2b3_bench_maml.err
