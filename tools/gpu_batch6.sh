#!/usr/bin/env bash
mkdir -p gpurun_out
python -m pytest tests/test_plan_gpu.py tests/test_plan_cache_gpu.py tests/test_parity_gpu.py -m gpu -q -p no:cacheprovider -k "fourconv or cache" 2>&1 | tail -60 > gpurun_out/r2b6_tests.log
python -m pytest tests/test_bf16_parity_gpu.py -m gpu -q -s -p no:cacheprovider 2>&1 | grep -E "\[bf16 parity\]|passed|failed|Error" | sed 's/^[.F]*//' > gpurun_out/r2b6_bf16.log
python bench.py --workload implicit_maml --steps 5 --no-cpu-baseline > gpurun_out/r2b6_bench_maml.json 2> gpurun_out/r2b6_bench_maml.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2b6_launches_maml.csv python bench.py --workload implicit_maml --steps 1 --warmup 3 --no-cpu-baseline --no-graph --e2e-steps 1 > gpurun_out/r2b6_ncu.log 2>&1
tail -30 gpurun_out/r2b6_tests.log; cut -c1-200 gpurun_out/r2b6_bf16.log | tail -8; cut -c1-300 gpurun_out/r2b6_bench_maml.json; tail -3 gpurun_out/r2b6_bench_maml.err
