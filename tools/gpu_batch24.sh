#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_plan_gpu.py tests/test_plan_cache_gpu.py tests/test_parity_gpu.py tests/test_bf16_parity_gpu.py tests/test_reference_gpu.py -m gpu -q -p no:cacheprovider -k "fourconv or cache" 2>&1 | tail -30 > gpurun_out/r2b24_tests.log
timeout 600 python bench.py --workload implicit_maml --steps 5 --no-cpu-baseline > gpurun_out/r2b24_bench.json 2> gpurun_out/r2b24_bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2b24_launches_maml.csv python bench.py --workload implicit_maml --steps 1 --warmup 3 --no-cpu-baseline --no-graph --e2e-steps 1 > gpurun_out/r2b24_ncu.log 2>&1
cat gpurun_out/r2b24_halo.log | cut -c1-200; tail -8 gpurun_out/r2b24_tests.log | cut -c1-200; cut -c1-200 gpurun_out/r2b24_bench.json; tail -2 gpurun_out/r2b24_bench.err
