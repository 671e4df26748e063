#!/usr/bin/env bash
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_plan_gpu.py tests/test_plan_cache_gpu.py tests/test_parity_gpu.py tests/test_bf16_parity_gpu.py tests/test_reference_gpu.py -m gpu -q -p no:cacheprovider -x 2>&1 | tail -6 > gpurun_out/r2b21_tests.log
timeout -s KILL 200 python tools/time_call.py implicit_maml > gpurun_out/r2b21_time_call.log 2>&1
timeout -s KILL 600 python bench.py --workload implicit_maml --steps 5 --no-cpu-baseline --no-extra > gpurun_out/r2b21_bench.json 2> gpurun_out/r2b21_bench.err
tail -3 gpurun_out/r2b21_tests.log | cut -c1-200; tail -3 gpurun_out/r2b21_time_call.log; python -c "
import json
d=json.loads(open('gpurun_out/r2b21_bench.json').read().strip().splitlines()[-1])
print('value',d['value'],'e2e',d['e2e']['value'])"; tail -2 gpurun_out/r2b21_bench.err
