#!/usr/bin/env bash
# dev call: native BatchNorm forward in the prologue (tests, A/B of the whole call), caller snapshot on CUDA
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
T0=$(date +%s)
timeout 600 python -m pytest tests/test_bn_fwd_gpu.py tests/test_callers_cpu.py -m gpu -q -p no:cacheprovider -x 2>&1 | tail -15 > gpurun_out/r2b26_tests.log
echo "tests done $(( $(date +%s) - T0 )) s" >> gpurun_out/r2b26_tests.log
timeout 600 python -m pytest tests/test_bf16_parity_gpu.py tests/test_reference_gpu.py tests/test_plan_cache_gpu.py -m gpu -q -p no:cacheprovider -k "fourconv or cache or maml" 2>&1 | tail -8 >> gpurun_out/r2b26_tests.log
echo "parity done $(( $(date +%s) - T0 )) s" >> gpurun_out/r2b26_tests.log
timeout 300 python tools/time_call.py implicit_maml > gpurun_out/r2b26_time_call_bn.log 2>&1
BB200_PROLOGUE_BN_MIN=0 timeout 300 python tools/time_call.py implicit_maml > gpurun_out/r2b26_time_call_nobn.log 2>&1
timeout 300 python tools/time_call.py bert_data_reweighting > gpurun_out/r2b26_time_call_bert.log 2>&1
timeout 600 python bench.py --workload implicit_maml --no-cpu-baseline --no-extra > gpurun_out/r2b26_bench.json 2> gpurun_out/r2b26_bench.err
echo "bench done $(( $(date +%s) - T0 )) s" >> gpurun_out/r2b26_tests.log
cat gpurun_out/r2b26_tests.log | cut -c1-220
tail -3 gpurun_out/r2b26_time_call_bn.log; tail -3 gpurun_out/r2b26_time_call_nobn.log; tail -3 gpurun_out/r2b26_time_call_bert.log
cut -c1-300 gpurun_out/r2b26_bench.json; tail -3 gpurun_out/r2b26_bench.err
