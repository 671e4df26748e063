#!/usr/bin/env bash
# round-2 final checkpoint: full GPU suite, default bench with CPU legs and all sub-records, launch lists of configs 2 and 5
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
T0=$(date +%s)
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -25 > gpurun_out/r2b30_tests.log
echo "tests done $(( $(date +%s) - T0 )) s" >> gpurun_out/r2b30_tests.log
timeout 900 python bench.py > gpurun_out/r2b30_bench.json 2> gpurun_out/r2b30_bench.err
echo "bench done $(( $(date +%s) - T0 )) s" >> gpurun_out/r2b30_tests.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2b30_launches_lenet.csv python bench.py --workload learning_to_reweight --steps 1 --warmup 3 --no-cpu-baseline --no-graph --e2e-steps 1 > gpurun_out/r2b30_ncu_lenet.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 6000 -c 3000 --csv --log-file gpurun_out/r2b30_launches_bert.csv python bench.py --workload bert_data_reweighting --steps 1 --warmup 3 --no-cpu-baseline --no-graph --e2e-steps 1 > gpurun_out/r2b30_ncu_bert.log 2>&1
echo "launch lists done $(( $(date +%s) - T0 )) s" >> gpurun_out/r2b30_tests.log
cat gpurun_out/r2b30_tests.log | cut -c1-220
cut -c1-300 gpurun_out/r2b30_bench.json; tail -3 gpurun_out/r2b30_bench.err
