#!/usr/bin/env bash
# round-2 re-entry checkpoint: full GPU suite, default bench (no CPU legs), launch list, ncu --set full of the K-loop kernels
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
T0=$(date +%s)
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -40 > gpurun_out/r2b25_tests.log
echo "tests done $(( $(date +%s) - T0 )) s" >> gpurun_out/r2b25_tests.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2b25_bench.json 2> gpurun_out/r2b25_bench.err
echo "bench done $(( $(date +%s) - T0 )) s" >> gpurun_out/r2b25_tests.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2b25_launches_maml.csv python bench.py --workload implicit_maml --steps 1 --warmup 3 --no-cpu-baseline --no-graph --e2e-steps 1 --no-extra > gpurun_out/r2b25_ncu_launches.log 2>&1
echo "launch list done $(( $(date +%s) - T0 )) s" >> gpurun_out/r2b25_tests.log
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:"cb_tf_mma|cb_reduce_mma|cb2_dense|cb2_final|cb2_stats|cb2_reduce_kernel|conv_halo|wgrad_halo" --launch-skip 40 -c 22 -o gpurun_out/r2b25_kloop python bench.py --workload implicit_maml --steps 1 --warmup 3 --no-cpu-baseline --no-graph --e2e-steps 1 --no-extra > gpurun_out/r2b25_ncu_full.log 2>&1
echo "ncu full done $(( $(date +%s) - T0 )) s" >> gpurun_out/r2b25_tests.log
ls -la gpurun_out/ | cut -c1-150
cat gpurun_out/r2b25_tests.log | cut -c1-220
cut -c1-400 gpurun_out/r2b25_bench.json; tail -3 gpurun_out/r2b25_bench.err
