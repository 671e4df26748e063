#!/usr/bin/env bash
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:"cb_tf_mma|cb_reduce_mma|cb2_dense|cb2_final|cb2_stats|cb2_reduce_kernel|conv_halo|wgrad_halo" --launch-skip 30 -c 16 -o gpurun_out/r2b23_kloop python bench.py --workload implicit_maml --steps 1 --warmup 3 --no-cpu-baseline --no-graph --e2e-steps 1 --no-extra > gpurun_out/r2b23_ncu.log 2>&1
tail -3 gpurun_out/r2b23_ncu.log | cut -c1-200
ls -la gpurun_out/r2b23_kloop.ncu-rep
