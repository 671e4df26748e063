#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"cb_tf_kernel|cb_reduce_kernel|cb2_dense_kernel" -s 6 -c 4 -f -o gpurun_out/r2b12_cb python bench.py --workload implicit_maml --steps 1 --warmup 3 --no-cpu-baseline --no-graph --e2e-steps 1 > gpurun_out/r2b12_ncu.log 2>&1
timeout 600 python tools/time_call.py implicit_maml > gpurun_out/r2b12_time_call.log 2>&1
tail -3 gpurun_out/r2b12_ncu.log; cat gpurun_out/r2b12_time_call.log | tail -5
