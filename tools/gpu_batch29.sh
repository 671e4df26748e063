#!/usr/bin/env bash
# ncu --set full of the LeNet small-channel kernels (second generation), one K-loop iteration
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:"conv_small" --launch-skip 30 -c 6 -o gpurun_out/r2b29_lenet python bench.py --workload learning_to_reweight --steps 1 --warmup 1 --no-cpu-baseline --no-graph --e2e-steps 1 > gpurun_out/r2b29_ncu.log 2>&1
tail -3 gpurun_out/r2b29_ncu.log | cut -c1-200
ncu -i gpurun_out/r2b29_lenet.ncu-rep --page raw --csv > gpurun_out/r2b29_lenet_raw.csv 2>/dev/null
ls -la gpurun_out/r2b29*
