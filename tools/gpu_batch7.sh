#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_conv_halo_gpu.py -m gpu -q -p no:cacheprovider -x 2>&1 | tail -8 > gpurun_out/r2b7_halo_bo1.log
BB200_HALO_BO=0 timeout 300 python -m pytest tests/test_conv_halo_gpu.py -m gpu -q -p no:cacheprovider -x 2>&1 | tail -8 > gpurun_out/r2b7_halo_bo0.log
BB200_NO_HALO=1 timeout 600 python -m pytest tests/test_plan_gpu.py tests/test_plan_cache_gpu.py tests/test_parity_gpu.py -m gpu -q -p no:cacheprovider -k "fourconv or cache" 2>&1 | tail -30 > gpurun_out/r2b7_tests_nohalo.log
BB200_NO_HALO=1 timeout 600 python bench.py --workload implicit_maml --steps 5 --no-cpu-baseline > gpurun_out/r2b7_bench_nohalo.json 2> gpurun_out/r2b7_bench_nohalo.err
echo "== halo bo1"; cat gpurun_out/r2b7_halo_bo1.log | cut -c1-200; echo "== halo bo0"; cat gpurun_out/r2b7_halo_bo0.log | cut -c1-200
echo "== tests nohalo"; tail -12 gpurun_out/r2b7_tests_nohalo.log | cut -c1-200; cut -c1-200 gpurun_out/r2b7_bench_nohalo.json; tail -2 gpurun_out/r2b7_bench_nohalo.err
