#!/usr/bin/env bash
mkdir -p gpurun_out
python -m pytest tests/test_bf16_parity_gpu.py -m gpu -q -s -p no:cacheprovider 2>&1 | grep -E "\[bf16 parity\]|passed|failed|Error" | sed 's/^[.F]*//' > gpurun_out/r2b4_bf16.log
python -m pytest tests -m gpu -q -p no:cacheprovider --deselect tests/test_bf16_parity_gpu.py 2>&1 | tail -70 > gpurun_out/r2b4_tests.log
ncu --set full --import-source on --clock-control none -k regex:"cb_(tf|reduce)" -c 4 -f -o gpurun_out/r2b4_cb python bench.py --workload implicit_maml --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/r2b4_ncu.log 2>&1
python bench.py --workload neural_architecture_search --steps 5 > gpurun_out/r2b4_bench_nas.json 2> gpurun_out/r2b4_bench_nas.err
python bench.py --steps 5 --no-cpu-baseline > gpurun_out/r2b4_bench.json 2> gpurun_out/r2b4_bench.err
cat gpurun_out/r2b4_bf16.log | cut -c1-250; tail -25 gpurun_out/r2b4_tests.log; tail -3 gpurun_out/r2b4_ncu.log; cut -c1-300 gpurun_out/r2b4_bench_nas.json; tail -2 gpurun_out/r2b4_bench_nas.err; cut -c1-200 gpurun_out/r2b4_bench.json
