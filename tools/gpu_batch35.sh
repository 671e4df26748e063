#!/usr/bin/env bash
# round-2 final: full GPU suite + default bench line (CPU legs, all sub-records)
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
T0=$(date +%s)
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -12 > gpurun_out/r2b35_tests.log
echo "tests done $(( $(date +%s) - T0 )) s" >> gpurun_out/r2b35_tests.log
timeout 900 python bench.py > gpurun_out/r2b35_bench.json 2> gpurun_out/r2b35_bench.err
echo "bench done $(( $(date +%s) - T0 )) s" >> gpurun_out/r2b35_tests.log
cat gpurun_out/r2b35_tests.log | cut -c1-220
cut -c1-300 gpurun_out/r2b35_bench.json; tail -3 gpurun_out/r2b35_bench.err
