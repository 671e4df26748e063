#!/usr/bin/env bash
# round-2 batch 1: full GPU test suite (new bf16 / reference / conv-TMA tests) + the default bench line
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -s -p no:cacheprovider 2>&1 | tail -150 > gpurun_out/r2b1_tests.log
python bench.py --steps 5 --warmup 3 > gpurun_out/r2b1_bench.json 2> gpurun_out/r2b1_bench.err
tail -c 3000 gpurun_out/r2b1_tests.log
