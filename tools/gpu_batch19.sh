#!/usr/bin/env bash
# checkpoint run: full GPU suite, smoke, default bench (+extras, cpu baseline), reference arm, call breakdown
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
timeout -s KILL 1200 python -m pytest tests -q -m gpu -p no:cacheprovider -x 2>&1 | tail -6 > gpurun_out/r2b19_tests.log
timeout -s KILL 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2b19_smoke.log 2>&1
timeout -s KILL 600 python bench.py > gpurun_out/r2b19_default.json 2> gpurun_out/r2b19_default.err
timeout -s KILL 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2b19_default_ref.json 2> gpurun_out/r2b19_default_ref.err
timeout -s KILL 200 python tools/time_call.py implicit_maml > gpurun_out/r2b19_time_call.log 2>&1
tail -3 gpurun_out/r2b19_tests.log | cut -c1-200; tail -1 gpurun_out/r2b19_smoke.log; cut -c1-1500 gpurun_out/r2b19_default.json; tail -2 gpurun_out/r2b19_default.err; cut -c1-600 gpurun_out/r2b19_default_ref.json; cat gpurun_out/r2b19_time_call.log | tail -3
