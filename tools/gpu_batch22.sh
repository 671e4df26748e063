#!/usr/bin/env bash
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_conv_halo_gpu.py tests/test_conv_tma_gpu.py tests/test_gemm_tma_gpu.py -m gpu -q -p no:cacheprovider -x 2>&1 | tail -4 > gpurun_out/r2b22_unit.log
timeout -s KILL 200 python tools/halo_bench.py > gpurun_out/r2b22_halobench.log 2>&1
timeout -s KILL 900 python -m pytest tests/test_plan_gpu.py tests/test_bf16_parity_gpu.py -m gpu -q -p no:cacheprovider -x 2>&1 | tail -4 > gpurun_out/r2b22_tests.log
timeout -s KILL 600 python bench.py --workload implicit_maml --steps 5 --no-cpu-baseline --no-extra > gpurun_out/r2b22_bench.json 2> gpurun_out/r2b22_bench.err
timeout -s KILL 600 python bench.py --workload bert_data_reweighting --steps 5 --no-cpu-baseline --no-extra > gpurun_out/r2b22_bert.json 2> gpurun_out/r2b22_bert.err
timeout -s KILL 200 python tools/probe_tma.py > gpurun_out/r2b22_probe_tma.log 2>&1
tail -2 gpurun_out/r2b22_unit.log | cut -c1-200; cat gpurun_out/r2b22_halobench.log; tail -2 gpurun_out/r2b22_tests.log | cut -c1-200
python -c "
import json
for f in ('r2b22_bench','r2b22_bert'):
    d=json.loads(open('gpurun_out/'+f+'.json').read().strip().splitlines()[-1])
    print(f,'value',d['value'],'e2e',d['e2e']['value'])"
tail -5 gpurun_out/r2b22_probe_tma.log
