#!/usr/bin/env bash
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
for i in 1 2 3; do
  timeout 300 python bench.py --workload implicit_maml --no-cpu-baseline --no-extra --steps 3 --e2e-steps 8 > gpurun_out/r2b31_maml_$i.json 2> gpurun_out/r2b31_maml_$i.err
  python -c "
import json; d=json.loads(open('gpurun_out/r2b31_maml_$i.json').read().strip().splitlines()[-1]); print(round(d['value'],1), round(d['e2e']['value'],1), d['e2e']['step_ms'])"
done
