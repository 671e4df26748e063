#!/usr/bin/env bash
# dev call: small-plane channel sum, SIMT GEMM tile-size threshold A/B on LeNet
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_plan_gpu.py tests/test_parity_gpu.py -m gpu -q -p no:cacheprovider -k "lenet or fourconv or generic or mlp or resnet or reweight" 2>&1 | tail -4 > gpurun_out/r2b32_tests.log
for mid in 296 148 64 0; do
  BB200_GEMM_MID=$mid timeout 300 python bench.py --workload learning_to_reweight --no-cpu-baseline --steps 5 --e2e-steps 3 > gpurun_out/r2b32_lenet_$mid.json 2> gpurun_out/r2b32_lenet_$mid.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2b32_lenet_$mid.json").read().strip().splitlines()[-1])
    print("mid<$mid", round(d["value"],1), "it/s e2e", round(d["e2e"]["value"],1), [(n["node"][:34], n["ms"]) for n in d["roofline"]["top_nodes"][:6]])
except Exception as e:
    print("$mid failed", e)
PY
done > gpurun_out/r2b32_ab.log 2>&1
cat gpurun_out/r2b32_tests.log gpurun_out/r2b32_ab.log | cut -c1-420
