#!/usr/bin/env bash
# dev call: second-generation small-channel conv kernels (conv_small2.cu): plan / parity tests, LeNet A/B, cfg4 line
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
T0=$(date +%s)
timeout 900 python -m pytest tests/test_plan_gpu.py tests/test_parity_gpu.py tests/test_reference_gpu.py -m gpu -q -p no:cacheprovider -k "lenet or fourconv or generic or logistic or mlp or resnet or reweight" 2>&1 | tail -15 > gpurun_out/r2b28_tests.log
echo "tests done $(( $(date +%s) - T0 )) s" >> gpurun_out/r2b28_tests.log
for v in v2 v1 px4 px5 px7; do
  case $v in
    v2) env="";;
    v1) env="BB200_CONV_SMALL_V1=1";;
    px4) env="BB200_CORR2_PX=4";;
    px5) env="BB200_CORR2_PX=5";;
    px7) env="BB200_CORR2_PX=7";;
  esac
  env $env timeout 300 python bench.py --workload learning_to_reweight --no-cpu-baseline --steps 5 --e2e-steps 2 > gpurun_out/r2b28_lenet_$v.json 2> gpurun_out/r2b28_lenet_$v.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2b28_lenet_$v.json").read().strip().splitlines()[-1])
    print("$v", round(d["value"],1), "it/s e2e", round(d["e2e"]["value"],1), [(n["node"][:40], n["ms"]) for n in d["roofline"]["top_nodes"][:5]])
except Exception as e:
    print("$v failed", e)
PY
done > gpurun_out/r2b28_lenet_ab.log 2>&1
echo "lenet A/B done $(( $(date +%s) - T0 )) s" >> gpurun_out/r2b28_tests.log
timeout 600 python bench.py --workload neural_architecture_search --steps 5 --e2e-steps 2 > gpurun_out/r2b28_nas.json 2> gpurun_out/r2b28_nas.err
echo "nas done $(( $(date +%s) - T0 )) s" >> gpurun_out/r2b28_tests.log
cat gpurun_out/r2b28_tests.log | cut -c1-220; cat gpurun_out/r2b28_lenet_ab.log | cut -c1-400; cut -c1-600 gpurun_out/r2b28_nas.json; tail -2 gpurun_out/r2b28_nas.err
