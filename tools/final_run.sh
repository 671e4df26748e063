export PYTHONUNBUFFERED=1
timeout -s KILL 900 python -m pytest tests -q -m gpu -p no:cacheprovider 2>&1 | tail -4
timeout -s KILL 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout -s KILL 400 python bench.py > gpurun_out/FINAL_default.json 2> gpurun_out/FINAL_default.err; tail -c 300 gpurun_out/FINAL_default.json
timeout -s KILL 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/FINAL_default_ref.json 2> gpurun_out/FINAL_default_ref.err; tail -c 300 gpurun_out/FINAL_default_ref.json
for w in bert_data_reweighting implicit_maml; do timeout -s KILL 500 python bench.py --workload $w --steps 5 --warmup 3 > gpurun_out/FINAL_$w.json 2> gpurun_out/FINAL_$w.err; tail -c 200 gpurun_out/FINAL_$w.json; done
for w in implicit_maml_n25 neural_architecture_search logistic_regression_hpo; do timeout -s KILL 300 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/FINAL_$w.json 2> gpurun_out/FINAL_$w.err; done
timeout -s KILL 300 compute-sanitizer --tool memcheck --print-limit 3 python -m pytest tests/test_gemm_tma_gpu.py tests/test_plan_gpu.py -q -p no:cacheprovider -k "shape2-layout1 or shape5-layout3 or fourconv_mini_bf16_tc or roberta_bf16_attn_tc or declines" > gpurun_out/sanitizer_tma.log 2>&1; grep -E "passed|failed|ERROR SUMMARY" gpurun_out/sanitizer_tma.log | tail -3
