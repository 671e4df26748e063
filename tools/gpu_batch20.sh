#!/usr/bin/env bash
# two GPUs: the NCCL DDP sync test and the N=2 bench (replicas, weak scaling)
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_reference_gpu.py -q -m gpu -p no:cacheprovider -k "nccl" 2>&1 | tail -4 > gpurun_out/r2b20_nccl.log
timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/r2b20_bench2.json 2> gpurun_out/r2b20_bench2.err
tail -3 gpurun_out/r2b20_nccl.log; cut -c1-700 gpurun_out/r2b20_bench2.json; tail -3 gpurun_out/r2b20_bench2.err
