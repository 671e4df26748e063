#!/usr/bin/env bash
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
timeout 600 python tools/debug_bn_parity.py > gpurun_out/r2b27_debug.log 2>&1
tail -60 gpurun_out/r2b27_debug.log | cut -c1-200
