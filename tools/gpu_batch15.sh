#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
{
for v in n1 n2; do
BB200_LIB=/root/repo/.ab/lib_$v.so timeout 300 python tools/halo_bench.py
done
} > gpurun_out/r2b15.log 2>&1
tail -60 gpurun_out/r2b15.log
