#!/bin/bash
# Round 2, GPU session 15 (one GPU): k_shade_miss prefetch variants on C3 (0 none, 1 ray two items ahead, 2 vertex
# terms, 3 both); four-class vote of the multi-object trace kernel on C5 (vote2 = node / leaf only; exit weights 1, 2, 4).
TAG=${1:-r2s15}
mkdir -p gpurun_out
AB=gpurun_out/ab_${TAG}.txt
: > $AB
run() { echo "== lib[$LIBV] args[$*]" >> $AB; SPB_B200_LIB=$LIBV timeout 200 python bench.py --steps 6 --warmup 3 --quick "$@" 2>&1 | cut -c1-400 >> $AB; }
for v in variants/miss0.so variants/miss1.so variants/miss2.so variants/miss3.so variants/miss0.so variants/miss3.so; do LIBV=$v; run; done
for v in variants/vote2.so "" variants/vote4w1.so variants/vote4w4.so variants/vote2.so ""; do LIBV=$v; run --workload c5 --spp 16; done
cat $AB | grep -o '== .*\|"ms_per_step": [0-9.]*, "kernel_ms": [0-9.]*, "trace_ms": [0-9.]*'
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_${TAG}.log 2>&1
tail -5 gpurun_out/pytest_gpu_${TAG}.log
