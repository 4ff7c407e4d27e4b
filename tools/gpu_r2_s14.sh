#!/bin/bash
# Round 2, GPU session 14 (one GPU): k_trace compiled for single-object scenes (SINGLE) + k_shade_miss run two items
# ahead (L2 prefetch) against session 13's 58.9 ms; C5 unchanged?; the GPU suite.
TAG=${1:-r2s14}
mkdir -p gpurun_out
AB=gpurun_out/ab_${TAG}.txt
: > $AB
run() { echo "== args[$*]" >> $AB; timeout 200 python bench.py --steps 6 --warmup 3 --quick "$@" 2>&1 | cut -c1-400 >> $AB; }
run
run --no-pipeline
run --workload c5 --spp 16
run
cat $AB | grep -o '== .*\|"ms_per_step": [0-9.]*, "kernel_ms": [0-9.]*, "trace_ms": [0-9.]*'
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_${TAG}.log 2>&1
tail -5 gpurun_out/pytest_gpu_${TAG}.log
