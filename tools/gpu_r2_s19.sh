#!/bin/bash
# Round 2, GPU session 19 (one GPU): grids of the shading kernels -- fixed 8 CTAs per SM (grid0), every CTA resident for
# k_shade_miss only (grid1), for the hit kernels too (default) -- on C3 and C5; then the GPU suite on the default build.
TAG=${1:-r2s19}
mkdir -p gpurun_out
AB=gpurun_out/ab_${TAG}.txt
: > $AB
run() { echo "== lib[$LIBV] args[$*]" >> $AB; SPB_B200_LIB=$LIBV timeout 200 python bench.py --steps 5 --warmup 3 --quick "$@" 2>&1 | cut -c1-400 >> $AB; }
for v in variants/grid0.so variants/grid1.so "" variants/grid0.so variants/grid1.so ""; do LIBV=$v; run; done
for v in variants/grid0.so ""; do LIBV=$v; run --workload c5 --spp 16; done
cat $AB | grep -o '== .*\|"ms_per_step": [0-9.]*, "kernel_ms": [0-9.]*, "trace_ms": [0-9.]*'
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_${TAG}.log 2>&1
tail -3 gpurun_out/pytest_gpu_${TAG}.log
