#!/bin/bash
# Round 2, GPU session 24 (one GPU): escaped rays shaded by the trace kernel where it retires them (sp_b200_SetMissFusion)
# against the miss queue + k_shade_miss, C3 and C5; the parity tests with the fusion on.
TAG=${1:-r2s24}
mkdir -p gpurun_out
AB=gpurun_out/ab_${TAG}.txt
: > $AB
run() { echo "== args[$*]" >> $AB; timeout 200 python bench.py --steps 5 --warmup 3 --quick "$@" 2>&1 | cut -c1-400 >> $AB; }
run --fuse-miss 0
run --fuse-miss 1
run --fuse-miss 0
run --fuse-miss 1
run --fuse-miss 0 --workload c5 --spp 16
run --fuse-miss 1 --workload c5 --spp 16
cat $AB | grep -o '== .*\|"ms_per_step": [0-9.]*, "kernel_ms": [0-9.]*, "trace_ms": [0-9.]*'
SPB_B200_FUSE_MISS=1 timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_fused_${TAG}.log 2>&1
tail -3 gpurun_out/pytest_gpu_fused_${TAG}.log
