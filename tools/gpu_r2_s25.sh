#!/bin/bash
# Round 2, GPU session 25 (one GPU): fused miss shading, mode 2 (vertex terms prefetched when the ray starts) against
# modes 1 and 0 on C3; the default (by scene) on C3 and C5.
TAG=${1:-r2s25}
mkdir -p gpurun_out
AB=gpurun_out/ab_${TAG}.txt
: > $AB
run() { echo "== args[$*]" >> $AB; timeout 200 python bench.py --steps 5 --warmup 3 --quick "$@" 2>&1 | cut -c1-400 >> $AB; }
run --fuse-miss 0
run --fuse-miss 1
run --fuse-miss 2
run --fuse-miss 1
run --fuse-miss 2
run
run --workload c5 --spp 16
cat $AB | grep -o '== .*\|"ms_per_step": [0-9.]*, "kernel_ms": [0-9.]*, "trace_ms": [0-9.]*'
