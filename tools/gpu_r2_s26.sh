#!/bin/bash
# Round 2, GPU session 26 (one GPU): fused miss shading mode 2 (vertex terms prefetched at ray start) against mode 1; the
# GPU suite on the final library (fusion on by scene: every single-object test runs it), the final bench line.
TAG=${1:-r2s26}
mkdir -p gpurun_out
AB=gpurun_out/ab_${TAG}.txt
: > $AB
run() { echo "== args[$*]" >> $AB; timeout 200 python bench.py --steps 5 --warmup 3 --quick "$@" 2>&1 | cut -c1-400 >> $AB; }
run --fuse-miss 1
run --fuse-miss 2
run --fuse-miss 1
run --fuse-miss 2
cat $AB | grep -o '== .*\|"ms_per_step": [0-9.]*, "kernel_ms": [0-9.]*, "trace_ms": [0-9.]*'
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_${TAG}.log 2>&1
tail -3 gpurun_out/pytest_gpu_${TAG}.log
SPB_B200_FUSE_MISS=2 timeout 600 python -m pytest tests -m gpu -q -x -k "golden or five_bounces or wavefront or coverage or pipelined" > gpurun_out/pytest_gpu_mode2_${TAG}.log 2>&1
tail -2 gpurun_out/pytest_gpu_mode2_${TAG}.log
timeout 900 python bench.py --steps 10 --warmup 5 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
echo "bench exit $?"; cut -c1-330 gpurun_out/bench_${TAG}.json
