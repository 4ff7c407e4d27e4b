#!/bin/bash
# Round 2, GPU session 16 (one GPU): the pipelined-frames test repeated with its messages visible (one abort in session
# 15), the GPU suite, then C5 with the four-class vote: refill / eviction thresholds, and the variant whose inner loop
# counts lanes about to leave an object as busy.
TAG=${1:-r2s16}
mkdir -p gpurun_out
for i in 1 2 3 4; do timeout 120 python -m pytest tests -m gpu -q -x -s -k pipelined_frames 2>&1 | tail -4; done > gpurun_out/pytest_pipelined_${TAG}.log 2>&1
grep -c passed gpurun_out/pytest_pipelined_${TAG}.log; grep "sp_b200\|rror\|Abort" gpurun_out/pytest_pipelined_${TAG}.log | head -5
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_${TAG}.log 2>&1
tail -3 gpurun_out/pytest_gpu_${TAG}.log
AB=gpurun_out/ab_${TAG}.txt
: > $AB
run() { echo "== lib[$LIBV] args[$*]" >> $AB; SPB_B200_LIB=$LIBV timeout 200 python bench.py --steps 5 --warmup 3 --quick --workload c5 --spp 16 "$@" 2>&1 | cut -c1-400 >> $AB; }
LIBV=""
run
run --refill 1,0,8
run --refill 1,0,16
run --refill 1,0,20
run --evict 0,0
run --evict 8,8
run --evict 12,12
LIBV=variants/busyexit.so
run
run --refill 1,0,8
run --refill 1,0,16
run --evict 0,0
cat $AB | grep -o '== .*\|"ms_per_step": [0-9.]*, "kernel_ms": [0-9.]*, "trace_ms": [0-9.]*'
timeout 100 python bench.py --steps 6 --warmup 3 --quick 2>&1 | grep -o '"ms_per_step": [0-9.]*, "kernel_ms": [0-9.]*, "trace_ms": [0-9.]*'
