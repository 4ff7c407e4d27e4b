#!/bin/bash
# Round 2, GPU session 28 (one GPU): eviction thresholds and paths per pass re-checked on the final single-object kernel (C3);
# the pipelined-frames test with its new cases.
TAG=${1:-r2s28}
mkdir -p gpurun_out
AB=gpurun_out/ab_${TAG}.txt
: > $AB
run() { echo "== args[$*]" >> $AB; timeout 200 python bench.py --steps 5 --warmup 3 --quick "$@" 2>&1 | cut -c1-400 >> $AB; }
run
run --evict 6,0
run --evict 10,0
run --evict 12,0
run --evict 8,8
run --evict 8,4
run --paths-per-pass 16777216
run --paths-per-pass 50331648
run --paths-per-pass 8388608
run
cat $AB | grep -o '== .*\|"ms_per_step": [0-9.]*, "kernel_ms": [0-9.]*, "trace_ms": [0-9.]*'
timeout 300 python -m pytest tests -m gpu -q -x -k "pipelined" 2>&1 | tail -2
