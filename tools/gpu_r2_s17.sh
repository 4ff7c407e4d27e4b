#!/bin/bash
# Round 2, GPU session 17 (one GPU): the "fewest idle lane-slots" vote (variants/minwaste.so) against the majority vote
# on C3 and C5; C5 primary refill threshold and eviction thresholds with the four-class vote.
TAG=${1:-r2s17}
mkdir -p gpurun_out
AB=gpurun_out/ab_${TAG}.txt
: > $AB
run() { echo "== lib[$LIBV] args[$*]" >> $AB; SPB_B200_LIB=$LIBV timeout 200 python bench.py --steps 5 --warmup 3 --quick "$@" 2>&1 | cut -c1-400 >> $AB; }
for v in "" variants/minwaste.so "" variants/minwaste.so; do LIBV=$v; run; done
for v in "" variants/minwaste.so; do LIBV=$v; run --workload c5 --spp 16; done
LIBV=""
run --workload c5 --spp 16 --refill 12,0,0
run --workload c5 --spp 16 --refill 16,0,0
run --workload c5 --spp 16 --refill 1,16,0
run --workload c5 --spp 16 --evict 16,16
run --workload c5 --spp 16 --evict 12,0
run --workload c5 --spp 16 --evict 0,12
cat $AB | grep -o '== .*\|"ms_per_step": [0-9.]*, "kernel_ms": [0-9.]*, "trace_ms": [0-9.]*'
