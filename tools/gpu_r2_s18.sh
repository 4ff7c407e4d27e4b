#!/bin/bash
# Round 2, GPU session 18 (one GPU): occupancy of the shading kernels (k_shade_miss at 7 / 8 CTAs per SM = 36 / 32
# registers, k_shade_hit_tiles at 5 CTAs) on C3; C5 with the scene-dependent thresholds and a few eviction thresholds more.
TAG=${1:-r2s18}
mkdir -p gpurun_out
AB=gpurun_out/ab_${TAG}.txt
: > $AB
run() { echo "== lib[$LIBV] args[$*]" >> $AB; SPB_B200_LIB=$LIBV timeout 200 python bench.py --steps 5 --warmup 3 --quick "$@" 2>&1 | cut -c1-400 >> $AB; }
for v in "" variants/miss8.so variants/miss7.so variants/tiles5.so "" variants/miss8.so; do LIBV=$v; run; done
LIBV=""
run --workload c5 --spp 16
run --workload c5 --spp 16 --evict 20,20
run --workload c5 --spp 16 --evict 24,24
run --workload c5 --spp 16 --evict 16,0
run --workload c5 --spp 16 --evict 20,16
cat $AB | grep -o '== .*\|"ms_per_step": [0-9.]*, "kernel_ms": [0-9.]*, "trace_ms": [0-9.]*'
