#!/bin/bash
# Round 2, GPU session 9 (one GPU): straggler eviction (EVICT / RESUME launches of k_trace) -- GPU suite, then
# A/B of eviction thresholds against packet mode and per-lane refill on C3 and C5; L2 fetch granularity A/B.
TAG=${1:-r2s9}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_${TAG}.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu_${TAG}.log
tail -8 gpurun_out/pytest_gpu_${TAG}.log
AB=gpurun_out/ab_${TAG}.txt
: > $AB
run() { echo "== env[$ENVS] args[$*]" >> $AB; env $ENVS timeout 200 python bench.py --steps 6 --warmup 3 --quick "$@" 2>&1 | cut -c1-330 >> $AB; }
ENVS=""
run
run --refill 1,1,12
run --refill 1,12,12
run --evict 12,0
run --evict 16,0
run --evict 8,0
run --evict 20,0
run --evict 12,12
run --evict 16,16
run --evict 20,20
run --evict 24,24
run --evict 16,24
ENVS="SPB_B200_L2_FETCH=32" run
ENVS="SPB_B200_L2_FETCH=128" run
ENVS="SPB_B200_L2_FETCH=32" run --evict 16,16
ENVS=""
run --workload c5 --spp 16
run --workload c5 --spp 16 --evict 12,12
run --workload c5 --spp 16 --evict 16,16
run --workload c5 --spp 16 --evict 24,24
run --workload c5 --spp 16 --evict 0,16
cat $AB | grep -o '== .*\|"ms_per_step": [0-9.]*'
