#!/bin/bash
# C5 (instanced 10 M triangle scene): launch list + full ncu of the first two k_trace launches of a frame.
TAG=${1:-c5}
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 1 --warmup 3 --quick --workload c5 --spp 16 > gpurun_out/bench_under_ncu_${TAG}.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_trace -s ${SKIP:-30} -c 2 -f -o gpurun_out/prof_${TAG} \
    python bench.py --steps 1 --warmup 3 --quick --workload c5 --spp 16 > gpurun_out/ncu_full_${TAG}.log 2>&1
tail -2 gpurun_out/ncu_full_${TAG}.log | cut -c1-200
