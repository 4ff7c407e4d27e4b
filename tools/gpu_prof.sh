#!/bin/bash
# Launch list + one full ncu capture of the trace kernels of one pass (primary + every bounce).
# usage (under gpurun): bash tools/gpu_prof.sh <tag> [bench args]
TAG=${1:-x}; shift
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 1 --warmup 3 --quick "$@" > gpurun_out/bench_under_ncu_${TAG}.log 2>&1
# skip the warm-up frames' trace launches, then take one pass: 5 trace launches
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_trace -s ${SKIP:-40} -c ${COUNT:-5} -f -o gpurun_out/prof_${TAG} \
    python bench.py --steps 1 --warmup 3 --quick "$@" > gpurun_out/ncu_full_${TAG}.log 2>&1
tail -3 gpurun_out/ncu_full_${TAG}.log | cut -c1-300
