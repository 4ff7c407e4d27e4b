#!/bin/bash
# Round 2, GPU session 22 (one GPU): device-side 4-wide collapse of the LBVH builder: its GPU test with builder timings,
# the same timings with the collapse on the host (round 1's path), compute-sanitizer over a device-built C5 thumbnail.
TAG=${1:-r2s22}
mkdir -p gpurun_out
SPB_TIMING_OUT=gpurun_out/timing_lbvh_device_collapse_${TAG}.txt timeout 600 python -m pytest tests -m gpu -q -x -k "lbvh" 2>&1 | tail -3
SPB_B200_LBVH_HOST_COLLAPSE=1 SPB_TIMING_OUT=gpurun_out/timing_lbvh_host_collapse_${TAG}.txt timeout 600 python -m pytest tests -m gpu -q -x -k "lbvh" 2>&1 | tail -3
cat gpurun_out/timing_lbvh_device_collapse_${TAG}.txt; echo; cat gpurun_out/timing_lbvh_host_collapse_${TAG}.txt
timeout 900 compute-sanitizer --tool memcheck python tools/sanitize_run.py c5 > gpurun_out/sanitizer_memcheck_c5_${TAG}.log 2>&1; tail -4 gpurun_out/sanitizer_memcheck_c5_${TAG}.log
timeout 900 compute-sanitizer --tool racecheck python tools/sanitize_run.py c5 > gpurun_out/sanitizer_racecheck_c5_${TAG}.log 2>&1; tail -4 gpurun_out/sanitizer_racecheck_c5_${TAG}.log
