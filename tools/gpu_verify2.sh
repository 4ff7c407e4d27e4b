#!/bin/bash
# Last short GPU session: the whole GPU suite on the final library (with builder timings), then a bench line.
TAG=${1:-r03c}
mkdir -p gpurun_out
SPB_TIMING_OUT=gpurun_out/timing_${TAG}.txt timeout 70 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_${TAG}.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu_${TAG}.log
tail -15 gpurun_out/pytest_gpu_${TAG}.log; cat gpurun_out/timing_${TAG}.txt
timeout 40 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
echo "bench exit $?" >> gpurun_out/bench_${TAG}.err
cut -c1-160 gpurun_out/bench_${TAG}.json; tail -2 gpurun_out/bench_${TAG}.err
