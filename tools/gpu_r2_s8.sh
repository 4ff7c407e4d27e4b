#!/bin/bash
# Round 2, GPU session 8 (one GPU): the library with the speculated covered-block count; whole GPU suite,
# quick lines, the bench line, and a full ncu of the two largest kernels after k_trace
# (k_shade_miss of bounces 0 and 1, k_shade_hit_tiles) with source lines.
TAG=${1:-r2s8}
mkdir -p gpurun_out
SPB_TIMING_OUT=gpurun_out/timing_${TAG}.txt timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_${TAG}.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu_${TAG}.log
tail -8 gpurun_out/pytest_gpu_${TAG}.log
timeout 200 python bench.py --steps 8 --warmup 3 --quick > gpurun_out/quick_c3_${TAG}.json 2>&1; cut -c1-260 gpurun_out/quick_c3_${TAG}.json
timeout 600 python bench.py --steps 10 --warmup 5 --cpu-seconds 6 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
echo "bench exit $?" >> gpurun_out/bench_${TAG}.err; tail -3 gpurun_out/bench_${TAG}.err; cut -c1-400 gpurun_out/bench_${TAG}.json
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_shade_miss -s 60 -c 2 -f -o gpurun_out/prof_miss_${TAG} \
    python bench.py --steps 1 --warmup 3 --quick > gpurun_out/ncu_full_miss_${TAG}.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_shade_hit_tiles -s 12 -c 1 -f -o gpurun_out/prof_hit_tiles_${TAG} \
    python bench.py --steps 1 --warmup 3 --quick > gpurun_out/ncu_full_hit_tiles_${TAG}.log 2>&1
ls -la gpurun_out/*${TAG}*
