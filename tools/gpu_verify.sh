#!/bin/bash
# Short GPU session: parity tests (with baker timings), smoke, ncu capture of the environment-baking
# kernels, then a bench line.  Usage (under gpurun): bash tools/gpu_verify.sh <tag>
TAG=${1:-r03}
mkdir -p gpurun_out
SPB_TIMING_OUT=gpurun_out/cubemap_timing_${TAG}.txt timeout 110 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_${TAG}.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu_${TAG}.log
timeout 40 python -c "import __graft_entry__ as e; e.smoke()" > gpurun_out/smoke_${TAG}.log 2>&1
echo "smoke exit $?" >> gpurun_out/smoke_${TAG}.log
timeout 60 ncu --set full --clock-control none --import-source on -k regex:'k_cube_map|k_irradiance' -f \
    -o gpurun_out/prof_cubemap_${TAG} python tools/cubemap_profile.py > gpurun_out/ncu_cubemap_${TAG}.log 2>&1
timeout 90 python bench.py --steps 5 --warmup 3 --cpu-seconds 5 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
echo "bench exit $?" >> gpurun_out/bench_${TAG}.err
tail -3 gpurun_out/pytest_gpu_${TAG}.log; tail -2 gpurun_out/smoke_${TAG}.log; cat gpurun_out/cubemap_timing_${TAG}.txt
tail -2 gpurun_out/ncu_cubemap_${TAG}.log; cut -c1-200 gpurun_out/bench_${TAG}.json; tail -2 gpurun_out/bench_${TAG}.err
