#!/bin/bash
# First GPU session of the next round (one GPU): everything the last round could not measure.
# Usage (under gpurun): bash tools/gpu_next_round.sh <tag>          (about 4 minutes of box time)
#   1. the standard round (tests, smoke, both bench arms, launch list, traffic, full ncu of k_trace)
#   2. baker / builder timings and the bakers' ncu summary
#   3. strip granularity A/B on one GPU (tileHeight 16 vs 8 must not change the frame time)
# Multi-GPU follow-up (separate call, gpurun --gpus 8):
#   python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
#       bench.py --gpus 8 --steps 10 --warmup 8 [--strip-rows 8]
TAG=${1:-r04}
bash tools/gpu_round.sh ${TAG}
SPB_TIMING_OUT=gpurun_out/timing_${TAG}.txt timeout 120 python -m pytest tests -m gpu -q \
    -k "cube_map or lbvh or work_queue" > gpurun_out/pytest_new_${TAG}.log 2>&1
timeout 60 python tools/cubemap_profile.py --time > gpurun_out/cubemap_timing_${TAG}.json 2>&1
timeout 120 ncu --set full --clock-control none --import-source on -k regex:'k_cube_map|k_irradiance|k_lbvh' -f \
    -o gpurun_out/prof_cubemap_${TAG} python tools/cubemap_profile.py > gpurun_out/ncu_cubemap_${TAG}.log 2>&1
for rows in 16 8; do
    timeout 120 python bench.py --steps 5 --warmup 3 --quick --strip-rows ${rows} >> gpurun_out/ab_strip_rows_${TAG}.txt 2>&1
done
tools/l2_peak > gpurun_out/l2_peak_${TAG}.json 2>&1
tail -3 gpurun_out/pytest_new_${TAG}.log; cat gpurun_out/timing_${TAG}.txt gpurun_out/ab_strip_rows_${TAG}.txt
