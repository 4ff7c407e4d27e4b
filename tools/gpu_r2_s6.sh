#!/bin/bash
# Round 2, GPU session 6 (one GPU): first machine + single-object entry at ray start (variant "oldfused")
# against the second machine and round 1's kernel; parity of that variant; ncu of a C5 bounce-1 launch.
TAG=${1:-r2s6}
mkdir -p gpurun_out
OUT=gpurun_out/ab_${TAG}.txt; : > $OUT
for v in oldfused default old oldfused; do
  lib=""; [ "$v" != "default" ] && lib="variants/${v}.so"
  echo "== ${v} c3" >> $OUT
  SPB_B200_LIB=$lib timeout 200 python bench.py --steps 8 --warmup 3 --quick 2>&1 | cut -c1-140 >> $OUT
done
cat $OUT
SPB_B200_LIB=variants/oldfused.so timeout 600 python -m pytest tests -m gpu -x -q -k "not multi_device and not reference_unit" > gpurun_out/pytest_gpu_oldfused_${TAG}.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu_oldfused_${TAG}.log
tail -4 gpurun_out/pytest_gpu_oldfused_${TAG}.log
timeout 600 python -m pytest tests -m gpu -x -q -k "slab_kats or metrics_mesh or reference_unit or c5_instanced_scene_1080p or multi_device" > gpurun_out/pytest_gpu_new_${TAG}.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu_new_${TAG}.log
tail -6 gpurun_out/pytest_gpu_new_${TAG}.log
for v in old default; do
  lib=""; [ "$v" != "default" ] && lib="variants/${v}.so"
  SPB_B200_LIB=$lib timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_trace -s 46 -c 1 -f -o gpurun_out/prof_c5_${v}_${TAG} \
      python bench.py --steps 1 --warmup 3 --quick --workload c5 --spp 16 > gpurun_out/ncu_c5_${v}_${TAG}.log 2>&1
  tail -1 gpurun_out/ncu_c5_${v}_${TAG}.log | cut -c1-200
done
