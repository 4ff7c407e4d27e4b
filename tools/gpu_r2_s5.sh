#!/bin/bash
# Round 2, GPU session 5 (one GPU): whole GPU suite on the final traversal machine, A/B on C5 and C3
# (default / round-1 kernel / TLAS staged in shared memory), ncu of the C5 bounce-1 trace launch for the
# default and the staged build (the north star's shared-memory staging, measured).
TAG=${1:-r2s5}
mkdir -p gpurun_out
SPB_TIMING_OUT=gpurun_out/timing_${TAG}.txt timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_${TAG}.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu_${TAG}.log
tail -12 gpurun_out/pytest_gpu_${TAG}.log; cat gpurun_out/timing_${TAG}.txt
OUT=gpurun_out/ab_${TAG}.txt; : > $OUT
for v in default tlas64 old; do
  lib=""; [ "$v" != "default" ] && lib="variants/${v}.so"
  echo "== ${v} c5" >> $OUT
  SPB_B200_LIB=$lib timeout 300 python bench.py --steps 3 --warmup 3 --quick --workload c5 --spp 16 2>&1 | cut -c1-200 >> $OUT
done
for v in default old; do
  lib=""; [ "$v" != "default" ] && lib="variants/${v}.so"
  echo "== ${v} c3" >> $OUT
  SPB_B200_LIB=$lib timeout 200 python bench.py --steps 5 --warmup 3 --quick 2>&1 | cut -c1-200 >> $OUT
done
cat $OUT
for v in default tlas64; do
  lib=""; [ "$v" != "default" ] && lib="variants/${v}.so"
  SPB_B200_LIB=$lib timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_trace -s 19 -c 1 -f -o gpurun_out/prof_c5_${v}_${TAG} \
      python bench.py --steps 1 --warmup 3 --quick --workload c5 --spp 16 > gpurun_out/ncu_c5_${v}_${TAG}.log 2>&1
  tail -1 gpurun_out/ncu_c5_${v}_${TAG}.log | cut -c1-200
done
ls -la gpurun_out/*.ncu-rep | tail -3
