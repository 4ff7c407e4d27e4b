#!/bin/bash
# Round 2, GPU session 2: parity of the second traversal machine, A/B against round 1's kernel and
# occupancy variants (prebuilt by tools/build_variants.sh), ncu of the new kernel, L2 peak re-run.
TAG=${1:-r2s2}
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_${TAG}.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu_${TAG}.log
tail -3 gpurun_out/pytest_gpu_${TAG}.log
OUT=gpurun_out/ab_${TAG}.txt; : > $OUT
for v in default mb4 mb6 old; do
  lib=""; [ "$v" != "default" ] && lib="variants/${v}.so"
  echo "== ${v} c3" >> $OUT
  SPB_B200_LIB=$lib timeout 200 python bench.py --steps 5 --warmup 3 --quick 2>&1 | cut -c1-200 >> $OUT
done
for v in default old; do
  lib=""; [ "$v" != "default" ] && lib="variants/${v}.so"
  echo "== ${v} c5" >> $OUT
  SPB_B200_LIB=$lib timeout 300 python bench.py --steps 3 --warmup 3 --quick --workload c5 --spp 16 2>&1 | cut -c1-200 >> $OUT
done
cat $OUT
timeout 120 tools/l2_peak > gpurun_out/l2_peak_${TAG}.json 2> gpurun_out/l2_peak_${TAG}.err
cat gpurun_out/l2_peak_${TAG}.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_trace -s 60 -c 2 -f -o gpurun_out/prof_${TAG} \
    python bench.py --steps 1 --warmup 3 --quick > gpurun_out/ncu_full_${TAG}.log 2>&1
tail -2 gpurun_out/ncu_full_${TAG}.log | cut -c1-200
ls -la gpurun_out/prof_${TAG}.ncu-rep
