#!/bin/bash
# Round 2, GPU session 1 (one GPU): state of the round-1 library on this round's box, the on-chip
# peaks re-measured (tools/l2_peak.cu sweep), compute-sanitizer over the smoke scene and a C5
# thumbnail, strip-granularity A/B and the C5 frame.   usage (under gpurun): bash tools/gpu_r2_s1.sh <tag>
TAG=${1:-r2s1}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu_${TAG}.txt 2>&1
nproc >> gpurun_out/gpu_${TAG}.txt
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_${TAG}.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu_${TAG}.log
tail -3 gpurun_out/pytest_gpu_${TAG}.log
timeout 120 tools/l2_peak > gpurun_out/l2_peak_${TAG}.json 2> gpurun_out/l2_peak_${TAG}.err
cat gpurun_out/l2_peak_${TAG}.json
for tool in memcheck racecheck initcheck; do
  for part in smoke c5; do
    timeout 420 compute-sanitizer --tool ${tool} --print-limit 20 python tools/sanitize_run.py ${part} \
        > gpurun_out/sanitizer_${tool}_${part}_${TAG}.log 2>&1
    echo "exit $?" >> gpurun_out/sanitizer_${tool}_${part}_${TAG}.log
    echo "== ${tool} ${part}"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|exit |done" gpurun_out/sanitizer_${tool}_${part}_${TAG}.log | tail -4
  done
done
for rows in 16 8; do
  echo "== strip rows ${rows}" >> gpurun_out/ab_strip_rows_${TAG}.txt
  timeout 120 python bench.py --steps 5 --warmup 3 --quick --strip-rows ${rows} 2>&1 | cut -c1-200 >> gpurun_out/ab_strip_rows_${TAG}.txt
done
cat gpurun_out/ab_strip_rows_${TAG}.txt
timeout 300 python bench.py --steps 3 --warmup 3 --quick --workload c5 --spp 16 2>&1 | cut -c1-300 > gpurun_out/bench_c5_${TAG}.txt
cat gpurun_out/bench_c5_${TAG}.txt
