#!/bin/bash
# One GPU session: parity tests, smoke, bench (both arms), launch list and a full ncu capture of
# every traversal-kernel launch of one frame.  Usage (under gpurun): bash tools/gpu_round.sh <tag>
TAG=${1:-r01}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu_${TAG}.txt 2>&1
nproc >> gpurun_out/gpu_${TAG}.txt
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_${TAG}.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu_${TAG}.log
timeout 300 python -c "import __graft_entry__ as e; e.smoke()" > gpurun_out/smoke_${TAG}.log 2>&1
echo "smoke exit $?" >> gpurun_out/smoke_${TAG}.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
echo "bench exit $?" >> gpurun_out/bench_${TAG}.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_${TAG}.json 2>> gpurun_out/bench_${TAG}.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 1 --warmup 3 --quick > gpurun_out/bench_under_ncu_${TAG}.log 2>&1
# one frame = TRACES k_trace launches (bands x 5); skip the three warm-up frames.
# (a) DRAM traffic + duration of every k_trace launch of one frame (small report);
# (b) the full set with source for the first two launches of that frame (primary + bounce 1).
TRACES=${TRACES:-20}
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_bytes.sum,l1tex__t_bytes.sum \
    --clock-control none -k regex:k_trace -s $((TRACES * 3)) -c ${TRACES} -f -o gpurun_out/traffic_${TAG} \
    python bench.py --steps 1 --warmup 3 --quick > gpurun_out/ncu_traffic_${TAG}.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_trace -s $((TRACES * 3)) -c 2 -f -o gpurun_out/prof_${TAG} \
    python bench.py --steps 1 --warmup 3 --quick > gpurun_out/ncu_full_${TAG}.log 2>&1
ls -la gpurun_out
tail -5 gpurun_out/pytest_gpu_${TAG}.log; cat gpurun_out/smoke_${TAG}.log | tail -3; cat gpurun_out/bench_${TAG}.json | cut -c1-300; cat gpurun_out/bench_${TAG}.err | tail -5
