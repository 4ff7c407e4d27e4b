#!/bin/bash
# Round 2, GPU session 20 (one GPU): the final library -- bench line, launch list of the bench command, DRAM traffic of
# every k_trace launch of a frame (C3 and C5), full ncu of the primary / EVICT / RESUME launches of a C3 pass and of
# the bounce-1 EVICT / RESUME launches of a C5 pass.  (GPU suite: session 19, same library.)
TAG=${1:-r2s20}
mkdir -p gpurun_out
timeout 200 python bench.py --steps 8 --warmup 3 --quick > gpurun_out/quick_c3_${TAG}.json 2>&1; cut -c1-330 gpurun_out/quick_c3_${TAG}.json
timeout 300 python bench.py --steps 3 --warmup 3 --quick --workload c5 --spp 16 > gpurun_out/quick_c5_${TAG}.json 2>&1; cut -c1-330 gpurun_out/quick_c5_${TAG}.json
timeout 900 python bench.py --steps 10 --warmup 5 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
echo "bench exit $?" >> gpurun_out/bench_${TAG}.err; tail -3 gpurun_out/bench_${TAG}.err; cut -c1-400 gpurun_out/bench_${TAG}.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference_${TAG}.json 2> gpurun_out/bench_reference_${TAG}.err
cut -c1-400 gpurun_out/bench_reference_${TAG}.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 1 --warmup 3 --quick > gpurun_out/bench_under_ncu_${TAG}.log 2>&1
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_bytes.sum,l1tex__t_bytes.sum \
    --clock-control none -k regex:k_trace -s 72 -c 24 -f -o gpurun_out/traffic_c3_${TAG} \
    python bench.py --steps 1 --warmup 3 --quick > gpurun_out/ncu_traffic_c3_${TAG}.log 2>&1
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_bytes.sum,l1tex__t_bytes.sum \
    --clock-control none -k regex:k_trace -s 81 -c 27 -f -o gpurun_out/traffic_c5_${TAG} \
    python bench.py --steps 1 --warmup 3 --quick --workload c5 --spp 16 > gpurun_out/ncu_traffic_c5_${TAG}.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_trace -s 72 -c 3 -f -o gpurun_out/prof_c3_${TAG} \
    python bench.py --steps 1 --warmup 3 --quick > gpurun_out/ncu_full_c3_${TAG}.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_trace -s 82 -c 2 -f -o gpurun_out/prof_c5_${TAG} \
    python bench.py --steps 1 --warmup 3 --quick --workload c5 --spp 16 > gpurun_out/ncu_full_c5_${TAG}.log 2>&1
ls -la gpurun_out/*${TAG}*
