#!/bin/bash
# Round 2, GPU session 21 (one GPU): session 20 again with what comes back kept under gpurun's 64 MiB (its four
# reports were 116 MB and none of its files returned): reports are summarised ON the box (tools/ncu_summary.py,
# ncu_lines.py, ncu_traffic.py) and xz-compressed; the traffic reports are dropped after their JSON is written.
TAG=${1:-r2s21}
mkdir -p gpurun_out
O=gpurun_out
timeout 200 python bench.py --steps 8 --warmup 3 --quick > $O/quick_c3_${TAG}.json 2>&1; cut -c1-330 $O/quick_c3_${TAG}.json
timeout 300 python bench.py --steps 3 --warmup 3 --quick --workload c5 --spp 16 > $O/quick_c5_${TAG}.json 2>&1; cut -c1-330 $O/quick_c5_${TAG}.json
C3RAYS=$(python -c "import json;print(json.load(open('$O/quick_c3_${TAG}.json'))['traced_rays_per_step'])")
C5RAYS=$(python -c "import json;print(json.load(open('$O/quick_c5_${TAG}.json'))['traced_rays_per_step'])")
timeout 900 python bench.py --steps 10 --warmup 5 > $O/bench_${TAG}.json 2> $O/bench_${TAG}.err
echo "bench exit $?" >> $O/bench_${TAG}.err; tail -3 $O/bench_${TAG}.err; cut -c1-300 $O/bench_${TAG}.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference_${TAG}.json 2> $O/bench_reference_${TAG}.err
cut -c1-200 $O/bench_reference_${TAG}.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_${TAG}.csv \
    python bench.py --steps 1 --warmup 3 --quick > $O/bench_under_ncu_${TAG}.log 2>&1
python tools/launch_summary.py $O/launches_${TAG}.csv > $O/launches_${TAG}.txt 2>&1; head -12 $O/launches_${TAG}.txt
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_bytes.sum,l1tex__t_bytes.sum \
    --clock-control none -k regex:k_trace -s 72 -c 24 -f -o /tmp/traffic_c3 \
    python bench.py --steps 1 --warmup 3 --quick > $O/ncu_traffic_c3_${TAG}.log 2>&1
python tools/ncu_traffic.py /tmp/traffic_c3.ncu-rep c3 64 3840 $O/trace_traffic_c3_${TAG}.json $C3RAYS
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_bytes.sum,l1tex__t_bytes.sum \
    --clock-control none -k regex:k_trace -s 81 -c 27 -f -o /tmp/traffic_c5 \
    python bench.py --steps 1 --warmup 3 --quick --workload c5 --spp 16 > $O/ncu_traffic_c5_${TAG}.log 2>&1
python tools/ncu_traffic.py /tmp/traffic_c5.ncu-rep c5 16 3840 $O/trace_traffic_c5_${TAG}.json $C5RAYS
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_trace -s 72 -c 3 -f -o /tmp/prof_c3 \
    python bench.py --steps 1 --warmup 3 --quick > $O/ncu_full_c3_${TAG}.log 2>&1
(python tools/ncu_summary.py /tmp/prof_c3.ncu-rep; python tools/ncu_lines.py /tmp/prof_c3.ncu-rep k_trace 45) > $O/trace_c3_ncu_full_${TAG}.txt 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_trace -s 82 -c 2 -f -o /tmp/prof_c5 \
    python bench.py --steps 1 --warmup 3 --quick --workload c5 --spp 16 > $O/ncu_full_c5_${TAG}.log 2>&1
(python tools/ncu_summary.py /tmp/prof_c5.ncu-rep; python tools/ncu_lines.py /tmp/prof_c5.ncu-rep k_trace 45) > $O/trace_c5_ncu_full_${TAG}.txt 2>&1
xz -9 -T0 -c /tmp/prof_c3.ncu-rep > $O/prof_c3_${TAG}.ncu-rep.xz
xz -9 -T0 -c /tmp/prof_c5.ncu-rep > $O/prof_c5_${TAG}.ncu-rep.xz
du -sh $O; ls -la $O
