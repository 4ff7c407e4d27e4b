#!/bin/bash
# Round 2, GPU session 10 (one GPU): early node fetch (variants/early1.so) against the default build on C3 and C5,
# parity tests on the variant, full ncu of the primary / EVICT / RESUME launches of a C3 frame (default build).
TAG=${1:-r2s10}
mkdir -p gpurun_out
AB=gpurun_out/ab_${TAG}.txt
: > $AB
run() { echo "== lib[$LIBV] args[$*]" >> $AB; SPB_B200_LIB=$LIBV timeout 200 python bench.py --steps 6 --warmup 3 --quick "$@" 2>&1 | cut -c1-330 >> $AB; }
LIBV=""; run
LIBV=variants/early1.so; run
LIBV=""; run --evict 0,0
LIBV=variants/early1.so; run --evict 0,0
LIBV=""; run --workload c5 --spp 16
LIBV=variants/early1.so; run --workload c5 --spp 16
LIBV=""; run
LIBV=variants/early1.so; run
cat $AB | grep -o '== .*\|"ms_per_step": [0-9.]*'
SPB_B200_LIB=variants/early1.so timeout 600 python -m pytest tests -m gpu -q -x -k "golden or coverage or c5_instanced_scene or five_bounces or multi_object" > gpurun_out/pytest_early1_${TAG}.log 2>&1
tail -3 gpurun_out/pytest_early1_${TAG}.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_trace -s 72 -c 3 -f -o gpurun_out/prof_c3_${TAG} \
    python bench.py --steps 1 --warmup 3 --quick > gpurun_out/ncu_full_c3_${TAG}.log 2>&1
SPB_B200_LIB=variants/early1.so timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_trace -s 72 -c 2 -f -o gpurun_out/prof_c3_early1_${TAG} \
    python bench.py --steps 1 --warmup 3 --quick > gpurun_out/ncu_full_c3_early1_${TAG}.log 2>&1
ls -la gpurun_out/*${TAG}*
