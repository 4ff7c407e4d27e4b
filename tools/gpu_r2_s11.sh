#!/bin/bash
# Round 2, GPU session 11 (one GPU): hybrid shared-memory stack of the trace kernel (variants/smemK.so) against the
# default build on C3 and C5; parity tests on the variant; full ncu of the EVICT launch with the variant.
TAG=${1:-r2s11}
mkdir -p gpurun_out
AB=gpurun_out/ab_${TAG}.txt
: > $AB
run() { echo "== lib[$LIBV] args[$*]" >> $AB; SPB_B200_LIB=$LIBV timeout 200 python bench.py --steps 6 --warmup 3 --quick "$@" 2>&1 | cut -c1-330 >> $AB; }
for v in "" variants/smem8.so variants/smem12.so variants/smem6.so variants/smem8early.so; do LIBV=$v; run; done
for v in "" variants/smem8.so variants/smem12.so variants/smem8early.so; do LIBV=$v; run --workload c5 --spp 16; done
for v in "" variants/smem8.so; do LIBV=$v; run --evict 0,0; done
cat $AB | grep -o '== .*\|"ms_per_step": [0-9.]*'
SPB_B200_LIB=variants/smem8.so timeout 600 python -m pytest tests -m gpu -q -x -k "golden or coverage or c5_instanced_scene or five_bounces or multi_object" > gpurun_out/pytest_smem8_${TAG}.log 2>&1
tail -3 gpurun_out/pytest_smem8_${TAG}.log
SPB_B200_LIB=variants/smem8.so timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_trace -s 73 -c 2 -f -o gpurun_out/prof_c3_smem8_${TAG} \
    python bench.py --steps 1 --warmup 3 --quick > gpurun_out/ncu_full_c3_smem8_${TAG}.log 2>&1
ls -la gpurun_out/*${TAG}*
