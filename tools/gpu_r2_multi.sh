#!/bin/bash
# Multi-GPU bench under torchrun (one process per GPU).  usage (under gpurun --gpus N): bash tools/gpu_r2_multi.sh <tag> <N> [bench args]
TAG=${1:-r2m}; N=${2:-2}; shift; shift
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/gpus_${TAG}.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node ${N} --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus ${N} --steps 10 --warmup 5 "$@" > gpurun_out/bench_n${N}_${TAG}.json 2> gpurun_out/bench_n${N}_${TAG}.err
echo "bench exit $?" >> gpurun_out/bench_n${N}_${TAG}.err
grep -v "^\[W\|^W0\|NCCL\|\*\*\*" gpurun_out/bench_n${N}_${TAG}.err | tail -15
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_n${N}_${TAG}.json'))
    for k in ('value','ms_per_step','e2e','ranks','strips'):
        print(k, json.dumps(d.get(k))[:700])
    print('parity', json.dumps(d.get('parity'))[:600])
    print('c5', json.dumps((d.get('secondary') or {}).get('c5'))[:500])
except Exception as e:
    print('no json', e)
PY
