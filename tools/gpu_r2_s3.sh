#!/bin/bash
# Round 2, GPU session 3: new parity tests (full-frame CRCs, multi-device behind the C ABI) and e2e with copy overlap.
TAG=${1:-r2s3}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_${TAG}.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu_${TAG}.log
tail -30 gpurun_out/pytest_gpu_${TAG}.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
echo "bench exit $?" >> gpurun_out/bench_${TAG}.err
python -c "
import json; d=json.load(open('gpurun_out/bench_${TAG}.json')); print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e'])"
tail -3 gpurun_out/bench_${TAG}.err
