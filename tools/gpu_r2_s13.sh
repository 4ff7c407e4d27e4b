#!/bin/bash
# Round 2, GPU session 13 (one GPU): bench with two frames in flight (sp_b200_RenderRowsBegin / End) against one call
# per step: device-resident value and e2e.
TAG=${1:-r2s13}
mkdir -p gpurun_out
for f in "" "--no-pipeline"; do
  timeout 300 python bench.py --steps 10 --warmup 5 --no-parity --no-secondary --no-cpu-baseline $f > gpurun_out/bench_${TAG}${f:+_nopipe}.json 2> gpurun_out/bench_${TAG}${f:+_nopipe}.err
  echo "== [$f] rc $?"
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_${TAG}${f:+_nopipe}.json'))
    print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', {k:v for k,v in d['e2e'].items() if k!='how'})
except Exception as e: print('no json', e)
PY
  tail -3 gpurun_out/bench_${TAG}${f:+_nopipe}.err
done
timeout 300 python -m pytest tests -m gpu -q -x -k "pipelined or golden or texture or flush" 2>&1 | tail -3
