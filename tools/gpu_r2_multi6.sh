#!/bin/bash
# N=8 bench line with the driver's arguments (e2e host phases of rank 0 in the line), then N=2.
TAG=${1:-r2m6}
mkdir -p gpurun_out
for n in 8 2; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 \
      bench.py --gpus $n --steps 20 --warmup 5 --no-secondary > gpurun_out/bench_n${n}_${TAG}.json 2> gpurun_out/bench_n${n}_${TAG}.err
  echo "== N=$n rc $?"
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_n${n}_${TAG}.json'))
    print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', {k:v for k,v in d['e2e'].items() if k!='how'}, 'kernel', [round(r['kernel_ms'],2) for r in d['ranks']])
    print('parity', json.dumps(d.get('parity'))[:200])
except Exception as e: print('no json', e)
PY
  tail -3 gpurun_out/bench_n${n}_${TAG}.err
done
