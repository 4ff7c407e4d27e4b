#!/bin/bash
# N=8 bench line with the driver's arguments: two frames in flight in both timed loops, host barrier in the e2e steps.
TAG=${1:-r2m8}
mkdir -p gpurun_out
n=8
o=gpurun_out/bench_n${n}_${TAG}
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus $n --steps 20 --warmup 5 --no-secondary > $o.json 2> $o.err
echo "== N=$n rc $?"
python - <<PY
import json
try:
    d=json.load(open('$o.json'))
    print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', {k:v for k,v in d['e2e'].items() if k!='how'}, 'kernel', [round(r['kernel_ms'],2) for r in d['ranks']])
    print('parity', json.dumps(d.get('parity'))[:200])
except Exception as e: print('no json', e)
PY
tail -3 $o.err
