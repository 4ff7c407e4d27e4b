#!/bin/bash
# The bench line with the driver's arguments at N GPUs (final library), C5 secondary included.
N=${1:-8}
TAG=${2:-r2m9}
mkdir -p gpurun_out
o=gpurun_out/bench_n${N}_${TAG}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus $N --steps 20 --warmup 5 > $o.json 2> $o.err
echo "== N=$N rc $?"
python - <<PY
import json
try:
    d=json.load(open('$o.json'))
    e={k:v for k,v in d['e2e'].items() if k not in ('how','ranks')}
    print('value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'e2e', e)
    print('kernel', [round(r['kernel_ms'],2) for r in d['ranks']])
    print('e2e ranks', d['e2e'].get('ranks'))
    print('c5', json.dumps(d.get('secondary',{}).get('c5'))[:400])
    print('parity', json.dumps(d.get('parity'))[:200])
except Exception as ex: print('no json', ex)
PY
tail -3 $o.err
