#!/bin/bash
# N=4: e2e steps with NCCL limited to 2 / 4 channels (the environment map's all-gather runs beside the kernels).
TAG=${1:-r2m10}
mkdir -p gpurun_out
n=4
for c in 2 4; do
  o=gpurun_out/bench_n${n}_ch${c}_${TAG}
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 \
      bench.py --gpus $n --steps 20 --warmup 5 --no-secondary --no-parity --nccl-channels $c > $o.json 2> $o.err
  echo "== N=$n channels $c rc $?"
  python - <<PY
import json
try:
    d=json.load(open('$o.json'))
    e=d['e2e']
    print('value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'e2e ms', round(e['ms_per_step'],3), 'gather', [round(r['gather_ms'],2) for r in d['ranks']])
    print('e2e ranks', e.get('ranks'))
except Exception as ex: print('no json', ex)
PY
  tail -2 $o.err
done
