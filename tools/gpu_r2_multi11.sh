#!/bin/bash
# N=8: strip granularity 4 rows against 8 (the cut's granularity is what is left of the imbalance).
TAG=${1:-r2m11}
mkdir -p gpurun_out
n=8
for rows in 4 8; do
  o=gpurun_out/bench_n${n}_rows${rows}_${TAG}
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 \
      bench.py --gpus $n --steps 20 --warmup 5 --quick --strip-rows $rows > $o.json 2> $o.err
  echo "== N=$n strip rows $rows rc $?"
  python - <<PY
import json
try:
    d=json.load(open('$o.json'))
    print('ms', round(d['ms_per_step'],3), 'kernel', [round(r['kernel_ms'],2) for r in d['ranks']], 'strips', [b[1] for b in d['strips']])
except Exception as ex: print('no json', ex)
PY
  tail -2 $o.err | cut -c1-200
done
