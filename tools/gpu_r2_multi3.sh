#!/bin/bash
TAG=${1:-r2m4}
mkdir -p gpurun_out
for w in 5 12; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 \
      bench.py --gpus 8 --steps 20 --warmup $w --quick > gpurun_out/quick_n8_w${w}_${TAG}.json 2> gpurun_out/quick_n8_w${w}_${TAG}.err
  echo "== warmup $w rc $?"; python - <<PY
import json
d=json.load(open('gpurun_out/quick_n8_w${w}_${TAG}.json'))
print(d['ms_per_step'], [round(r['kernel_ms'],2) for r in d['ranks']])
for h in d['rebalance_history']: print(h)
PY
done
