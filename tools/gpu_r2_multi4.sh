#!/bin/bash
# N=8 quick line with the rebalance history, then the full N=8 and N=4, N=2 lines (driver's arguments: --steps 20 --warmup 5)
TAG=${1:-r2m5}
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus 8 --steps 20 --warmup 5 --quick > gpurun_out/quick_n8_${TAG}.json 2> gpurun_out/quick_n8_${TAG}.err
python - <<PY
import json
d=json.load(open('gpurun_out/quick_n8_${TAG}.json'))
print('N=8 quick', d['ms_per_step'], [round(r['kernel_ms'],2) for r in d['ranks']])
for h in d['rebalance_history']: print(h)
PY
for n in 8 4 2; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 \
      bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/bench_n${n}_${TAG}.json 2> gpurun_out/bench_n${n}_${TAG}.err
  echo "== N=$n rc $?"
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_n${n}_${TAG}.json'))
    print('value', d['value'], 'ms', d['ms_per_step'], 'e2e ms', d['e2e']['ms_per_step'], 'kernel', [round(r['kernel_ms'],2) for r in d['ranks']])
    print('parity', json.dumps(d.get('parity'))[:300]); print('c5', (d.get('secondary') or {}).get('c5',{}).get('ms_per_step'), (d.get('secondary') or {}).get('c5',{}).get('rank_kernel_ms'))
except Exception as e: print('no json', e)
PY
done
