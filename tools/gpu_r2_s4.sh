#!/bin/bash
# Round 2, GPU session 4: the rewritten bench line on one GPU (parity block, C5 secondary, roofline vs L2).
TAG=${1:-r2s4}
mkdir -p gpurun_out
timeout 600 python bench.py --steps 5 --warmup 3 --cpu-seconds 4 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
echo "bench exit $?" >> gpurun_out/bench_${TAG}.err
tail -5 gpurun_out/bench_${TAG}.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_${TAG}.json'))
for k in ('value','ms_per_step','value_traced_only','e2e','parity','clocks'):
    print(k, json.dumps(d.get(k))[:900])
print('roofline', {k:v for k,v in d['roofline'].items() if k in ('bound','achieved','peak','frac','traffic','traversal_grays_per_s','frac_of_l2_plateau')})
print('c5', json.dumps(d.get('secondary'))[:1500])
PY
