#!/bin/bash
# N=8 and N=4 back to back (quick lines: value + per-rank kernel times), then the full N=8 line.
TAG=${1:-r2m3}
mkdir -p gpurun_out
for n in 8 4; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 \
      bench.py --gpus $n --steps 20 --warmup 5 --quick > gpurun_out/quick_n${n}_${TAG}.json 2> gpurun_out/quick_n${n}_${TAG}.err
  echo "== N=$n rc $?"; cut -c1-1500 gpurun_out/quick_n${n}_${TAG}.json
done
bash tools/gpu_r2_multi.sh ${TAG} 8
