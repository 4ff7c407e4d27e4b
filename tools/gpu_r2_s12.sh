#!/bin/bash
# Round 2, GPU session 12 (one GPU): the whole GPU suite with sp_b200_RenderRowsBegin / End in the library, then a
# quick C3 line (device-resident) to see that the one-call form did not slow down.
TAG=${1:-r2s12}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_${TAG}.log 2>&1
tail -15 gpurun_out/pytest_gpu_${TAG}.log
timeout 200 python bench.py --steps 6 --warmup 3 --quick 2>&1 | cut -c1-400 > gpurun_out/quick_c3_${TAG}.json
cat gpurun_out/quick_c3_${TAG}.json
