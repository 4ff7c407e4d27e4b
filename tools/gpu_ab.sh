#!/bin/bash
# A/B of tuning knobs on the GPU box: per variant, rebuild the wavefront unit with extra nvcc
# defines and run a quick bench.  usage: bash tools/gpu_ab.sh <tag> "<defs>|<bench args>" ...
mkdir -p gpurun_out
OUT=gpurun_out/ab_${1:-x}.txt
: > $OUT
shift
for v in "$@"; do
  defs="${v%%|*}"; args=""
  if [[ "$v" == *"|"* ]]; then args="${v#*|}"; fi
  echo "== defs[$defs] args[$args]" >> $OUT
  SPB_NVCC_DEFS="$defs" python -c "import __graft_entry__ as e; e.build_library()" >> $OUT 2>&1
  timeout 300 python bench.py --steps 3 --warmup 3 --quick $args 2>&1 | cut -c1-200 >> $OUT
done
# leave the default build behind
SPB_NVCC_DEFS="" python -c "import __graft_entry__ as e; e.build_library()" >> $OUT 2>&1
cat $OUT
