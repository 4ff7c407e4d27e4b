#!/bin/bash
# A/B of k_trace tuning knobs on the GPU box: rebuilds spb_wavefront.o per variant, quick bench.
mkdir -p gpurun_out
OUT=gpurun_out/ab_${1:-x}.txt
: > $OUT
shift
for defs in "$@"; do
  echo "== $defs" >> $OUT
  SPB_NVCC_DEFS="$defs" python -c "import __graft_entry__ as e; e.build_library()" >> $OUT 2>&1
  timeout 300 python bench.py --steps 3 --warmup 3 --quick >> $OUT 2>&1
done
cat $OUT
