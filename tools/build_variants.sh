#!/bin/bash
# Builds A/B variants of libspb200.so HERE (no GPU minutes spent compiling on the box): each variant
# recompiles the wavefront unit with extra nvcc defines and is kept as variants/<name>.so; the default
# build is restored at the end.  usage: bash tools/build_variants.sh name1="<defs>" name2="<defs>" ...
# Run one with:  SPB_B200_LIB=variants/<name>.so python bench.py --quick ...
set -e
mkdir -p variants
for v in "$@"; do
  name="${v%%=*}"; defs="${v#*=}"
  echo "== $name: $defs"
  SPB_NVCC_DEFS="$defs" python -c "import __graft_entry__ as e; e.build_library()"
  cp vk_cinematic_b200/libspb200.so variants/${name}.so
done
SPB_NVCC_DEFS="" python -c "import __graft_entry__ as e; e.build_library()"
ls -la variants
