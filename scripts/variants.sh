#!/bin/bash
# Build A/B variants of libbesst_b200.so with -D tuning knobs (run HERE, no GPU needed):
#   scripts/variants.sh name1 "-DBESST_K1T_BATCH=16" name2 "-DBESST_KB_MIN_CTAS=3" ...
# then on the GPU box: BESST_B200_LIB=besst_b200/variants/libbesst_b200.<name>.so python bench.py --no-cpu --no-e2e
set -e
cd "$(dirname "$0")/.."
mkdir -p besst_b200/variants
while [ $# -ge 2 ]; do
  name=$1; defs=$2; shift 2
  objs=""
  for f in besst_api besst_links besst_sort besst_edges besst_metrics besst_bamdev; do
    extra=""; [ $f == besst_edges ] && extra="-fmad=false"
    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC $extra $defs -c besst_b200/csrc/$f.cu -o /tmp/var_${name}_$f.o &
    objs="$objs /tmp/var_${name}_$f.o"
  done
  wait
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o besst_b200/variants/libbesst_b200.$name.so $objs
  echo built besst_b200/variants/libbesst_b200.$name.so "($defs)"
done
