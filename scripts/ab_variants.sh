#!/bin/bash
# bench every besst_b200/variants/libbesst_b200.<name>.so (same ABI, different tuning knobs)
mkdir -p gpurun_out/variants
for so in besst_b200/variants/libbesst_b200.*.so; do
  name=$(basename $so .so); name=${name#libbesst_b200.}
  BESST_B200_LIB=$PWD/$so timeout -s KILL 200 python bench.py --no-cpu --no-e2e --steps 8 --warmup 3 > gpurun_out/variants/$name.json 2> gpurun_out/variants/$name.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/variants/$name.json")); k=d["kernels"]
    print("%-10s step %.4f ms | K1 %.4f group %.4f ks %.4f gather %.4f"%("$name", d["ms_per_step"], k["k_extract_links"]["ms_per_step"], k["k_group_blocks"]["ms_per_step"], k["k_ks_block"]["ms_per_step"], k["k_edge_reduce"]["ms_per_step"]))
except Exception as e:
    print("$name failed", e)
PY
done
