#!/bin/bash
# quick A/B on one GPU: parity tests, then the bench (no CPU leg) for each environment setting given
# usage: scripts/ab.sh <tag> "<ENV=.. ENV2=..>" ["<ENV=..>" ...]
TAG=$1; shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout -s KILL 600 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1
echo "pytest exit $?" >> $OUT/pytest_gpu.log
tail -4 $OUT/pytest_gpu.log
i=0
for envs in "$@"; do
  i=$((i+1))
  echo "== variant $i: $envs"
  env $envs timeout -s KILL 300 python bench.py --no-cpu --no-e2e --steps 5 --warmup 3 > $OUT/bench_$i.json 2> $OUT/bench_$i.err
  python - <<PY
import json
try:
    d=json.load(open("$OUT/bench_$i.json"))
    print("ms/step", d["ms_per_step"], "value %.3g"%d["value"], "clock", d["clocks"]["sm_mhz"])
    for k,v in d["kernels"].items(): print("   %-20s %8.4f ms  x%-4g %s"%(k, v["ms_per_step"], v["launches_per_step"], v["GBps"]))
except Exception as e:
    print("bench failed", e); print(open("$OUT/bench_$i.err").read()[-2000:])
PY
done
