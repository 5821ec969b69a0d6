#!/bin/bash
# parity tests + the default bench line (with e2e and cpu legs)
TAG=$1; shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout -s KILL 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1
echo "pytest exit $?" >> $OUT/pytest_gpu.log
tail -15 $OUT/pytest_gpu.log
env "$@" timeout -s KILL 600 python bench.py > $OUT/bench.json 2> $OUT/bench.err
python - <<PY
import json
try:
    d=json.load(open("$OUT/bench.json"))
    print("ms/step", d["ms_per_step"], "value %.4g"%d["value"], "clock", d["clocks"])
    print("e2e", d["e2e"]); print("roofline", d["roofline"]); print("cpu", d["cpu_baseline"]); print("parity", d["parity"])
    for k,v in d["kernels"].items(): print("   %-20s %8.4f ms  x%-4g %s"%(k, v["ms_per_step"], v["launches_per_step"], v["GBps"]))
except Exception as e:
    print("bench failed", e); print(open("$OUT/bench.err").read()[-3000:])
PY
