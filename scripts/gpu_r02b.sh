#!/bin/bash
# round 2, GPU call b: full parity suite (new tests), bench N=1 default + config2 with the PE-level leg
OUT=gpurun_out/r02b
mkdir -p $OUT
timeout 2400 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1
echo "pytest exit $?" >> $OUT/pytest_gpu.log
tail -5 $OUT/pytest_gpu.log
timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.err
python - <<PY
import json
d=json.loads(open("$OUT/bench.json").read().strip().splitlines()[-1])
print("config3", d["ms_per_step"], d["value"], d["e2e"], d["libmetrics"], d["pe_level"], d["parity"]["integers_bit_exact"])
PY
tail -5 $OUT/bench.err
timeout 900 python bench.py --workload config2 > $OUT/bench_config2.json 2> $OUT/bench_config2.err
python - <<PY
import json
d=json.loads(open("$OUT/bench_config2.json").read().strip().splitlines()[-1])
print("config2", d["ms_per_step"], d["value"], d["e2e"], d["libmetrics"], d["pe_level"], d["parity"]["integers_bit_exact"])
PY
tail -5 $OUT/bench_config2.err
