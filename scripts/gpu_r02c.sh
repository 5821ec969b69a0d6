#!/bin/bash
# round 2, GPU call c: parity suite, bench N=1 packed (default) and plain records, config2 with the PE-level leg
OUT=gpurun_out/r02c
mkdir -p $OUT
timeout 2400 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1
echo "pytest exit $?" >> $OUT/pytest_gpu.log
tail -5 $OUT/pytest_gpu.log
show() { python - <<PY
import json
d=json.loads(open("$1").read().strip().splitlines()[-1])
print("$2", d["ms_per_step"], "%.4g"%d["value"], "e2e", d["e2e"], "roofline", d["roofline"]["frac"], d["roofline"]["whole_step"]["frac"], "pe", d.get("pe_level"), "parity", d["parity"]["integers_bit_exact"] if d.get("parity") else None)
print({k: v["ms_per_step"] for k, v in d["kernels"].items()})
PY
}
timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.err; show $OUT/bench.json packed; tail -3 $OUT/bench.err
timeout 900 python bench.py --records plain --no-cpu > $OUT/bench_plain.json 2> $OUT/bench_plain.err; show $OUT/bench_plain.json plain; tail -3 $OUT/bench_plain.err
timeout 900 python bench.py --workload config2 > $OUT/bench_config2.json 2> $OUT/bench_config2.err; show $OUT/bench_config2.json config2; tail -3 $OUT/bench_config2.err
