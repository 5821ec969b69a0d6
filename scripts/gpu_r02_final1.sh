#!/bin/bash
# final single-GPU evidence of round 2: smoke, GPU parity suite, bench (both arms)
OUT=gpurun_out/r02_final
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; tail -1 $OUT/smoke.log
timeout 1500 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; tail -2 $OUT/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err
python - <<PY
import json
d=json.loads(open("$OUT/bench.json").read().strip().splitlines()[-1])
print("N=1", d["ms_per_step"], "%.4g"%d["value"], "e2e", d["e2e"]["ms_per_step"], "%.4g"%d["e2e"]["value"], "roofline", d["roofline"]["frac"], d["roofline"]["whole_step"]["frac"], "parity", d["parity"]["integers_bit_exact"], d["parity"]["gap_equal"], d["parity"]["score_max_rel_diff"])
print("ingest", d.get("ingest")); print("pe", d.get("pe_level")); print("libmetrics", d.get("libmetrics")); print("cpu", d["cpu_baseline"]); print(d["clocks"], d["gpu_launches"])
print({k: v["ms_per_step"] for k, v in d["kernels"].items()})
r=json.loads(open("$OUT/bench_reference.json").read().strip().splitlines()[-1]); print("reference arm", "%.4g"%r["value"], r["config"])
PY
grep "\[bench" $OUT/bench.err | tail -4
