#!/bin/bash
# N=8 bench only, tight limits (a stuck collective must not burn the GPU budget)
OUT=gpurun_out/r02_n8c
mkdir -p $OUT
timeout -s KILL 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29571 \
    bench.py --gpus 8 --steps 10 --warmup 3 --watchdog 170 > $OUT/bench_n8.json 2> $OUT/bench_n8.err
echo "exit $?"
grep "\[bench" $OUT/bench_n8.err | tail -12
python - <<PY
import json
try:
    d=json.loads([l for l in open("$OUT/bench_n8.json") if l.startswith("{")][-1])
    print("N=%d ms/step"%d["n_gpus"], d["ms_per_step"], "value %.4g"%d["value"], "e2e", d["e2e"], "phases", d.get("dist_phases_ms"))
    print("parity", {k: v for k, v in (d.get("parity") or {}).items() if k != "libraries"})
except Exception as e:
    print("bench failed", e)
PY
grep -v "NCCL INFO" $OUT/bench_n8.err | grep -E "Traceback|Error|rror:" | head -5
