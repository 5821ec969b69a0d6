#!/bin/bash
# N-GPU check: NCCL paths (run-level and tuple-level exchange) against the oracle, then the bench at N ranks
N=${1:-2}; TAG=${2:-multi$N}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi -L > $OUT/smi.txt
run() { timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
run 29511 tests/dist_check.py > $OUT/dist_check_runs.log 2>&1; grep -E "DIST_CHECK_OK|Error|error" $OUT/dist_check_runs.log | tail -3
run 29512 tests/dist_check.py small_mp > $OUT/dist_check_runs2.log 2>&1; grep -E "DIST_CHECK_OK|Error|error" $OUT/dist_check_runs2.log | tail -3
BESST_DIST_EXCHANGE=tuples run 29513 tests/dist_check.py > $OUT/dist_check_tuples.log 2>&1; grep -E "DIST_CHECK_OK|Error|error" $OUT/dist_check_tuples.log | tail -3
run 29514 bench.py --gpus $N --steps 5 --warmup 3 > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err
python - <<PY
import json
try:
    d=json.load(open("$OUT/bench_n$N.json"))
    print("N=%d ms/step"%d["n_gpus"], d["ms_per_step"], "value %.4g"%d["value"], "wall", d["wall_ms_per_step"], "e2e", d["e2e"])
    for k,v in d["kernels"].items(): print("   %-20s %8.4f ms  x%-4g"%(k, v["ms_per_step"], v["launches_per_step"]))
except Exception as e:
    print("bench failed", e); print(open("$OUT/bench_n$N.err").read()[-3000:])
PY
if [ "$3" == "tuples" ]; then
BESST_DIST_EXCHANGE=tuples run 29515 bench.py --gpus $N --steps 5 --warmup 3 --no-e2e > $OUT/bench_tuples_n$N.json 2> $OUT/bench_tuples_n$N.err
python -c "
import json; d=json.load(open('$OUT/bench_tuples_n$N.json')); print('tuple exchange: ms/step', d['ms_per_step'], 'value %.4g'%d['value'])"
fi
