#!/bin/bash
# N-GPU check: the three exchange modes (peer stores fused into the pack kernel / NCCL all-to-all of runs /
# NCCL all-to-all of tuples) against the oracle, then the bench at N ranks.  usage: gpu_multi.sh N tag [compare]
N=${1:-2}; TAG=${2:-multi$N}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi -L > $OUT/smi.txt
# wide observations (ins_size_threshold > 65535): the int32-pair format of the exchange
WIDE=${WIDE:-1}
run() { timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
PORT=29540
for mode in ${MODES:-peer runs tuples}; do
  for cfg in small_mp_cont small_mp; do
    PORT=$((PORT+1))
    BESST_DIST_EXCHANGE=$mode run $PORT tests/dist_check.py $cfg > $OUT/dist_check_${mode}_$cfg.log 2>&1
    echo "$mode $cfg: $(grep -E 'DIST_CHECK_OK|Error|error|unavailable' $OUT/dist_check_${mode}_$cfg.log | tail -2)"
  done
done
if [ "$WIDE" == "1" ]; then
  run 29559 tests/dist_check.py small_mp 70000 > $OUT/dist_check_wide.log 2>&1
  echo "peer wide: $(grep -E 'DIST_CHECK_OK|Error|error' $OUT/dist_check_wide.log | tail -2)"
fi
show() { python - <<PY
import json
try:
    d=json.loads([l for l in open("$1") if l.startswith("{")][-1])
    print("$2 N=%d ms/step"%d["n_gpus"], d["ms_per_step"], "value %.4g"%d["value"], "e2e", d["e2e"], "phases", d.get("dist_phases_ms"))
except Exception as e:
    print("bench failed", e); print(open("$1".replace(".json",".err")).read()[-3000:])
PY
}
run 29560 bench.py --gpus $N --steps 5 --warmup 3 > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err; show $OUT/bench_n$N.json peer
if [ "$3" == "compare" ]; then
BESST_DIST_EXCHANGE=runs run 29561 bench.py --gpus $N --steps 5 --warmup 3 > $OUT/bench_runs_n$N.json 2> $OUT/bench_runs_n$N.err; show $OUT/bench_runs_n$N.json runs
fi
