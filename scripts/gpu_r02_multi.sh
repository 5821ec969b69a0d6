#!/bin/bash
# round 2 multi-GPU call: dist_check (three exchange modes) against the single-pass oracle, then the bench at N ranks with
# its default workload (config3 global x2 at N=2, config4 at N=4, config5 at N=8) incl. the whole-workload parity leg.
# usage: gpu_r02_multi.sh N tag [extra bench args]
N=${1:-2}; TAG=${2:-r02_n$N}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi -L > $OUT/smi.txt; nvidia-smi topo -m >> $OUT/smi.txt 2>&1; nproc >> $OUT/smi.txt; free -g >> $OUT/smi.txt
run() { timeout -s KILL ${TMO:-900} python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
PORT=29540
for mode in ${MODES:-peer runs tuples}; do
  PORT=$((PORT+1))
  BESST_DIST_EXCHANGE=$mode run $PORT tests/dist_check.py small_mp_cont > $OUT/dist_check_${mode}.log 2>&1
  echo "$mode: $(grep -E 'DIST_CHECK_OK|DIST_PE_OK|Error|error|unavailable' $OUT/dist_check_${mode}.log | tail -3)"
done
show() { python - <<PY
import json
try:
    d=json.loads([l for l in open("$1") if l.startswith("{")][-1])
    print("N=%d %s: ms/step"%(d["n_gpus"], d["config"]["workload"][:30]), d["ms_per_step"], "value %.4g"%d["value"], "e2e", d["e2e"], "phases", d.get("dist_phases_ms"))
    print("parity", {k: v for k, v in (d.get("parity") or {}).items() if k != "libraries"}, "cpu", d.get("cpu_baseline"))
except Exception as e:
    print("bench failed", e); print(open("$1".replace(".json",".err")).read()[-3000:])
PY
}
NCCL_DEBUG=INFO run 29560 bench.py --gpus $N --steps 10 --warmup 3 "${@:3}" > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err; show $OUT/bench_n$N.json
grep -E "nranks|NVLS|P2P" $OUT/bench_n$N.err | head -5
