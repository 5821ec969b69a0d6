#!/bin/bash
N=${1:-2}; OUT=gpurun_out/phases$N; mkdir -p $OUT
BESST_DIST_TIMING=1 timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus $N --steps 5 --warmup 3 --no-e2e > $OUT/bench.json 2> $OUT/bench.err
python -c "
import json
d=json.loads([l for l in open('$OUT/bench.json') if l.startswith('{')][-1]); print('ms/step (with phase syncs)', d['ms_per_step']); print(d['dist_phases_ms'])"
