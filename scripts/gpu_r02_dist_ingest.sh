#!/bin/bash
# N GPUs of one box: file -> per-rank device ingest -> distributed graph build vs the oracle, aggregate ingest rate;
# on one GPU first: the part tests through the C ABI.   usage: gpu_r02_dist_ingest.sh N
N=${1:-2}
O=gpurun_out/r02_dist_ingest
mkdir -p $O
timeout -s KILL 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29571 tests/dist_ingest_check.py ${PAIRS:-2400000} > $O/dist_ingest_n$N.log 2>&1; echo "dist rc=$?"
grep -E "DIST_INGEST|DIST_ENTRY|Error|error|assert" $O/dist_ingest_n$N.log | tail -5
cp gpurun_out/dist_ingest_n$N.json $O/ 2>/dev/null
