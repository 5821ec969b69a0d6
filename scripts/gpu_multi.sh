#!/bin/bash
# N-GPU check: NCCL path against the oracle, then the bench at N ranks
N=${1:-2}; TAG=${2:-multi$N}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi -L > $OUT/smi.txt
timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/dist_check.py > $OUT/dist_check.log 2>&1
tail -3 $OUT/dist_check.log
timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tests/dist_check.py small_mp > $OUT/dist_check2.log 2>&1
tail -1 $OUT/dist_check2.log
timeout -s KILL 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 5 --warmup 3 > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err
tail -c 2500 $OUT/bench_n$N.json
tail -5 $OUT/bench_n$N.err
