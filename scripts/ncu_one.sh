#!/bin/bash
# full ncu capture of the kernels matching a regex during one build of the bench workload
# usage: scripts/ncu_one.sh <tag> <regex> <count> [ENV=..]
TAG=$1; RX=$2; CNT=$3; shift 3
OUT=gpurun_out/$TAG
mkdir -p $OUT
env "$@" timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:"$RX" -c $CNT -f -o $OUT/prof \
    python bench.py --steps 1 --warmup 0 --no-cpu --no-e2e > $OUT/ncu.log 2>&1
tail -3 $OUT/ncu.log | cut -c1-300
ls -la $OUT
