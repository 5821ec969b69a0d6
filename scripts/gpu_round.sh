#!/bin/bash
# One gpurun call: GPU parity tests, the bench (both arms), the ncu launch list of the same
# command and one full capture of the two heaviest kernels.  Outputs under gpurun_out/<tag>/.
# usage: scripts/gpu_round.sh <tag> [skip-tests]
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
if [ "$2" != "skip-tests" ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1
  echo "pytest exit $?" >> $OUT/pytest_gpu.log
  tail -3 $OUT/pytest_gpu.log
fi
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err
timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.err
tail -c 3000 $OUT/bench.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
    --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu > $OUT/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:'k_extract_links|k_group_blocks|k_ks_block|k_edge_gather|k_run_write|k_tile_offsets|k_tile_reduce' -c 8 -f -o $OUT/prof \
    python bench.py --steps 1 --warmup 0 --no-cpu --no-e2e > $OUT/ncu_full.log 2>&1
ls -la $OUT
