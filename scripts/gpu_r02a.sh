#!/bin/bash
# round 2, first GPU call: parity tests (incl. full-size config 3 vs the oracle), bench N=1 (both arms), multi-library workload at N=1
OUT=gpurun_out/r02a
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
nproc > $OUT/nproc.txt; free -g >> $OUT/nproc.txt
timeout 1800 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1
echo "pytest exit $?" >> $OUT/pytest_gpu.log
tail -5 $OUT/pytest_gpu.log
timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.err
tail -c 2500 $OUT/bench.json; tail -5 $OUT/bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err
timeout 900 python bench.py --workload config4 --scale 0.5 --steps 3 > $OUT/bench_config4_half_n1.json 2> $OUT/bench_config4_half_n1.err
tail -c 1500 $OUT/bench_config4_half_n1.json; tail -5 $OUT/bench_config4_half_n1.err
