#!/bin/bash
# ncu evidence for round 2 (one GPU): launch list of the bench command + full-set capture of the heavy kernels with source
OUT=gpurun_out/r02_ncu
mkdir -p $OUT
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-libmetrics > $OUT/ncu_launches.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on \
    -k regex:'k_extract_links|k_group_blocks|k_ks_block|k_edge_gather|k_tile_offsets|k_tile_reduce' -c 6 -f -o $OUT/prof \
    python bench.py --steps 1 --warmup 0 --no-cpu --no-e2e --no-libmetrics > $OUT/ncu_full.log 2>&1
ls -la $OUT
