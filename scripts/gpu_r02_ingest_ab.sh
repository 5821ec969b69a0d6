#!/bin/bash
# one GPU: A/B of the inflate kernel's group width / table size variants on one 4 M-record BAM file
O=gpurun_out/r02_ingest_ab
mkdir -p $O
F=/tmp/probe.bam
python scripts/ingest_probe.py 2000000 --style=htslib --file=$F --tag=default > $O/default.log 2>&1; tail -1 $O/default.log
for so in besst_b200/variants/libbesst_b200.*.so; do
  name=$(basename $so .so); name=${name#libbesst_b200.}
  BESST_B200_LIB=$PWD/$so timeout -s KILL 120 python scripts/ingest_probe.py --style=htslib --file=$F --tag=$name > $O/$name.log 2>&1
  tail -1 $O/$name.log | cut -c1-160
done
cp gpurun_out/ingest_*.json $O/
