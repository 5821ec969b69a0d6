#!/bin/bash
# one GPU: ingest with chunk-streamed upload: GPU tests of the ingest, then read-thread / chunk-size variants on one file
O=gpurun_out/r02_ingest_io
mkdir -p $O
timeout 300 python -m pytest tests/test_bamdev.py tests/test_bamdev_parts.py -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -1 $O/pytest.log
F=/tmp/probe.bam
python scripts/ingest_probe.py 2000000 --style=htslib --file=$F --tag=default > $O/default.log 2>&1; tail -1 $O/default.log | cut -c1-150
for t in 4 8 16; do
  BESST_BAM_READ_THREADS=$t python scripts/ingest_probe.py --style=htslib --file=$F --tag=threads$t > $O/threads$t.log 2>&1; tail -1 $O/threads$t.log | cut -c1-150
done
BESST_BAM_CHUNK=$((8<<20)) python scripts/ingest_probe.py --style=htslib --file=$F --tag=chunk8 > $O/chunk8.log 2>&1; tail -1 $O/chunk8.log | cut -c1-150
BESST_BAM_CHUNK=$((1<<30)) python scripts/ingest_probe.py --style=htslib --file=$F --tag=chunk1g > $O/chunk1g.log 2>&1; tail -1 $O/chunk1g.log | cut -c1-150
cp gpurun_out/ingest_*.json $O/ 2>/dev/null
