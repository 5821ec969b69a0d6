#!/bin/bash
# one GPU: the device BAM ingest -- parity tests, compute-sanitizer memcheck on a small multi-window ingest, throughput probe
mkdir -p gpurun_out/r02_ingest
O=gpurun_out/r02_ingest
timeout 400 python -m pytest tests/test_bamdev.py -m gpu -x -q -s > $O/pytest_bamdev.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_bamdev.log
tail -5 $O/pytest_bamdev.log
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/ingest_probe.py 30000 --small > $O/memcheck.log 2>&1; echo "memcheck rc=$?" | tee -a $O/memcheck.log
tail -4 $O/memcheck.log
timeout 300 python scripts/ingest_probe.py 2000000 > $O/probe.log 2>&1; echo "probe rc=$?" | tee -a $O/probe.log
tail -3 $O/probe.log
cp gpurun_out/ingest_probe.json $O/ 2>/dev/null
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/smi.txt
