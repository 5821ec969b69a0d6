#!/bin/bash
# one GPU: ncu --set full of k_bgzf_inflate (one launch), launch list of one ingest, bench config2 with the ingest leg
O=gpurun_out/r02_ingest
mkdir -p $O
timeout 300 ncu --set full --import-source on --clock-control none -k regex:k_bgzf_inflate -c 1 -f -o $O/inflate python scripts/ingest_probe.py 2000000 --style=htslib > $O/ncu_run.log 2>&1; echo "ncu rc=$?"
ncu -i $O/inflate.ncu-rep --page raw --csv > $O/inflate_raw.csv 2>/dev/null
ncu -i $O/inflate.ncu-rep --page source --csv > $O/inflate_source.csv 2>/dev/null
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_b -c 40 --csv --log-file $O/ingest_launches.csv python scripts/ingest_probe.py 2000000 --style=htslib > /dev/null 2>&1; echo "launch list rc=$?"
timeout 400 python bench.py --workload config2 --steps 5 --warmup 3 > $O/bench_config2.json 2> $O/bench_config2.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('$O/bench_config2.json')); print(json.dumps(d['ingest'])); print(d['ms_per_step'], d['value'], d['parity']['integers_bit_exact'])"
