"""GPU: device BAM ingest on a synthetic library written as a BAM file: throughput of besst_bam_ingest (wall, device
phases) next to the host-thread reader, both BGZF writer styles, column equality.  `python scripts/ingest_probe.py [pairs]`"""
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from besst_b200 import bamio, synth   # noqa: E402
from besst_b200.engine import CudaEngine   # noqa: E402


def main():
    pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 2000000
    small = "--small" in sys.argv
    lib = synth.make_library(max(50, pairs // 2000), pairs, "rf", 3000.0, 500.0, 0.0, seed=5)
    batch = lib.to_batch()
    eng = CudaEngine(0)
    out = {}
    d = tempfile.mkdtemp()
    for style in ("htslib", "packed"):
        path = os.path.join(d, style + ".bam")
        t = time.time()
        bamio.write_bam_columns(path, batch, style=style)
        t_write = time.time() - t
        if small:
            os.environ["BESST_BAM_WINDOW"] = str(1 << 20)
        runs = []
        for _ in range(3):
            dev = eng.ingest_bam(path)
            runs.append(dict(dev.stats))
        host = dev.to_host()
        for f in ("tid", "mtid", "pos", "mpos", "tlen", "qlen", "flag", "mapq"):
            assert np.array_equal(getattr(host, f), getattr(batch, f)), f
        t = time.time()
        nat = bamio.read_bam_native(path)
        t_host = time.time() - t
        best = min(runs, key=lambda s: s["seconds_total"])
        out[style] = {"records": len(batch), "write_s": t_write, "device": best, "device_first_call_s": runs[0]["seconds_total"],
                      "device_records_per_s": len(batch) / best["seconds_total"],
                      "inflate_GBps": best["uncompressed_bytes"] / 1e6 / max(best["ms_inflate"], 1e-9),
                      "host_threads": {"seconds_total": t_host, "records_per_s": len(batch) / t_host, **{k: nat.stats[k] for k in ("threads", "seconds_inflate", "seconds_decode")}}}
        print(style, json.dumps(out[style]))
    eng.close()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "ingest_probe.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
