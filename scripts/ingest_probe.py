"""GPU: device BAM ingest on a synthetic library written as a BAM file: throughput of besst_bam_ingest (wall, device
phases) next to the host-thread reader, both BGZF writer styles, column equality against the host-thread reader.
`python scripts/ingest_probe.py [pairs] [--style=htslib|packed] [--small] [--file=PATH (reused if present)] [--tag=NAME]`"""
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

FIELDS = ("tid", "mtid", "pos", "mpos", "tlen", "qlen", "flag", "mapq")


def opt(name, default=None):
    for a in sys.argv[1:]:
        if a.startswith("--%s=" % name):
            return a.split("=", 1)[1]
    return default


def main():
    nums = [a for a in sys.argv[1:] if a.isdigit()]
    pairs = int(nums[0]) if nums else 2000000
    styles = [opt("style")] if opt("style") else ["htslib", "packed"]
    tag = opt("tag", "probe")
    eng = CudaEngine(0)
    out = {}
    d = tempfile.mkdtemp()
    for style in styles:
        path = opt("file") or os.path.join(d, style + ".bam")
        t_write = 0.0
        if not os.path.exists(path):
            batch = synth.make_library(max(50, pairs // 2000), pairs, "rf", 3000.0, 500.0, 0.0, seed=5).to_batch()
            t = time.time()
            bamio.write_bam_columns(path, batch, style=style)
            t_write = time.time() - t
        if "--small" in sys.argv:
            os.environ["BESST_BAM_WINDOW"] = str(1 << 20)
        t = time.time()
        nat = bamio.read_bam_native(path)
        t_host = time.time() - t
        runs = []
        for _ in range(4):
            dev = eng.ingest_bam(path)
            runs.append(dict(dev.stats))
        host = dev.to_host()
        for f in FIELDS:
            assert np.array_equal(getattr(host, f), getattr(nat, f)), f
        assert np.array_equal(host.packed, nat.packed)
        best = min(runs[1:], key=lambda s: s["seconds_total"])
        n = len(nat)
        out[style] = {"records": n, "write_s": t_write, "device": best, "device_first_call_s": runs[0]["seconds_total"],
                      "device_records_per_s": n / best["seconds_total"],
                      "inflate_GBps": best["uncompressed_bytes"] / 1e6 / max(best["ms_inflate"], 1e-9),
                      "host_threads": {"seconds_total": t_host, "records_per_s": n / t_host, **{k: nat.stats[k] for k in ("threads", "seconds_inflate", "seconds_decode")}}}
        print("%s %s: inflate %.2f ms = %.1f GB/s out, wall %.1f ms = %.1f M records/s (host threads %.1f M records/s) | %s" % (
            tag, style, best["ms_inflate"], out[style]["inflate_GBps"], 1e3 * best["seconds_total"], n / best["seconds_total"] / 1e6,
            n / t_host / 1e6, json.dumps(best)))
    eng.close()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "ingest_%s.json" % tag), "w"), indent=1)


if __name__ == "__main__":
    main()
