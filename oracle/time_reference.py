"""ORACLE / TEST INFRASTRUCTURE -- time the reference's own Python path (its unmodified
bytecode through oracle/ref_harness.py, records pre-decoded so that BAM inflate is excluded on
both sides) on one core of this machine.  Only runs where /root/reference exists.

    python oracle/time_reference.py [synth config] [scale]
"""
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (ROOT, HERE, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import ref_harness  # noqa: E402
from besst_b200 import synth  # noqa: E402


def main():
    config = sys.argv[1] if len(sys.argv) > 1 else "config3"
    scale = float(sys.argv[2]) if len(sys.argv) > 2 else 0.005
    lib = synth.make_config(config, scale=scale)
    batch = lib.to_batch()
    opts = dict(orientation=lib.orientation, mean=lib.mu, stddev=lib.sigma, readlen=100)
    ref_harness.load_reference()
    t0 = time.perf_counter()
    out = ref_harness.run_reference(batch, opts, run_libmetrics=True)
    dt = time.perf_counter() - t0
    print("reference Python path: %s x%g = %d contigs, %d pairs (%d records): %.2f s -> %.3g read-pairs/s on 1 core; "
          "G %d edges, G_prime %d edges" % (config, scale, len(batch.references), lib.n_pairs, len(batch), dt,
                                            lib.n_pairs / dt, len(out["G"]["edges"]), len(out["G_prime"]["edges"])))


if __name__ == "__main__":
    main()
