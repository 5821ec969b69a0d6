"""Run under torchrun on N GPUs of one box: BAM FILE -> every rank inflates and decodes its part on its own GPU
(besst_b200.dist.ingest_bam_distributed) -> DistributedGraphBuild.step on the device-resident parts -> merged CSR, against
the single-pass C oracle on the whole library (rank 0 checks; every rank must finish).  Then the aggregate ingest rate on a
larger file (the max over ranks of the wall clock of the call, barrier on both sides).

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 \
      tests/dist_ingest_check.py [pairs of the throughput file]
"""
import json
import os
import sys
import tempfile
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (ROOT, os.path.join(ROOT, "oracle"), HERE):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import helpers  # noqa: E402
import oracle_lib  # noqa: E402
from besst_b200 import abi, bamio, synth  # noqa: E402
from besst_b200.dist import CudaBackend, DistributedGraphBuild, ingest_bam_distributed  # noqa: E402
from besst_b200.engine import CudaEngine  # noqa: E402


def main():
    pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 2400000
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    host_group = dist.new_group(backend="gloo")   # the (first, landing) rows are Python objects
    d = tempfile.gettempdir()
    small, big = os.path.join(d, "besst_dist_small.bam"), os.path.join(d, "besst_dist_big.bam")
    lib = synth.make_config("small_mp_cont")
    batch = lib.to_batch()
    if rank == 0:
        oracle_lib.build()
        bamio.write_bam_columns(small, batch, style="packed")   # records straddle BGZF blocks and part boundaries
        # a PE library whose first half holds the 1e6 insert-size samples of libmetrics (reached after ~2.3 M records)
        bamio.write_bam_columns(big, synth.make_library(max(50, pairs // 10000), pairs, "fr", 550.0, 50.0, 0.0, seed=17).to_batch())
    dist.barrier()
    eng = CudaEngine(local)
    torch.cuda.set_stream(torch.cuda.Stream(device=dev))
    eng.set_stream(torch.cuda.current_stream().cuda_stream)

    # ---- file -> parts in HBM -> distributed graph build == the oracle on the whole library --------------------------------
    part, info = ingest_bam_distributed(eng, small, rank, world, group=host_group)
    assert sum(info["counts"]) == len(batch), (info, len(batch))
    host = part.to_host()
    lo = info["record_base"]
    for f in ("tid", "mtid", "pos", "mpos", "tlen", "qlen", "flag", "mapq"):
        assert np.array_equal(getattr(host, f), getattr(batch, f)[lo:lo + len(part)]), f
    params = abi.make_params(lib.orientation, 11, 100.0, lib.mu, lib.sigma, lib.mu + 6 * lib.sigma)
    objs = helpers.later_library_objects(batch.references, batch.lengths, lib.mu + 4 * lib.sigma, seed=3)
    table = helpers.table_for(batch, objs)
    eng.set_table(table)
    runner = DistributedGraphBuild(CudaBackend(eng, dev), rank, world)
    runner.step(params, part.abi_records)
    merged = runner.fetch_global()
    if rank == 0:
        want, _, _, consistent = oracle_lib.graph_build(table.rows, table.n_scaffolds, params, batch)
        assert consistent
        helpers.assert_graph_equal(merged, want, label="file -> %d GPUs -> graph" % world)
        print("DIST_INGEST_OK world=%d: %d records in parts %s (repeats %d) -> %d edges, %d links == oracle" % (
            world, len(batch), info["counts"], info["repeats"], merged.n_edges, merged.n_links))
    dist.barrier()

    # ---- the Python entry points on a PATH under the process group: every rank ingests its part on its GPU, rank 0's
    #      library metrics are broadcast (or the whole file is read when the sampled prefix is not inside part 0), PE builds
    #      from the parts -- same graphs as with the host reader (every rank decodes the whole file and slices it) ----------
    from besst_b200 import records
    whole_file_reads = []
    real_reader = bamio.read_bam_native
    bamio.read_bam_native = lambda *a, **k: (whole_file_reads.append(1), real_reader(*a, **k))[1]
    for label, path, orient in (("small", small, lib.orientation), ("big", big, "fr")):
        hdr = eng.ingest_bam(path, part=(0, max(world, 64)))   # header only (a sliver of the file)
        lengths = dict(zip(hdr.references, hdr.lengths))
        sigs, reads = {}, {}
        for mode in ("host", "device"):
            os.environ["BESST_B200_INGEST"] = mode
            records._open_cache.clear()
            del whole_file_reads[:]
            out = helpers.run_dropin(None, dict(orientation=orient, mean=None, stddev=None, readlen=None), eng,
                                     fasta_lengths=lengths, bam_path=path)
            sigs[mode] = {k: out[k] for k in ("G", "G_prime", "param", "objects")}
            reads[mode] = len(whole_file_reads)
            dist.barrier()
        os.environ.pop("BESST_B200_INGEST", None)
        assert sigs["host"] == sigs["device"], "entry points on %s: device parts != host reader" % label
        assert len(sigs["device"]["G_prime"]["edges"]) > 0
        # small: the sampled prefix is the whole file -> every rank reads it once for the metrics; big (world 2): it lies
        # inside part 0 -> no rank reads the whole file
        if rank == 0:
            print("DIST_ENTRY_OK world=%d %s: get_metrics + PE on a path, ingest in parts on the devices == host reader "
                  "(%d G_prime edges; whole-file reads in device mode: %d)" % (world, label, len(sigs["device"]["G_prime"]["edges"]), reads["device"]))
        dist.barrier()
    bamio.read_bam_native = real_reader

    # ---- aggregate ingest rate -----------------------------------------------------------------------------------------------
    ingest_bam_distributed(eng, big, rank, world, group=host_group)   # warm-up: page cache, buffers
    best = None
    for _ in range(3):
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        part, info = ingest_bam_distributed(eng, big, rank, world, group=host_group)
        torch.cuda.synchronize()
        t = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if best is None or float(t.item()) < best[0]:
            best = (float(t.item()), dict(part.stats), info)
    if rank == 0:
        n = sum(best[2]["counts"])
        line = {"world": world, "records": n, "wall_ms_max_over_ranks": round(1e3 * best[0], 2), "records_per_s": n / best[0],
                "rank0": {k: best[1][k] for k in ("compressed_bytes", "uncompressed_bytes", "blocks", "windows", "ms_inflate", "seconds_read")},
                "counts": best[2]["counts"], "repeats": best[2]["repeats"]}
        print("DIST_INGEST_RATE " + json.dumps(line))
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        json.dump(line, open(os.path.join(ROOT, "gpurun_out", "dist_ingest_n%d.json" % world), "w"), indent=1)
        os.remove(small)
        os.remove(big)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
