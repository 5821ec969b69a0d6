"""Run under torchrun on N GPUs of one box: the NCCL path of besst_b200/dist.py against the
single-pass C oracle (rank 0 checks; every rank must finish).

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
      --master-port 29511 tests/dist_check.py [config]
"""
import os
import sys

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
from besst_b200 import abi, synth  # noqa: E402
from besst_b200.dist import CudaBackend, DistributedGraphBuild  # noqa: E402
from besst_b200.engine import CudaEngine  # noqa: E402


def main():
    config = sys.argv[1] if len(sys.argv) > 1 else "small_mp_cont"
    threshold = float(sys.argv[2]) if len(sys.argv) > 2 else None   # e.g. 70000: observations travel as int32 pairs, not 2 x u16
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    if rank == 0:
        oracle_lib.build()
    dist.barrier()
    lib = synth.make_config(config)
    batch = lib.to_batch()
    params = abi.make_params(lib.orientation, 11, 100.0, lib.mu, lib.sigma, threshold if threshold else lib.mu + 6 * lib.sigma)
    objs = helpers.later_library_objects(batch.references, batch.lengths, lib.mu + 4 * lib.sigma, seed=3)
    table = helpers.table_for(batch, objs)
    n = len(batch)
    bounds = [(n * r // world) - ((n * r // world) % 16) for r in range(world)] + [n]
    sl = batch.slice(bounds[rank], bounds[rank + 1])
    cols = {k: torch.from_numpy(np.ascontiguousarray(v).view(np.int16) if k == "flag" else np.ascontiguousarray(v)).to(dev)
            for k, v in sl.device_arrays().items()}
    ptrs = {k: v.data_ptr() for k, v in cols.items()}
    ptrs["n"] = len(sl)
    rec = abi.make_records(ptrs, on_device=True)
    eng = CudaEngine(local)
    eng.set_table(table)
    torch.cuda.set_stream(torch.cuda.Stream(device=dev))
    eng.set_stream(torch.cuda.current_stream().cuda_stream)
    runner = DistributedGraphBuild(CudaBackend(eng, dev), rank, world)
    for _ in range(2):   # twice: buffers are reused
        runner.step(params, rec)
    merged = runner.fetch_global()
    if rank == 0:
        want, _, _, consistent = oracle_lib.graph_build(table.rows, table.n_scaffolds, params, batch)
        assert consistent
        helpers.assert_graph_equal(merged, want, label="dist world=%d %s" % (world, config))
        print("DIST_CHECK_OK world=%d %s: %d edges, %d links" % (world, config, merged.n_edges, merged.n_links))
    dist.barrier()
    # the reference-facing entry point itself under the process group: besst_b200.CreateGraph.PE takes the multi-GPU build
    # (every rank its BAM-order slice, merged CSR on every rank) and must leave the reference's graphs on EVERY rank
    import test_golden_reference as tg
    for case in ("small_mp_given", "small_pe_later", "testset1_travis"):
        tg.check_case(case, eng)
    dist.barrier()
    if rank == 0:
        print("DIST_PE_OK world=%d: CreateGraph.PE == reference goldens on every rank" % world)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
