"""world_size 2 and 3 on CPU (gloo): the host logic of the multi-GPU decomposition
(besst_b200/dist.py) -- halo from preceding ranks, stable bucket exchange, global
first-appearance ordinals, all-reduced coverage/counters, merge -- must reproduce the
single-pass oracle bit for bit (integers)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, config, cuts, out_dir, exchange):
    os.environ["BESST_DIST_EXCHANGE"] = exchange
    expect_redo = cuts == "split_duplicate"
    for p in (ROOT, os.path.join(ROOT, "oracle"), HERE):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch.distributed as dist
    import helpers
    import oracle_lib
    from besst_b200 import abi, synth
    from besst_b200.dist import DistributedGraphBuild
    from dist_backend_numpy import NumpyBackend
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    try:
        lib = synth.make_config(config)
        batch = lib.to_batch()
        params = abi.make_params(lib.orientation, 11, 100.0, lib.mu, lib.sigma, lib.mu + 6 * lib.sigma)
        objs = helpers.later_library_objects(batch.references, batch.lengths, lib.mu + 4 * lib.sigma, seed=3)
        table = helpers.table_for(batch, objs)
        n = len(batch)
        if cuts == "split_duplicate":   # cut between two adjacent identical read2 records: the second slice starts with a duplicate
            same = ((batch.tid[1:] == batch.tid[:-1]) & (batch.pos[1:] == batch.pos[:-1]) & (batch.mtid[1:] == batch.mtid[:-1]) &
                    (batch.mpos[1:] == batch.mpos[:-1]) & (batch.flag[1:] == batch.flag[:-1]) & (batch.tid[1:] != batch.mtid[1:]) &
                    ((batch.flag[1:] & 0x80) != 0) & ((batch.flag[1:] & 0x4) == 0) & (batch.mapq[1:] >= 11) & (batch.mapq[:-1] >= 11))
            cand = np.nonzero(same)[0] + 1
            assert len(cand) > 0
            cuts = [cand[len(cand) // 2] / float(n) + 1e-12]
        bounds = [0] + [int(n * c) for c in cuts] + [n]
        runner = DistributedGraphBuild(NumpyBackend(table), rank, world)
        runner.step(params, batch.slice(bounds[rank], bounds[rank + 1]))
        if expect_redo and exchange == "runs":
            assert runner.redo_count > 0, "the cut was meant to split a duplicate pair"
        merged = runner.fetch_global()
        if rank == 0:
            want, _, _, _ = oracle_lib.graph_build(table.rows, table.n_scaffolds, params, batch)
            for f in ("edge_u", "edge_v", "nr_links", "obs_sum", "obs_sq", "row_ptr", "fishy", "obs_u", "obs_v", "aligned_len"):
                assert np.array_equal(getattr(merged, f), getattr(want, f)), f
            assert np.array_equal(merged.flags & abi.EDGE_LL, want.flags & abi.EDGE_LL)
            assert np.array_equal(merged.counters[:12], want.counters[:12]), (merged.counters[:12], want.counters[:12])
            # first-appearance order of the edges (networkx insertion order) is the global one
            assert np.array_equal(np.argsort(merged.first_idx, kind="stable"), np.argsort(want.first_idx, kind="stable"))
            assert np.array_equal(merged.first_idx, want.first_idx)
            open(os.path.join(out_dir, "ok"), "w").write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,config,cuts", [
    (2, "small_mp", [0.5]),
    (3, "small_pe", [0.2, 0.21]),       # a tiny middle slice
    (3, "tiny", [0.0, 0.6]),            # an empty first slice: the halo must pass through
    (2, "small_mp", "split_duplicate"), # the second slice starts with a duplicate of the first slice's last call
])
@pytest.mark.parametrize("exchange", ["runs", "tuples"])
def test_distributed_build_equals_single_pass(tmp_path, world, config, cuts, exchange):
    import oracle_lib
    oracle_lib.build()   # before the workers race to do it
    port = _free_port()
    mp.spawn(_worker, args=(world, port, config, cuts, str(tmp_path), exchange), nprocs=world, join=True)
    assert os.path.exists(os.path.join(str(tmp_path), "ok"))


# ---- CreateGraph.PE under a torch.distributed job: the entry point itself runs the multi-rank build -------
class _DistEngine(object):
    """OracleEngine for the single-process calls (libmetrics) + the numpy dist backend for the build."""

    def __init__(self):
        from oracle_engine import OracleEngine
        self._o = OracleEngine()
        self.libmetrics = self._o.libmetrics
        self.gapest_batch = self._o.gapest_batch

    def graph_build(self, *a, **k):
        raise AssertionError("PE must take the distributed build when the process group has more than one rank")

    def make_dist_backend(self, table):
        from dist_backend_numpy import NumpyBackend
        return NumpyBackend(table)


def _pe_worker(rank, world, port, case, out_dir):
    for p in (ROOT, os.path.join(ROOT, "oracle"), HERE):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch.distributed as dist
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    try:
        import test_golden_reference as tg
        tg.check_case(case, _DistEngine())   # every rank ends with the reference's graphs, params, objects, counter lines
        open(os.path.join(out_dir, "ok%d" % rank), "w").write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("case", ["small_pe_no_score_later", "testset1_travis_no_score"])
def test_entry_point_PE_runs_the_distributed_build_under_a_process_group(tmp_path, case):
    """north-star: 'called from the existing Python entry points' -- besst_b200.CreateGraph.PE in a world-2 job
    against the goldens minted from the reference's bytecode (no_score cases: the numpy backend has no scores)."""
    import oracle_lib
    oracle_lib.build()
    mp.spawn(_pe_worker, args=(2, _free_port(), case, str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(os.path.join(str(tmp_path), "ok0")) and os.path.exists(os.path.join(str(tmp_path), "ok1"))
