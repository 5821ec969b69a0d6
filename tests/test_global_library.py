"""ONE global BAM-ordered library generated slice by slice (besst_b200/synth.make_library_slice), the
later-library contig table at scale (synth.later_library_rows), the slice-wise oracle
(oracle/slice_oracle.py) and the mergeable per-edge digests (besst_b200/digest.py) -- the pieces
bench.py's config 4 / config 5 workloads and their full-size parity leg are made of.  CPU, gloo."""
import os
import socket
import sys
from types import SimpleNamespace

import numpy as np
import pytest
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (ROOT, os.path.join(ROOT, "oracle"), HERE):
    if p not in sys.path:
        sys.path.insert(0, p)

from besst_b200 import abi, digest, synth  # noqa: E402

N_CONTIGS, N_PAIRS, SEED = 900, 260000, 77


def _slices(world, orientation="rf", mu=3000.0, sigma=500.0, cont=0.25):
    return [synth.make_library_slice(N_CONTIGS, N_PAIRS, orientation, mu, sigma, cont, SEED, r, world) for r in range(world)]


def _concat(slices):
    from besst_b200.records import RecordBatch
    parts = [s.to_batch() for s in slices]
    cols = {f: np.concatenate([getattr(p, f) for p in parts]) for f in ("tid", "mtid", "pos", "mpos", "tlen", "qlen", "flag", "mapq")}
    return RecordBatch(references=parts[0].references, lengths=parts[0].lengths, **cols)


@pytest.mark.parametrize("world", [1, 2, 5])
def test_slices_concatenate_to_one_sorted_bam_with_pairs_across_the_cuts(world):
    slices = _slices(world)
    bounds = synth.slice_bounds(N_CONTIGS, world)
    for r, s in enumerate(slices):
        tid = s.cols["tid"].numpy()
        assert tid.size and tid.min() >= bounds[r] and tid.max() < bounds[r + 1]
    b = _concat(slices)
    key = b.tid.astype(np.int64) * (1 << 32) + b.pos
    assert np.all(np.diff(key) >= 0)
    total = sum(s.n_records for s in slices)
    assert abs(total / 2.0 - N_PAIRS * 1.01) < 0.03 * N_PAIRS
    # every record whose mate sits on another contig has that mate record, whichever slice it is in
    inter = (b.tid != b.mtid) & ((b.flag & 0xC) == 0)
    me = np.stack([b.tid[inter], b.pos[inter], b.mtid[inter], b.mpos[inter]], axis=1).astype(np.int64)
    mate = me[:, [2, 3, 0, 1]]
    as_set = set(map(tuple, me.tolist()))
    assert all(tuple(m) in as_set for m in mate.tolist())
    if world > 1:
        owner = np.searchsorted(np.asarray(bounds[1:]), b.tid, side="right")
        mate_owner = np.searchsorted(np.asarray(bounds[1:]), b.mtid, side="right")
        crossing = inter & (owner != mate_owner)
        assert crossing.sum() > 20   # pairs really span the cuts


def test_later_library_rows_describe_consistent_multi_contig_scaffolds():
    lengths = synth.make_contigs(5000, 5)[0].numpy()
    rows, n_scaf, n_large = synth.later_library_rows(lengths, 5000.0, seed=9)
    present = rows["state"] != 0
    assert (~present).sum() == 5000 // 17
    sc = rows["scaffold"][present]
    assert sc.min() == 0 and sc.max() == n_scaf - 1 and np.unique(sc).size == n_scaf
    large = rows["state"][present] == abi.CTG_LARGE
    assert np.all(sc[large] < n_large) and np.all(sc[~large] >= n_large)
    members = np.bincount(sc)
    assert members.max() == 3 and (members > 1).sum() > n_scaf // 4
    # members tile their scaffold left to right; the last one ends at scaf_length
    end = rows["position"][present] + rows["length"][present]
    slen = rows["scaf_length"][present]
    assert np.all(end <= slen)
    last_end = np.zeros(n_scaf, np.int64)
    np.maximum.at(last_end, sc, end)
    assert np.array_equal(last_end[sc], slen)
    assert np.all(slen[large] >= 5000) and np.all(slen[~large] < 5000)
    first_of = np.zeros(n_scaf, bool)
    first_of[sc[rows["position"][present] == 0]] = True
    assert first_of.all()
    assert 0.3 < rows["direction"][present].mean() < 0.7


def test_digest_tables_merge_like_a_single_pass():
    """oracle over the whole library == merge of oracle runs over its slices (true start states)"""
    import oracle_lib
    oracle_lib.build()
    slices = _slices(3)
    whole = _concat(slices)
    rows, n_scaf, _ = synth.later_library_rows(slices[0].lengths.numpy(), 5000.0, seed=4)
    params = abi.make_params("rf", 11, 100.0, 3000.0, 500.0, 6000.0)
    want_res, _, fishy, _ = oracle_lib.graph_build(rows, n_scaf, params, whole)
    want = digest.apply_fishy(digest.edge_table(want_res), [fishy])
    assert np.array_equal(want["fishy"], want_res.fishy)
    tables, fdicts, base, halo = [], [], 0, (-1, -1)
    for s in slices:
        p = abi.make_params("rf", 11, 100.0, 3000.0, 500.0, 6000.0, halo=halo)
        res, _, f, _ = oracle_lib.graph_build(rows, n_scaf, p, s.to_batch())
        tables.append(digest.edge_table(res, first_base=base))
        fdicts.append(f)
        base += res.n_links
        if res.counters[abi.CNT_CALLS] > 0:
            halo = (int(res.counters[abi.CNT_LAST_OBS1]), int(res.counters[abi.CNT_LAST_OBS2]))
    got = digest.apply_fishy(digest.merge_slices(tables), fdicts)
    rep = digest.compare(got, want)
    assert rep["integers_bit_exact"], rep
    assert (got["parts"] > 1).sum() > 0   # edges with links on both sides of a cut
    # a single flipped observation is detected
    res.obs_u[res.n_links // 2] += 1
    tables[-1] = digest.edge_table(res, first_base=base - res.n_links)
    assert not digest.compare(digest.apply_fishy(digest.merge_slices(tables), fdicts), want)["integers_bit_exact"]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    import torch.distributed as dist
    import slice_oracle
    from besst_b200.dist import DistributedGraphBuild
    from dist_backend_numpy import NumpyBackend
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    try:
        lib = synth.make_library_slice(N_CONTIGS, N_PAIRS, "rf", 3000.0, 500.0, 0.25, SEED, rank, world)
        rows, n_scaf, n_large = synth.later_library_rows(lib.lengths.numpy(), 5000.0, seed=4)
        params = abi.make_params("rf", 11, 100.0, 3000.0, 500.0, 6000.0)
        batch = lib.to_batch()
        runner = DistributedGraphBuild(NumpyBackend(SimpleNamespace(rows=rows, n_scaffolds=n_scaf, n_large_scaffolds=n_large)), rank, world)
        runner.step(params, batch)
        got = slice_oracle.gather_owned(dist, rank, world, runner.fetch_local())
        want = slice_oracle.sliced_oracle(dist, rank, world, rows, n_scaf, params, batch)
        if rank == 0:
            rep = slice_oracle.compare(got, want)
            assert rep["integers_bit_exact"], rep
            assert rep["multi_slice_edges"] > 0 and rep["links"] > 10000
            open(os.path.join(out_dir, "ok"), "w").write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_distributed_build_of_a_global_library_equals_the_sliced_oracle(tmp_path, world):
    import oracle_lib
    oracle_lib.build()
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert os.path.exists(os.path.join(str(tmp_path), "ok"))
