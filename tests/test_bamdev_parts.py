"""The multi-GPU BAM ingest: besst_bam_ingest_part (rank r of n inflates and decodes its part of the file) and the
neighbour check of besst_b200.dist.ingest_bam_distributed.

CPU: the part logic is the window loop of bam_ingest.hpp, run here through the host rendering of the kernels
(libbesst_bgzf_hostcheck.so, test tooling): the concatenation of the parts equals the whole file for 1..16 parts, ragged
files with records straddling BGZF blocks, windows and part boundaries, header longer than a part; the virtual-offset
chain is consistent; with blind seeds (every guess wrong where a block starts inside a record) the world-2/3 gloo run of
ingest_bam_distributed repairs the parts.  GPU (-m gpu): the same through the C ABI on one device, part after part."""
import ctypes as C
import os
import socket
import sys

import numpy as np
import pytest

import test_bamdev as t
import test_bamio as tb
from besst_b200 import bamio, synth

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


class _HostPart(object):
    def __init__(self, cols, first, landing):
        self.cols, self.first_voffset, self.landing_voffset = cols, first, landing

    def __len__(self):
        return int(self.cols["tid"].shape[0])


class HostcheckEngine(object):
    """stand-in for CudaEngine.ingest_bam over the host rendering of the ingest (tests only)"""

    def __init__(self, window=0, carry=0, tail=0):
        self.window, self.carry, self.tail = window, carry, tail
        self.calls = []

    def ingest_bam(self, path, part=(0, 1), start_voffset=-1, blind_seeds=False, head_records=1000, check_crc=True):
        L = t.hostcheck()
        L.bgzf_hc_ingest_part.restype = C.c_void_p
        L.bgzf_hc_ingest_part.argtypes = [C.c_char_p, C.c_int64, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int64, C.c_int, C.c_int,
                                          C.c_int64, C.c_int64, C.c_char_p, C.c_int]
        err = C.create_string_buffer(512)
        h = L.bgzf_hc_ingest_part(os.fsencode(path), self.window, 0, self.carry, int(check_crc), int(blind_seeds), head_records, part[0], part[1],
                                  start_voffset, self.tail, err, 512)
        if not h:
            raise IOError(err.value.decode())
        n = L.bgzf_hc_n(h)
        cols = {}
        for i, (k, dt) in enumerate(t.COLS):
            p = L.bgzf_hc_column(h, i)
            cols[k] = np.frombuffer((C.c_char * (n * np.dtype(dt).itemsize)).from_address(p), dtype=dt, count=n).copy() if n else np.zeros(0, dt)
        out = _HostPart(cols, int(L.bgzf_hc_stat(h, 7)), int(L.bgzf_hc_stat(h, 8)))
        L.bgzf_hc_close(h)
        self.calls.append((part, start_voffset))
        return out


class _PartEngine(object):
    """what the entry points need from an engine under a process group, without a GPU: ingest_bam over the host rendering of
    the device ingest (a RecordBatch of the part's columns + its virtual offsets), library metrics from the C oracle, the
    numpy backend for the distributed build"""

    def __init__(self):
        from oracle_engine import OracleEngine
        self._o = OracleEngine()
        self.libmetrics = self._o.libmetrics
        self.gapest_batch = self._o.gapest_batch
        self.ingest_calls = 0

    def graph_build(self, *a, **k):
        raise AssertionError("PE must take the distributed build when the process group has more than one rank")

    def make_dist_backend(self, table):
        from dist_backend_numpy import NumpyBackend
        return NumpyBackend(table)

    def ingest_bam(self, path, part=(0, 1), start_voffset=-1, **kw):
        from besst_b200.records import RecordBatch
        L = t.hostcheck()
        L.bgzf_hc_ingest_part.restype = C.c_void_p
        L.bgzf_hc_ingest_part.argtypes = [C.c_char_p, C.c_int64, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int64, C.c_int, C.c_int,
                                          C.c_int64, C.c_int64, C.c_char_p, C.c_int]
        err = C.create_string_buffer(512)
        h = L.bgzf_hc_ingest_part(os.fsencode(path), 0, 0, 0, 1, 0, 1000, part[0], part[1], start_voffset, 0, err, 512)
        if not h:
            raise IOError(err.value.decode())
        n, nh = L.bgzf_hc_n(h), L.bgzf_hc_n_head(h)

        def col(i, dt, count):
            p = L.bgzf_hc_column(h, i)
            return np.frombuffer((C.c_char * (count * np.dtype(dt).itemsize)).from_address(p), dtype=dt, count=count).copy() if count else np.zeros(0, dt)
        cols = {k: col(i, dt, n) for i, (k, dt) in enumerate(t.COLS[:8])}
        rlen, alen = np.zeros(n, np.int32), np.zeros(n, np.int32)
        rlen[:nh], alen[:nh] = col(9, np.int32, nh), col(10, np.int32, nh)
        batch = RecordBatch(references=[L.bgzf_hc_ref_name(h, i).decode() for i in range(L.bgzf_hc_n_refs(h))],
                            lengths=[int(L.bgzf_hc_ref_length(h, i)) for i in range(L.bgzf_hc_n_refs(h))], rlen=rlen, alen=alen,
                            packed=col(8, np.uint32, n), **cols)
        batch.first_voffset, batch.landing_voffset = int(L.bgzf_hc_stat(h, 7)), int(L.bgzf_hc_stat(h, 8))
        L.bgzf_hc_close(h)
        self.ingest_calls += 1
        return batch


def _ragged_file(path, n, block_bytes, n_refs=37):
    rng = np.random.default_rng(n + block_bytes)
    refs = [("c%d,pos:%d-%d,rc:0" % (i, i * 1000, i * 1000 + 900), 900 + i) for i in range(n_refs)]
    tb.write_bam(path, refs, tb._random_records(rng, n, len(refs)), block_bytes)
    return bamio.read_bam(path)


def _check_parts(parts, want):
    land = None
    for q in parts:
        if q.first_voffset >= 0 and land is not None:
            assert q.first_voffset == land
        if q.landing_voffset >= 0:
            land = q.landing_voffset
    for k, _ in t.COLS[:8]:
        assert np.array_equal(np.concatenate([q.cols[k] for q in parts]), getattr(want, k)), k


@pytest.mark.parametrize("n,block_bytes,n_refs", [(5000, 700, 37), (5000, 65000, 37), (20000, 3000, 37), (300, 3000, 37), (300, 700, 400)])
def test_parts_concatenate_to_the_whole_file(tmp_path, n, block_bytes, n_refs):
    """(300, 700, 400): the header is longer than several parts -- they own no record and say so"""
    path = str(tmp_path / "t.bam")
    want = _ragged_file(path, n, block_bytes, n_refs)
    for n_parts in (1, 2, 3, 8, 16):
        for window, carry in ((0, 0), (70000, 2048)):
            eng = HostcheckEngine(window, carry, tail=200000)
            parts = [eng.ingest_bam(path, part=(p, n_parts)) for p in range(n_parts)]
            _check_parts(parts, want)
            assert sum(len(q) for q in parts) == n


def test_a_record_longer_than_the_tail_is_an_error(tmp_path):
    path = str(tmp_path / "t.bam")
    _ragged_file(path, 5000, 700)
    eng = HostcheckEngine(tail=64)   # 64 bytes of tail: the record cut by the part boundary cannot be completed
    with pytest.raises(IOError, match="range tail"):
        for p in range(4):
            eng.ingest_bam(path, part=(p, 4))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, path, blind, out_dir):
    for p in (ROOT, os.path.join(ROOT, "oracle"), HERE):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch.distributed as dist
    from besst_b200.dist import ingest_bam_distributed
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    try:
        eng = HostcheckEngine(tail=200000)
        dev, info = ingest_bam_distributed(eng, path, rank, world, blind_seeds=blind)
        np.savez(os.path.join(out_dir, "part%d.npz" % rank), repeats=info["repeats"], record_base=info["record_base"],
                 counts=np.asarray(info["counts"]), calls=len(eng.calls), **dev.cols)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,blind", [(2, False), (3, True)])
def test_distributed_ingest_protocol_gloo(tmp_path, world, blind):
    """blind seeds: every part behind the first starts its chain at the first byte of its first block -- inside a record for
    this file -- so the neighbour check has to catch it and the part is read again from the previous part's landing"""
    import torch.multiprocessing as mp
    path = str(tmp_path / "t.bam")
    want = _ragged_file(path, 20000, 3000)
    mp.spawn(_worker, args=(world, _free_port(), path, blind, str(tmp_path)), nprocs=world, join=True)
    parts = [np.load(str(tmp_path / ("part%d.npz" % r))) for r in range(world)]
    for k, _ in t.COLS[:8]:
        assert np.array_equal(np.concatenate([q[k] for q in parts]), getattr(want, k)), k
    assert [int(q["record_base"]) for q in parts] == list(np.cumsum([0] + [len(q["tid"]) for q in parts[:-1]]))
    repeats = int(parts[0]["repeats"])
    assert all(int(q["repeats"]) == repeats for q in parts)
    assert (repeats > 0) == blind
    if blind:
        assert sum(int(q["calls"]) for q in parts) == world + repeats


def _entry_worker(rank, world, port, path, lengths, opts, overrides, out_dir):
    for p in (ROOT, os.path.join(ROOT, "oracle"), HERE):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["BESST_B200_INGEST"] = "device"
    import json
    import torch.distributed as dist
    import helpers
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    try:
        eng = _PartEngine()
        from besst_b200 import bamio as _bamio
        whole_file_reads = []
        real = _bamio.read_bam_native
        _bamio.read_bam_native = lambda *a, **k: (whole_file_reads.append(1), real(*a, **k))[1]
        out = helpers.run_dropin(None, opts, eng, fasta_lengths=lengths, bam_path=path, param_overrides=overrides)
        assert eng.ingest_calls >= 1
        json.dump(dict({k: out[k] for k in ("G", "G_prime", "param", "objects")}, whole_file_reads=len(whole_file_reads)),
                  open(os.path.join(out_dir, "rank%d.json" % rank), "w"))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("pairs,world,estimate", [(2400000, 2, True), (150000, 3, True)])
def test_entry_points_on_a_path_under_a_process_group_ingest_in_parts(tmp_path, pairs, world, estimate):
    """libmetrics.get_metrics + CreateGraph.PE on a BAM PATH in a gloo job with BESST_B200_INGEST=device: every rank ingests
    its part (host rendering of the device ingest), rank 0's library metrics are broadcast when its part covers the sampled
    prefix (2.4 M PE pairs: the 1e6-sample cap is reached after 2.27 M records, inside part 0 of 2) and the whole file is read on host threads otherwise
    (150 k pairs), PE builds from the parts -- every rank must end with the graphs, parameters and objects of a
    single-process run on the same file"""
    import json
    import torch.multiprocessing as mp
    import helpers
    import oracle_lib
    from oracle_engine import OracleEngine
    oracle_lib.build()
    big = pairs > 1000000
    orient, mu, sd = ("fr", 550.0, 50.0) if big else ("rf", 3000.0, 500.0)
    lib = synth.make_library(max(50, pairs // 10000), pairs, orient, mu, sd, 0.0, seed=17)
    batch = lib.to_batch()
    path = str(tmp_path / "lib.bam")
    bamio.write_bam_columns(path, batch, level=0 if pairs > 1000000 else 1, style="packed")
    lengths = dict(zip(batch.references, [int(x) for x in batch.lengths]))
    opts = dict(orientation=orient, mean=None if estimate else mu, stddev=None if estimate else sd, readlen=None if estimate else 100)
    overrides = {"no_score": True}   # the numpy backend of the distributed build has no scores
    os.environ["BESST_B200_INGEST"] = "host"
    try:
        from besst_b200 import records
        records._open_cache.clear()
        want = helpers.run_dropin(None, opts, OracleEngine(), fasta_lengths=lengths, bam_path=path, param_overrides=overrides)
    finally:
        os.environ.pop("BESST_B200_INGEST", None)
    want = json.loads(json.dumps({k: want[k] for k in ("G", "G_prime", "param", "objects")}))
    mp.spawn(_entry_worker, args=(world, _free_port(), path, lengths, opts, overrides, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        got = json.load(open(str(tmp_path / ("rank%d.json" % r))))
        for k in ("G", "G_prime", "param", "objects"):
            assert got[k] == want[k], (r, k)
        if estimate:   # the metrics came from rank 0's part when it holds the sampled prefix, from the whole file otherwise
            assert got["whole_file_reads"] == (0 if pairs > 1000000 else 1), (r, got["whole_file_reads"])
    assert len(want["G_prime"]["edges"]) > 0


# ---- GPU ------------------------------------------------------------------------------------------------------------------

@pytest.fixture(scope="module")
def engine():
    from besst_b200.engine import CudaEngine
    eng = CudaEngine(0)
    yield eng
    eng.close()


@pytest.mark.gpu
@pytest.mark.parametrize("n,block_bytes", [(5000, 700), (20000, 3000)])
def test_device_parts_concatenate_to_the_whole_file(tmp_path, engine, n, block_bytes):
    path = str(tmp_path / "t.bam")
    want = _ragged_file(path, n, block_bytes)
    for n_parts in (2, 3, 8):
        for env in ({"BESST_BAM_TAIL": 200000}, {"BESST_BAM_TAIL": 200000, "BESST_BAM_WINDOW": 70000, "BESST_BAM_CARRY": 2048}):
            parts = []
            for p in range(n_parts):
                dev = t._with_env(env, lambda: engine.ingest_bam(path, part=(p, n_parts)))
                host = dev.to_host()
                parts.append(_HostPart({k: getattr(host, k) for k, _ in t.COLS[:8]}, dev.first_voffset, dev.landing_voffset))
            _check_parts(parts, want)


@pytest.mark.gpu
def test_device_parts_of_a_synthetic_library_and_a_forced_start(tmp_path, engine):
    lib = synth.make_library(200, 150000, "rf", 3000.0, 500.0, 0.0, seed=13)
    batch = lib.to_batch()
    path = str(tmp_path / "lib.bam")
    bamio.write_bam_columns(path, batch, style="packed")
    parts = []
    for p in range(4):
        dev = engine.ingest_bam(path, part=(p, 4))
        host = dev.to_host()
        parts.append(_HostPart({k: getattr(host, k) for k, _ in t.COLS[:8]}, dev.first_voffset, dev.landing_voffset))
    _check_parts(parts, t.batch_with_lengths(batch))
    # blind seeds put part 2's first record in the wrong place; the previous part's landing as start_voffset repairs it
    wrong = engine.ingest_bam(path, part=(2, 4), blind_seeds=True)
    assert wrong.first_voffset != parts[1].landing_voffset
    right = engine.ingest_bam(path, part=(2, 4), blind_seeds=True, start_voffset=parts[1].landing_voffset)
    assert right.first_voffset == parts[1].landing_voffset and len(right) == len(parts[2])
    assert np.array_equal(right.to_host().pos, parts[2].cols["pos"])
