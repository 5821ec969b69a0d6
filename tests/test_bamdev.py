"""The device BAM ingest (besst_bam_ingest: BGZF inflate, CRC-32, record boundaries, record decode on the GPU).

CPU part: the SAME source the GPU runs (besst_b200/csrc/bgzf_core.cuh + bam_ingest.hpp), compiled for the host with a
lane-serial warp policy (libbesst_bgzf_hostcheck.so, test tooling), against zlib -- every deflate block type, every stream
alignment, the 32-lane CRC -- and against the pure-Python BAM reader on files with records that straddle BGZF blocks and
windows, with tiny windows (carry path), with blind seeds (every block's record hop repaired by the host verification),
on corrupt and truncated files, and on the reference's own testdata where present.

GPU part (-m gpu): CudaEngine.ingest_bam through the C ABI against the same readers, the ingested columns fed to
besst_libmetrics / besst_graph_build without leaving HBM, and the drop-in entry points on a BAM path."""
import ctypes as C
import os
import zlib

import numpy as np
import pytest

import helpers
import test_bamio as tb
from besst_b200 import bamio, build, synth

REF_BAM = "/root/reference/testdata/testset1/mapped.bam"
COLS = [("tid", np.int32), ("mtid", np.int32), ("pos", np.int32), ("mpos", np.int32), ("tlen", np.int32), ("qlen", np.int32),
        ("flag", np.uint16), ("mapq", np.uint8), ("packed", np.uint32)]

_hc = None


def hostcheck():
    global _hc
    if _hc is None:
        L = C.CDLL(build.build_hostcheck())
        L.bgzf_hc_inflate.argtypes = [C.c_char_p, C.c_uint32, C.c_int, C.c_void_p, C.c_uint32]
        L.bgzf_hc_set_width.argtypes = [C.c_int]
        L.bgzf_hc_crc32.restype = C.c_uint32
        L.bgzf_hc_crc32.argtypes = [C.c_char_p, C.c_uint32, C.c_int]
        L.bgzf_hc_ingest.restype = C.c_void_p
        L.bgzf_hc_ingest.argtypes = [C.c_char_p, C.c_int64, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int64, C.c_char_p, C.c_int]
        for f in ("bgzf_hc_n", "bgzf_hc_n_head", "bgzf_hc_n_refs"):
            getattr(L, f).restype = C.c_int64
            getattr(L, f).argtypes = [C.c_void_p]
        L.bgzf_hc_ref_length.restype = C.c_int64
        L.bgzf_hc_ref_length.argtypes = [C.c_void_p, C.c_int64]
        L.bgzf_hc_stat.restype = C.c_int64
        L.bgzf_hc_stat.argtypes = [C.c_void_p, C.c_int]
        L.bgzf_hc_ref_name.restype = C.c_char_p
        L.bgzf_hc_ref_name.argtypes = [C.c_void_p, C.c_int64]
        L.bgzf_hc_column.restype = C.c_void_p
        L.bgzf_hc_column.argtypes = [C.c_void_p, C.c_int]
        L.bgzf_hc_close.argtypes = [C.c_void_p]
        _hc = L
    return _hc


def host_ingest(path, window=0, max_inflated=0, carry=0, crc=True, blind=False, head=1000):
    """the window loop + kernels' code, lane-serial on the host -> dict of columns, header, stats"""
    L = hostcheck()
    err = C.create_string_buffer(512)
    h = L.bgzf_hc_ingest(os.fsencode(path), window, max_inflated, carry, int(crc), int(blind), head, err, 512)
    if not h:
        raise IOError(err.value.decode())
    n, nh = L.bgzf_hc_n(h), L.bgzf_hc_n_head(h)
    out = {}
    for i, (k, dt) in enumerate(COLS):
        p = L.bgzf_hc_column(h, i)
        out[k] = np.frombuffer((C.c_char * (n * np.dtype(dt).itemsize)).from_address(p), dtype=dt, count=n).copy() if n else np.zeros(0, dt)
    for i, k in ((9, "rlen"), (10, "alen")):
        p = L.bgzf_hc_column(h, i)
        out[k] = np.frombuffer((C.c_char * (nh * 4)).from_address(p), dtype=np.int32, count=nh).copy() if nh else np.zeros(0, np.int32)
    out["references"] = [L.bgzf_hc_ref_name(h, i).decode() for i in range(L.bgzf_hc_n_refs(h))]
    out["lengths"] = [int(L.bgzf_hc_ref_length(h, i)) for i in range(L.bgzf_hc_n_refs(h))]
    out["stats"] = dict(zip(("compressed_bytes", "uncompressed_bytes", "blocks", "records", "windows", "rescans", "unpackable"),
                            (int(L.bgzf_hc_stat(h, i)) for i in range(7))))
    L.bgzf_hc_close(h)
    return out


def assert_columns_equal(got, want):
    """got: dict of columns (host_ingest) or a RecordBatch; want: RecordBatch"""
    g = got if isinstance(got, dict) else {k: getattr(got, k) for k, _ in COLS[:8]} | {
        "references": list(got.references), "lengths": list(got.lengths), "rlen": got.rlen, "alen": got.alen, "packed": got.packed}
    assert list(g["references"]) == list(want.references) and [int(x) for x in g["lengths"]] == [int(x) for x in want.lengths]
    for k, _ in COLS[:8]:
        assert np.array_equal(g[k], getattr(want, k)), k
    nh = min(len(g["rlen"]), 1000, len(want.rlen))
    assert np.array_equal(g["rlen"][:nh], want.rlen[:nh]) and np.array_equal(g["alen"][:nh], want.alen[:nh])


# ---- CPU: the device decoder's source, lane-serial, against zlib ------------------------------------------------------

def _payloads():
    rng = np.random.default_rng(7)
    record_like = (b"ACGT" * 7 + b"read12345\0" + bytes(rng.integers(30, 42, 100, dtype=np.uint8))) * 300
    return [b"", b"a", b"abc" * 1000, bytes(rng.integers(0, 256, 65000, dtype=np.uint8)), bytes(rng.integers(0, 4, 65536, dtype=np.uint8)),
            b"\0" * 65536, bytes(rng.integers(0, 256, 127, dtype=np.uint8)), bytes(rng.integers(0, 256, 128, dtype=np.uint8)),
            bytes(rng.integers(0, 256, 129, dtype=np.uint8)), record_like[:65536],
            bytes(np.repeat(rng.integers(0, 256, 300, dtype=np.uint8), rng.integers(1, 400, 300)))[:65536]]


@pytest.mark.parametrize("level", [0, 1, 6, 9])
def test_inflate_and_crc_equal_zlib(level):
    """stored (level 0), fixed and dynamic Huffman blocks, matches longer than their distance, every byte alignment of the
    stream inside its word buffer; the 32-lane Horner CRC-32 with and without a tail"""
    L = hostcheck()
    for data in _payloads():
        co = zlib.compressobj(level, zlib.DEFLATED, -15)
        c = co.compress(data) + co.flush()
        for mis in range(4):
            out = np.zeros(len(data) + 8, np.uint8)
            assert L.bgzf_hc_inflate(c, len(c), mis, out.ctypes.data, len(data)) == 0
            assert out[:len(data)].tobytes() == data
            assert L.bgzf_hc_crc32(data, len(data), mis) == (zlib.crc32(data) & 0xffffffff)


def test_inflate_fuzz_against_zlib():
    """400 random streams: every zlib level, window size 2^9..2^15, memLevel, the strategies (default, filtered, Huffman-only,
    RLE, fixed codes), SYNC / FULL flushes in the middle (empty stored blocks, several deflate blocks per stream), payloads
    from incompressible to one repeated byte; placement rounds of 32 and of 8 lanes"""
    L = hostcheck()
    rng = np.random.default_rng(123)

    def payload():
        kind, n = int(rng.integers(6)), int(rng.integers(1, 65537))
        if kind == 0:
            return bytes(rng.integers(0, 256, n, dtype=np.uint8))
        if kind == 1:
            return bytes(rng.integers(0, int(rng.integers(2, 20)), n, dtype=np.uint8))
        if kind == 2:
            return bytes(np.repeat(rng.integers(0, 256, max(1, n // 50), dtype=np.uint8), 50))[:n]
        if kind == 3:
            return (bytes(rng.integers(0, 256, int(rng.integers(1, 300)), dtype=np.uint8)) * 70000)[:n]
        if kind == 4:
            return b"".join(b"read%07d\0" % i + bytes(rng.integers(33, 74, 40, dtype=np.uint8)) for i in range(n // 52 + 1))[:n]
        return bytes(rng.choice(np.array([0, 255, 17, 34], np.uint8), n, p=[0.7, 0.1, 0.1, 0.1]))
    try:
        for trial in range(400):
            L.bgzf_hc_set_width(32 if trial % 2 else 8)
            data = payload()
            strategy = [zlib.Z_DEFAULT_STRATEGY, zlib.Z_FILTERED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE, zlib.Z_FIXED][int(rng.integers(5))]
            co = zlib.compressobj(int(rng.integers(0, 10)), zlib.DEFLATED, -int(rng.integers(9, 16)), int(rng.integers(1, 10)), strategy)
            parts, prev = [], 0
            for c in sorted(set(int(x) for x in rng.integers(0, len(data) + 1, int(rng.integers(0, 4))))):
                parts.append(co.compress(data[prev:c]))
                prev = c
                if rng.random() < 0.7:
                    parts.append(co.flush([zlib.Z_SYNC_FLUSH, zlib.Z_FULL_FLUSH, zlib.Z_NO_FLUSH][int(rng.integers(3))]))
            parts += [co.compress(data[prev:]), co.flush()]
            c = b"".join(parts)
            out = np.zeros(len(data) + 8, np.uint8)
            mis = int(rng.integers(4))
            assert L.bgzf_hc_inflate(c, len(c), mis, out.ctypes.data, len(data)) == 0, trial
            assert out[:len(data)].tobytes() == data, trial
            assert L.bgzf_hc_crc32(data, len(data), mis) == (zlib.crc32(data) & 0xffffffff)
    finally:
        L.bgzf_hc_set_width(32)


def test_inflate_rejects_damaged_streams():
    L = hostcheck()
    rng = np.random.default_rng(3)
    data = bytes(rng.integers(0, 64, 40000, dtype=np.uint8))
    co = zlib.compressobj(6, zlib.DEFLATED, -15)
    c = co.compress(data) + co.flush()
    out = np.zeros(len(data) + 8, np.uint8)
    assert L.bgzf_hc_inflate(c, len(c), 0, out.ctypes.data, len(data) - 1) != 0      # wrong ISIZE
    assert L.bgzf_hc_inflate(c[:len(c) // 2], len(c) // 2, 0, out.ctypes.data, len(data)) != 0   # truncated stream
    bad = 0
    for trial in range(200):   # random damage never crashes, never passes silently: an error code or a CRC mismatch
        d = bytearray(c)
        d[int(rng.integers(0, len(d)))] ^= 1 << int(rng.integers(0, 8))
        rc = L.bgzf_hc_inflate(bytes(d), len(d), int(rng.integers(0, 4)), out.ctypes.data, len(data))
        if rc != 0 or out[:len(data)].tobytes() != data:
            bad += 1
            assert rc != 0 or L.bgzf_hc_crc32(out[:len(data)].tobytes(), len(data), 0) != (zlib.crc32(data) & 0xffffffff)
    assert bad >= 150


@pytest.mark.parametrize("n,block_bytes", [(0, 3000), (1, 3000), (5000, 700), (5000, 65000), (20000, 3000)])
def test_host_rendered_ingest_equals_python_reader(tmp_path, n, block_bytes):
    """ragged CIGARs, soft / hard clips, SEQ '*', unmapped records, BGZF blocks cut anywhere (records straddle blocks AND
    windows): one big window, tiny windows with the carry path, blind seeds repaired by the verification"""
    rng = np.random.default_rng(n + block_bytes)
    refs = [("c%d,pos:%d-%d,rc:0" % (i, i * 1000, i * 1000 + 900), 900 + i) for i in range(37)]
    path = str(tmp_path / "t.bam")
    tb.write_bam(path, refs, tb._random_records(rng, n, len(refs)), block_bytes)
    py = bamio.read_bam(path)
    for window, carry, blind in [(0, 0, False), (8192, 4096, False), (70000, 1024, False), (70000, 1024, True), (0, 0, True)]:
        if window and window < block_bytes:
            continue
        got = host_ingest(path, window, 0, carry, blind=blind)
        assert_columns_equal(got, py)
        assert got["stats"]["records"] == n
        if window and n >= 5000:
            assert got["stats"]["windows"] >= 2
        if blind and n >= 5000 and block_bytes < 65000:
            assert got["stats"]["rescans"] > 0     # most blocks start inside a record: every such seed was repaired
        if not blind:
            assert got["stats"]["rescans"] == 0    # the plausibility seeds were all right
        if n:
            assert np.array_equal(got["packed"], py.flag.astype(np.uint32) | (py.mapq.astype(np.uint32) << 12) | (py.qlen.astype(np.uint32) << 20))


def _long_record_file(path, block_bytes):
    """400 records of which five are 150-400 KB long (long reads): each spans several BGZF blocks, whole blocks lie inside a
    record, flag / qlen do not fit the packed column"""
    rng = np.random.default_rng(3)
    refs = [("c%d" % i, 100000 + i) for i in range(9)]
    recs = tb._random_records(rng, 400, len(refs))
    for k in (5, 120, 121, 300, 399):
        n = int(rng.integers(150000, 400000))
        recs[k] = dict(tid=1, pos=100 + k, mapq=33, flag=99, l_seq=n, mtid=2, mpos=5, tlen=1234, cigar=[(4, 10), (0, n - 30), (4, 20)], name="long%d" % k)
    tb.write_bam(path, refs, recs, block_bytes)
    return bamio.read_bam(path)


@pytest.mark.parametrize("block_bytes", [65280, 3000])
def test_host_rendered_ingest_of_records_longer_than_a_block(tmp_path, block_bytes):
    path = str(tmp_path / "long.bam")
    py = _long_record_file(path, block_bytes)
    for window, carry, blind in [(0, 0, False), (0, 0, True), (2048, 1 << 20, False), (2048, 1 << 20, True)]:
        got = host_ingest(path, window, 0, carry, blind=blind)
        assert_columns_equal(got, py)
        assert got["stats"]["unpackable"] == 5
        assert (got["stats"]["rescans"] > 0) == blind
        if window:
            assert got["stats"]["windows"] >= 2
    with pytest.raises(IOError, match="carry buffer"):
        host_ingest(path, 2048, 0, 4096)    # a 4 KB carry cannot hold the head of a 150 KB record


def test_host_rendered_ingest_errors(tmp_path):
    rng = np.random.default_rng(5)
    refs = [("c%d" % i, 1000 + i) for i in range(5)]
    path = str(tmp_path / "t.bam")
    tb.write_bam(path, refs, tb._random_records(rng, 3000, len(refs)), 3000)
    raw = open(path, "rb").read()
    # a flipped payload byte: CRC mismatch or a deflate error, never silence
    d = bytearray(raw)
    d[len(d) // 2] ^= 0x10
    open(path, "wb").write(bytes(d))
    with pytest.raises(IOError, match="CRC32|inflate|BGZF"):
        host_ingest(path)
    # file cut inside a BGZF block
    open(path, "wb").write(raw[:len(raw) // 2])
    with pytest.raises(IOError, match="truncated"):
        host_ingest(path)
    # whole blocks, but the last record is cut
    blocks, o = [], 0
    while o < len(raw):
        bsize = int.from_bytes(raw[o + 16:o + 18], "little") + 1
        blocks.append(raw[o:o + bsize])
        o += bsize
    open(path, "wb").write(b"".join(blocks[:len(blocks) // 2]))
    with pytest.raises(IOError, match="partial record"):
        host_ingest(path)
    open(path, "wb").write(b"not a bam file at all" * 10)
    with pytest.raises(IOError, match="BGZF"):
        host_ingest(path)


@pytest.mark.parametrize("style", ["htslib", "packed"])
def test_host_rendered_ingest_of_a_synthetic_library(tmp_path, style):
    """write_bam_columns (the bench's file writer) round-trips through the host-thread reader and the device ingest's
    source: both BGZF writer styles, many windows"""
    lib = synth.make_library(60, 20000, "rf", 3000.0, 500.0, 0.0, seed=11)
    batch = lib.to_batch()
    path = str(tmp_path / "lib.bam")
    bamio.write_bam_columns(path, batch, style=style)
    nat = bamio.read_bam_native(path)
    assert_columns_equal(nat, batch_with_lengths(batch))
    got = host_ingest(path, 1 << 18, 1 << 20, 1 << 12)
    assert_columns_equal(got, nat)
    assert got["stats"]["windows"] > 3 and got["stats"]["unpackable"] == 0
    assert np.array_equal(got["packed"], nat.packed)


def batch_with_lengths(batch):
    """the writer emits l_seq = 100 and a CIGAR spanning qlen reference bases"""
    import copy
    b = copy.copy(batch)
    b.rlen = np.full(len(batch), 100, np.int32)
    b.alen = np.asarray(batch.qlen, np.int32)
    return b


@pytest.mark.skipif(not os.path.exists(REF_BAM), reason="reference testdata not present")
def test_host_rendered_ingest_of_the_reference_testdata():
    """all 9187 BGZF blocks / 1,999,958 records of testset1 (written by samtools): inflate == zlib (CRC checked per block),
    columns == the host-thread reader"""
    nat = bamio.read_bam_native(REF_BAM)
    got = host_ingest(REF_BAM, 4 << 20, 24 << 20, 1 << 16)
    assert_columns_equal(got, nat)
    assert got["stats"]["records"] == 1999958 and got["stats"]["rescans"] == 0 and got["stats"]["windows"] > 10
    assert np.array_equal(got["packed"], nat.packed)


# ---- GPU: besst_bam_ingest through the C ABI ----------------------------------------------------------------------------

@pytest.fixture(scope="module")
def engine():
    from besst_b200.engine import CudaEngine
    eng = CudaEngine(0)
    yield eng
    eng.close()


def _with_env(env, fn):
    old = {k: os.environ.get(k) for k in env}
    os.environ.update({k: str(v) for k, v in env.items()})
    try:
        return fn()
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


@pytest.mark.gpu
@pytest.mark.parametrize("n,block_bytes", [(0, 3000), (1, 3000), (5000, 700), (5000, 65000), (20000, 3000)])
def test_device_ingest_equals_python_reader(tmp_path, engine, n, block_bytes):
    rng = np.random.default_rng(n + block_bytes)
    refs = [("c%d,pos:%d-%d,rc:0" % (i, i * 1000, i * 1000 + 900), 900 + i) for i in range(37)]
    path = str(tmp_path / "t.bam")
    tb.write_bam(path, refs, tb._random_records(rng, n, len(refs)), block_bytes)
    py = bamio.read_bam(path)
    for env, blind in [({}, False), ({"BESST_BAM_WINDOW": 70000, "BESST_BAM_CARRY": 1024}, False),
                       ({"BESST_BAM_WINDOW": 70000, "BESST_BAM_CARRY": 1024}, True), ({}, True)]:
        dev = _with_env(env, lambda: engine.ingest_bam(path, blind_seeds=blind))
        assert len(dev) == n and dev.stats["records"] == n and dev.stats["crc_checked"] == 1
        assert_columns_equal(dev.to_host(), py)
        if env and n >= 5000:
            assert dev.stats["windows"] >= 2
        if not blind:
            assert dev.stats["rescans"] == 0
        elif n >= 5000 and block_bytes < 65000:
            assert dev.stats["rescans"] > 0


@pytest.mark.gpu
@pytest.mark.parametrize("block_bytes", [65280, 3000])
def test_device_ingest_of_records_longer_than_a_block(tmp_path, engine, block_bytes):
    from besst_b200._lib import BesstLibraryError
    path = str(tmp_path / "long.bam")
    py = _long_record_file(path, block_bytes)
    for env, blind in [({}, False), ({}, True), ({"BESST_BAM_WINDOW": 2048, "BESST_BAM_CARRY": 1 << 20}, False),
                       ({"BESST_BAM_WINDOW": 2048, "BESST_BAM_CARRY": 1 << 20}, True)]:
        dev = _with_env(env, lambda: engine.ingest_bam(path, blind_seeds=blind))
        assert_columns_equal(dev.to_host(), py)
        assert not dev.abi_records.packed      # flag | mapq | qlen does not fit 32 bits for the long records
        assert (dev.stats["rescans"] > 0) == blind
    with pytest.raises(BesstLibraryError, match="carry buffer"):
        _with_env({"BESST_BAM_WINDOW": 2048, "BESST_BAM_CARRY": 4096}, lambda: engine.ingest_bam(path))


@pytest.mark.gpu
def test_device_ingest_errors(tmp_path, engine):
    from besst_b200._lib import BesstLibraryError
    rng = np.random.default_rng(5)
    refs = [("c%d" % i, 1000 + i) for i in range(5)]
    path = str(tmp_path / "t.bam")
    tb.write_bam(path, refs, tb._random_records(rng, 3000, len(refs)), 3000)
    raw = open(path, "rb").read()
    d = bytearray(raw)
    d[len(d) // 2] ^= 0x10
    open(path, "wb").write(bytes(d))
    with pytest.raises(BesstLibraryError, match="CRC32|inflate|BGZF"):
        engine.ingest_bam(path)
    open(path, "wb").write(raw[:len(raw) // 2])
    with pytest.raises(BesstLibraryError, match="truncated"):
        engine.ingest_bam(path)
    with pytest.raises(BesstLibraryError, match="cannot open"):
        engine.ingest_bam(str(tmp_path / "missing.bam"))
    open(path, "wb").write(raw)   # and the engine is still usable afterwards
    assert len(engine.ingest_bam(path)) == 3000


@pytest.mark.gpu
@pytest.mark.parametrize("style", ["htslib", "packed"])
def test_device_ingest_feeds_the_graph_build_without_leaving_hbm(tmp_path, engine, style):
    """file -> besst_bam_ingest -> besst_libmetrics + besst_graph_build on the device-resident columns == the same calls
    on the host-decoded records; several windows; the result equals the oracle's"""
    import oracle_lib
    from besst_b200 import abi
    lib = synth.make_config("small_mp")
    batch = lib.to_batch()
    path = str(tmp_path / "lib.bam")
    bamio.write_bam_columns(path, batch, style=style)
    dev = _with_env({"BESST_BAM_WINDOW": 4 << 20, "BESST_BAM_MAX_INFLATED": 12 << 20}, lambda: engine.ingest_bam(path))
    assert len(dev) == len(batch) and dev.stats["windows"] > 3 and dev.abi_records.packed
    host = dev.to_host()
    assert_columns_equal(host, batch_with_lengths(batch))
    params = abi.make_params("rf", 11, 100.0, lib.mu, lib.sigma, lib.mu + 6 * lib.sigma)
    objs = helpers.first_library_objects(batch.references, batch.lengths, lib.mu + 4 * lib.sigma)
    table = helpers.table_for(batch, objs)
    got = engine.graph_build(table, params, dev)
    want, _, _, consistent = oracle_lib.graph_build(table.rows, table.n_scaffolds, params, batch)
    assert consistent
    helpers.assert_graph_equal(got, want, label="device ingest -> graph build")
    from besst_b200.libmetrics import metric_rows
    p0 = abi.make_params("rf", 11, 100.0, 0.0, 0.0, 0.0)
    rc_d, m_d, adj_d = engine.libmetrics(metric_rows(batch.lengths), p0, dev, batch.lengths, True)
    rc_h, m_h, adj_h = engine.libmetrics(metric_rows(batch.lengths), p0, batch, batch.lengths, True)
    assert rc_d == rc_h and np.array_equal(adj_d, adj_h)
    for f, _ in abi.LibMetricsOut._fields_:
        assert getattr(m_d, f) == getattr(m_h, f), f


@pytest.mark.gpu
@pytest.mark.parametrize("level", [0, 9])
def test_device_ingest_of_stored_and_best_compression_files(tmp_path, engine, level):
    """level 0: every BGZF block is a stored deflate block (`samtools view -u`): the warp-wide copy from the compressed
    stream; level 9: long matches, lazy matching"""
    lib = synth.make_library(80, 40000, "rf", 3000.0, 500.0, 0.0, seed=19)
    batch = lib.to_batch()
    path = str(tmp_path / "lib.bam")
    bamio.write_bam_columns(path, batch, level=level, style="packed")
    dev = _with_env({"BESST_BAM_WINDOW": 1 << 20}, lambda: engine.ingest_bam(path))
    assert dev.stats["windows"] > 1
    assert_columns_equal(dev.to_host(), batch_with_lengths(batch))


@pytest.mark.gpu
def test_entry_points_on_a_bam_path_with_device_ingest(tmp_path, engine):
    """libmetrics.get_metrics + CreateGraph.PE on a PATH: device ingest == host-thread ingest, graph for graph"""
    lib = synth.make_config("small_mp")
    batch = lib.to_batch()
    path = str(tmp_path / "lib.bam")
    bamio.write_bam_columns(path, batch)
    opts = dict(orientation="rf", mean=None, stddev=None, readlen=None)
    sigs = {}
    for mode in ("host", "device"):
        from besst_b200 import records
        records._open_cache.clear()
        out = _with_env({"BESST_B200_INGEST": mode}, lambda: helpers.run_dropin(batch, opts, engine, bam_path=path))
        sigs[mode] = (out["G"], out["G_prime"], out["param"], out["objects"])
    assert sigs["host"] == sigs["device"]
    assert len(sigs["device"][0]["edges"]) > 0


@pytest.mark.gpu
def test_device_ingest_throughput_report(tmp_path, engine, capsys):
    """not an assertion on speed: prints what the bench reports (records/s, inflated GB/s) for a 1.2 M-record file"""
    lib = synth.make_library(3000, 600000, "rf", 3000.0, 500.0, 0.0, seed=3)
    batch = lib.to_batch()
    path = str(tmp_path / "lib.bam")
    bamio.write_bam_columns(path, batch)
    engine.ingest_bam(path)   # warm-up: page cache, allocations
    dev = engine.ingest_bam(path)
    s = dev.stats
    with capsys.disabled():
        print("\ndevice ingest: %d records, %.1f MB -> %.1f MB in %.1f ms wall (read %.1f ms; inflate %.2f ms = %.1f GB/s inflated, scan %.2f ms, decode %.2f ms)" % (
            s["records"], s["compressed_bytes"] / 1e6, s["uncompressed_bytes"] / 1e6, 1e3 * s["seconds_total"], 1e3 * s["seconds_read"],
            s["ms_inflate"], s["uncompressed_bytes"] / 1e6 / max(s["ms_inflate"], 1e-9), s["ms_scan"], s["ms_decode"]))
    assert_columns_equal(dev.to_host(), batch_with_lengths(batch))
