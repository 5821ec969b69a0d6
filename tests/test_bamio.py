"""CPU: the native BAM ingest (libbesst_bamio.so: threaded BGZF inflate + fixed-core decode) against the
pure-Python reader, on BAM files written here (ragged CIGARs, soft/hard clips, SEQ '*', unmapped
records, many small BGZF blocks, records straddling blocks and inflate windows) and, where the
reference tree is present, on its own testdata."""
import os
import struct
import zlib

import numpy as np
import pytest

from besst_b200 import bamio
from besst_b200.abi import pack_record_columns as abi_pack

REF_BAM = "/root/reference/testdata/testset1/mapped.bam"


def _bgzf_block(data):
    comp = zlib.compressobj(6, zlib.DEFLATED, -15)
    c = comp.compress(data) + comp.flush()
    bsize = 12 + 6 + len(c) + 8
    hdr = b"\x1f\x8b\x08\x04" + b"\0\0\0\0" + b"\0\xff" + struct.pack("<H", 6) + b"BC" + struct.pack("<HH", 2, bsize - 1)
    return hdr + c + struct.pack("<II", zlib.crc32(data) & 0xffffffff, len(data))


def write_bam(path, refs, records, block_bytes=3000):
    """records: dicts with tid, pos, mapq, flag, l_seq, mtid, mpos, tlen, cigar [(op, len)], name"""
    text = b"@HD\tVN:1.0\tSO:coordinate\n"
    raw = b"BAM\x01" + struct.pack("<i", len(text)) + text + struct.pack("<i", len(refs))
    for name, ln in refs:
        nm = name.encode() + b"\0"
        raw += struct.pack("<i", len(nm)) + nm + struct.pack("<i", ln)
    for r in records:
        name = r.get("name", "q").encode() + b"\0"
        cig = b"".join(struct.pack("<I", (ln << 4) | op) for op, ln in r["cigar"])
        l_seq = r["l_seq"]
        body = struct.pack("<iiBBHHHiiii", r["tid"], r["pos"], len(name), r["mapq"], 4680, len(r["cigar"]), r["flag"], l_seq,
                           r["mtid"], r["mpos"], r["tlen"]) + name + cig + b"\x11" * ((l_seq + 1) // 2) + b"\x20" * l_seq
        raw += struct.pack("<i", len(body)) + body
    with open(path, "wb") as fh:
        for o in range(0, len(raw), block_bytes):
            fh.write(_bgzf_block(raw[o:o + block_bytes]))
        fh.write(_bgzf_block(b""))   # EOF marker


def _random_records(rng, n, n_refs):
    ops_pool = [[(0, 100)], [(4, 7), (0, 93)], [(0, 60), (1, 3), (0, 37)], [(5, 4), (4, 5), (0, 80), (2, 6), (0, 10), (4, 1)],
                [(0, 50), (3, 200), (0, 50)], [(4, 100)], [(7, 40), (8, 1), (7, 59)], []]
    recs = []
    for i in range(n):
        cig = ops_pool[int(rng.integers(len(ops_pool)))]
        l_seq = sum(ln for op, ln in cig if op in (0, 1, 4, 7, 8)) if cig else 100
        if rng.random() < 0.05:
            l_seq = 0   # SEQ '*'
        recs.append(dict(tid=int(rng.integers(-1, n_refs)), pos=int(rng.integers(0, 50000)), mapq=int(rng.integers(0, 61)),
                         flag=int(rng.integers(0, 4096)), l_seq=l_seq, mtid=int(rng.integers(-1, n_refs)),
                         mpos=int(rng.integers(0, 50000)), tlen=int(rng.integers(-9000, 9000)), cigar=cig,
                         name="read%d" % i if i % 7 else "r" * 200))
    return recs


def _assert_same(a, b, head=1000):
    assert a.references == b.references and list(a.lengths) == list(b.lengths)
    for f in ("tid", "mtid", "pos", "mpos", "tlen", "qlen", "flag", "mapq"):
        assert np.array_equal(getattr(a, f), getattr(b, f)), f
    assert np.array_equal(a.rlen[:head], b.rlen[:head]) and np.array_equal(a.alen[:head], b.alen[:head])


@pytest.mark.parametrize("n,block_bytes,threads", [(0, 3000, 2), (1, 3000, 1), (5000, 700, 4), (5000, 65000, 3), (20000, 3000, 8)])
def test_native_reader_equals_python_reader(tmp_path, n, block_bytes, threads):
    rng = np.random.default_rng(n + block_bytes)
    refs = [("c%d,pos:%d-%d,rc:0" % (i, i * 1000, i * 1000 + 900), 900 + i) for i in range(37)]
    path = str(tmp_path / "t.bam")
    write_bam(path, refs, _random_records(rng, n, len(refs)), block_bytes)
    py = bamio.read_bam(path)
    nat = bamio.read_bam_native(path, threads=threads)
    assert len(nat) == n
    _assert_same(nat, py)
    assert nat.stats["records"] == n and nat.stats["blocks"] >= 1
    if n:   # the packed column the graph build uploads: flag | mapq << 12 | qlen << 20
        assert nat.packed is not None and np.array_equal(nat.packed, abi_pack(py.flag, py.mapq, py.qlen))
    if n > 100:
        part = bamio.read_bam_native(path, threads=threads, max_records=100)
        assert len(part) == 100 and np.array_equal(part.pos, py.pos[:100])


def test_records_straddling_inflate_windows(tmp_path, monkeypatch):
    """The reader inflates a window of blocks at a time and carries a partial record into the next one."""
    rng = np.random.default_rng(11)
    refs = [("c%d" % i, 5000 + i) for i in range(9)]
    path = str(tmp_path / "w.bam")
    write_bam(path, refs, _random_records(rng, 8000, len(refs)), block_bytes=1500)
    py = bamio.read_bam(path)
    for window in (1, 4000, 50000):   # 1: every block is its own window
        monkeypatch.setenv("BESST_BAMIO_WINDOW", str(window))
        _assert_same(bamio.read_bam_native(path, threads=3), py)


def test_slices_keep_the_native_buffers_alive(tmp_path):
    """The zero-copy column views own a reference to the native handle through ndarray.base, so a slice
    (or a single column) outlives the parent batch."""
    import gc
    rng = np.random.default_rng(3)
    refs = [("c%d" % i, 5000 + i) for i in range(9)]
    path = str(tmp_path / "o.bam")
    write_bam(path, refs, _random_records(rng, 6000, len(refs)), block_bytes=3000)
    want = bamio.read_bam(path)
    import weakref
    batch = bamio.read_bam_native(path, threads=2)
    alive = weakref.ref(batch._owner)
    part = batch.slice(1000, 3000)
    col = batch.pos
    del batch
    gc.collect()
    assert alive() is not None and alive().ptr is not None   # held by the views, not by the batch
    junk = [np.full(6000, -7, np.int32) for _ in range(64)]   # would overwrite the freed heap
    assert np.array_equal(part.pos, want.pos[1000:3000]) and np.array_equal(part.flag, want.flag[1000:3000])
    assert np.array_equal(col, want.pos)
    del part, col, junk
    gc.collect()
    assert alive() is None   # and released with the last view


def test_native_reader_rejects_garbage(tmp_path):
    p = tmp_path / "x.bam"
    p.write_bytes(b"this is not a BAM file, not even gzip" * 10)
    with pytest.raises(IOError):
        bamio.read_bam_native(str(p))
    with pytest.raises(IOError):
        bamio.read_bam_native(str(tmp_path / "missing.bam"))
    # truncated in the middle of a block
    refs = [("c0", 1000)]
    good = tmp_path / "g.bam"
    write_bam(str(good), refs, _random_records(np.random.default_rng(1), 500, 1), 3000)
    data = good.read_bytes()
    (tmp_path / "t.bam").write_bytes(data[:len(data) // 2])
    with pytest.raises(IOError):
        bamio.read_bam_native(str(tmp_path / "t.bam"))


@pytest.mark.skipif(not os.path.exists(REF_BAM), reason="reference testdata not on this machine")
def test_native_reader_on_reference_testset1():
    py = bamio.read_bam(REF_BAM, max_records=300000)
    nat = bamio.read_bam_native(REF_BAM, max_records=300000)
    _assert_same(nat, py)
    full = bamio.read_bam_native(REF_BAM)
    assert len(full) == 1999958 and len(full.references) == 1836
    assert int((full.tid != full.mtid).sum()) == 1356734


def test_dropin_entry_points_accept_a_bam_path(tmp_path):
    """get_metrics / PE called with a file name: the native reader decodes the file once (cached for the second
    call) and the pass equals the one over the in-memory records the file was written from."""
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, os.path.join(os.path.dirname(here), "oracle"))
    import helpers
    from oracle_engine import OracleEngine
    from besst_b200 import records, synth
    batch = synth.make_config("tiny").to_batch()
    recs = []
    for i in range(len(batch)):
        q = int(batch.qlen[i])
        recs.append(dict(tid=int(batch.tid[i]), pos=int(batch.pos[i]), mapq=int(batch.mapq[i]), flag=int(batch.flag[i]), l_seq=q,
                         mtid=int(batch.mtid[i]), mpos=int(batch.mpos[i]), tlen=int(batch.tlen[i]), cigar=[(0, q)] if q else []))
    path = str(tmp_path / "lib.bam")
    write_bam(path, list(zip(batch.references, [int(x) for x in batch.lengths])), recs, block_bytes=30000)
    opts = dict(orientation="fr")
    want = helpers.run_dropin(batch, opts, OracleEngine())
    got = helpers.run_dropin(batch, opts, OracleEngine(), bam_path=path)
    assert got["G"] == want["G"] and got["G_prime"] == want["G_prime"] and got["param"] == want["param"]
    assert got["objects"] == want["objects"]
    assert len(records._open_cache) == 1   # decoded once for get_metrics + PE


def test_streamed_windows_concatenate_to_the_full_read(tmp_path, monkeypatch):
    rng = np.random.default_rng(3)
    refs = [("c%d" % i, 4000 + i) for i in range(5)]
    path = str(tmp_path / "s.bam")
    write_bam(path, refs, _random_records(rng, 6000, len(refs)), block_bytes=2500)
    full = bamio.read_bam(path)
    monkeypatch.setenv("BESST_BAMIO_WINDOW", "30000")
    parts, firsts, seen_refs = [], [], []

    def on_window(cols, first, references, lengths):
        parts.append({k: v.copy() for k, v in cols.items()})   # the views die with the call
        firsts.append(first)
        seen_refs.append((tuple(references), tuple(lengths)))

    stats = bamio.stream_bam_native(path, on_window, threads=3)
    assert len(parts) > 5 and stats["records"] == len(full)
    assert firsts == list(np.cumsum([0] + [len(p["tid"]) for p in parts[:-1]]))
    assert all(r == (tuple(full.references), tuple(full.lengths)) for r in seen_refs)
    for f in ("tid", "mtid", "pos", "mpos", "tlen", "qlen", "flag", "mapq"):
        assert np.array_equal(np.concatenate([p[f] for p in parts]), getattr(full, f)), f
    assert stats["stopped"] is False
    # a callback can stop the pass: not an error, the statistics come back with the flag set
    early = bamio.stream_bam_native(path, lambda *a: True, threads=2)
    assert early["stopped"] is True and 0 < early["records"] < len(full) and early["blocks"] == stats["blocks"]
    with pytest.raises(ZeroDivisionError):
        bamio.stream_bam_native(path, lambda *a: 1 // 0, threads=2)


def test_corrupted_block_is_caught_by_its_crc(tmp_path, monkeypatch):
    """A flipped byte inside a STORED deflate block leaves the stream and ISIZE valid: only the CRC32 of the
    gzip trailer can tell.  The reader checks it (BESST_BAMIO_NOCRC=1 turns the check off)."""
    rng = np.random.default_rng(5)
    refs = [("c%d" % i, 4000 + i) for i in range(5)]
    raw_path = str(tmp_path / "ok.bam")
    write_bam(raw_path, refs, _random_records(rng, 800, len(refs)), block_bytes=4000)
    data = bytearray(open(raw_path, "rb").read())
    # re-write the second block as a stored (uncompressed) deflate block and damage one payload byte
    def blocks(buf):
        o = 0
        while o < len(buf):
            bsize = struct.unpack_from("<H", buf, o + 16)[0] + 1
            yield o, bsize
            o += bsize
    offs = list(blocks(data))
    o, bsize = offs[1]
    payload = zlib.decompress(bytes(data[o + 18:o + bsize - 8]), -15)
    stored = b"\x01" + struct.pack("<HH", len(payload), len(payload) ^ 0xffff) + payload
    damaged = bytearray(stored)
    damaged[5 + len(payload) // 2 + 3] ^= 0x01   # inside a record's name/sequence bytes: lengths stay consistent
    new_block = (b"\x1f\x8b\x08\x04" + b"\0\0\0\0" + b"\0\xff" + struct.pack("<H", 6) + b"BC" + struct.pack("<HH", 2, 12 + 6 + len(damaged) + 8 - 1)
                 + bytes(damaged) + struct.pack("<II", zlib.crc32(payload) & 0xffffffff, len(payload)))
    bad_path = str(tmp_path / "bad.bam")
    open(bad_path, "wb").write(bytes(data[:o]) + new_block + bytes(data[o + bsize:]))
    with pytest.raises(IOError, match="CRC32"):
        bamio.read_bam_native(bad_path, threads=2)
    monkeypatch.setenv("BESST_BAMIO_NOCRC", "1")
    assert len(bamio.read_bam_native(bad_path, threads=2)) == 800   # accepted silently without the check
