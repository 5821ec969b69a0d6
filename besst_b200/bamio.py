"""BAM file -> RecordBatch.

`read_bam_native` is the product path: libbesst_bamio.so inflates the BGZF
blocks on a pool of host threads and decodes the fixed-core fields straight into the column layout
the engine consumes (include/besst_bamio.h) -- one pass over the file instead of the reference's
three to four pysam iterations per library (runBESST:162, libmetrics.py:63,257,293,
CreateGraph.py:111).  It raises if the library is not built.

`read_bam` is a pure-Python reader of the same format, kept as test tooling: the native reader is
checked against it (tests/test_bamio.py) and fixtures can be decoded without the native library.
Both follow the SAM/BAM spec's 32-byte fixed core; `qlen`/`alen` follow pysam 0.8.4's
`query_alignment_length` / `reference_length`.
"""
from __future__ import annotations

import gzip
import struct

import numpy as np

import ctypes as C
import os

from .records import RecordBatch

_HERE = os.path.dirname(os.path.abspath(__file__))
BAMIO_SO = os.path.join(_HERE, "libbesst_bamio.so")
_bamio = None


class _Columns(C.Structure):
    _fields_ = [("n", C.c_int64)] + [(k, C.c_void_p) for k in ("tid", "mtid", "pos", "mpos", "tlen", "qlen", "flag", "mapq",
                                                               "rlen", "alen")] + [("n_head", C.c_int64), ("packed", C.c_void_p)]


class BamStats(C.Structure):
    _fields_ = [("compressed_bytes", C.c_int64), ("uncompressed_bytes", C.c_int64), ("blocks", C.c_int64),
                ("records", C.c_int64), ("seconds_inflate", C.c_double), ("seconds_decode", C.c_double),
                ("seconds_total", C.c_double), ("threads", C.c_int32)]


WINDOW_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.POINTER(_Columns), C.c_int64)

BAMIO_EXPORTS = ["besst_bamio_abi_version", "besst_bam_read", "besst_bam_stream", "besst_bam_stopped", "besst_bam_n_refs", "besst_bam_ref_name", "besst_bam_ref_length",
                 "besst_bam_get_columns", "besst_bam_get_stats", "besst_bam_close"]


def load_bamio():
    global _bamio
    if _bamio is not None:
        return _bamio
    if not os.path.exists(BAMIO_SO):
        raise IOError("libbesst_bamio.so is not built (%s): run `python -m besst_b200.build`" % BAMIO_SO)
    L = C.CDLL(BAMIO_SO)
    L.besst_bam_read.restype = C.c_void_p
    L.besst_bam_read.argtypes = [C.c_char_p, C.c_int32, C.c_int64, C.c_int64, C.c_char_p, C.c_int32]
    L.besst_bam_stream.restype = C.c_void_p
    L.besst_bam_stream.argtypes = [C.c_char_p, C.c_int32, C.c_int64, C.c_int64, WINDOW_FN, C.c_void_p, C.c_char_p, C.c_int32]
    L.besst_bam_n_refs.restype = C.c_int64
    L.besst_bam_n_refs.argtypes = [C.c_void_p]
    L.besst_bam_ref_name.restype = C.c_char_p
    L.besst_bam_ref_name.argtypes = [C.c_void_p, C.c_int64]
    L.besst_bam_ref_length.restype = C.c_int64
    L.besst_bam_ref_length.argtypes = [C.c_void_p, C.c_int64]
    L.besst_bam_get_columns.argtypes = [C.c_void_p, C.POINTER(_Columns)]
    L.besst_bam_get_stats.argtypes = [C.c_void_p, C.POINTER(BamStats)]
    L.besst_bam_close.argtypes = [C.c_void_p]
    L.besst_bam_stopped.argtypes = [C.c_void_p]
    _bamio = L
    return L


class _BamHandle(object):
    """Owns the native column buffers; the numpy views of a RecordBatch keep it alive."""

    def __init__(self, lib, ptr):
        self.lib, self.ptr = lib, ptr

    def __del__(self):
        if getattr(self, "ptr", None):
            self.lib.besst_bam_close(self.ptr)
            self.ptr = None


def read_bam_native(path, threads=0, max_records=None, head_records=1000):
    """-> RecordBatch whose columns are zero-copy views of the native buffers (`batch.stats` has the
    inflate/decode timings).  rlen/alen are kept for the first `head_records` records only (the
    reference reads 1000, libmetrics.py:246-266)."""
    L = load_bamio()
    err = C.create_string_buffer(512)
    ptr = L.besst_bam_read(os.fsencode(path), int(threads), -1 if max_records is None else int(max_records), int(head_records),
                           err, len(err))
    if not ptr:
        raise IOError("besst_bam_read: %s" % err.value.decode(errors="replace"))
    h = _BamHandle(L, ptr)
    cols = _Columns()
    L.besst_bam_get_columns(ptr, C.byref(cols))
    n, nh = int(cols.n), int(cols.n_head)

    def view(p, count, dt):
        if count == 0 or not p:
            return np.zeros(0, dtype=dt)
        buf = (C.c_char * (count * np.dtype(dt).itemsize)).from_address(p)
        buf._owner = h   # the ctypes object becomes ndarray.base of every derived view: slices keep the handle alive
        a = np.frombuffer(buf, dtype=dt, count=count)
        return a

    n_ref = int(L.besst_bam_n_refs(ptr))
    references = [L.besst_bam_ref_name(ptr, i).decode("ascii") for i in range(n_ref)]
    lengths = [int(L.besst_bam_ref_length(ptr, i)) for i in range(n_ref)]
    rlen = np.zeros(n, np.int32)
    alen = np.zeros(n, np.int32)
    rlen[:nh] = view(cols.rlen, nh, np.int32)
    alen[:nh] = view(cols.alen, nh, np.int32)
    batch = RecordBatch(tid=view(cols.tid, n, np.int32), mtid=view(cols.mtid, n, np.int32), pos=view(cols.pos, n, np.int32),
                        mpos=view(cols.mpos, n, np.int32), tlen=view(cols.tlen, n, np.int32), qlen=view(cols.qlen, n, np.int32),
                        flag=view(cols.flag, n, np.uint16), mapq=view(cols.mapq, n, np.uint8), references=references,
                        lengths=lengths, rlen=rlen, alen=alen, packed=view(cols.packed, n, np.uint32) if cols.packed else None)
    st = BamStats()
    L.besst_bam_get_stats(ptr, C.byref(st))
    batch.stats = {k: getattr(st, k) for k, _ in BamStats._fields_}
    batch._owner = h
    return batch

_CORE = struct.Struct("<iiBBHHHiiii")   # refID pos l_read_name mapq bin n_cigar flag l_seq next_refID next_pos tlen


def read_bam(path, max_records=None):
    with gzip.open(path, "rb") as fh:
        data = fh.read()
    if data[:4] != b"BAM\x01":
        raise IOError("%s is not a BAM file" % path)
    (l_text,) = struct.unpack_from("<i", data, 4)
    off = 8 + l_text
    (n_ref,) = struct.unpack_from("<i", data, off)
    off += 4
    references, lengths = [], []
    for _ in range(n_ref):
        (l_name,) = struct.unpack_from("<i", data, off)
        off += 4
        references.append(data[off:off + l_name - 1].decode("ascii"))
        off += l_name
        (l_ref,) = struct.unpack_from("<i", data, off)
        off += 4
        lengths.append(l_ref)

    n_total = len(data)
    tid, mtid, pos, mpos, tlen, qlen, flag, mapq, rlen, alen = ([] for _ in range(10))
    unpack_core = _CORE.unpack_from
    count = 0
    while off < n_total:
        (block_size,) = struct.unpack_from("<i", data, off)
        (ref_id, p, l_read_name, mq, _bin, n_cigar, fl, l_seq, next_ref, next_pos, tl) = unpack_core(data, off + 4)
        q_start, q_end, ref_span = 0, l_seq, 0
        if n_cigar:
            coff = off + 36 + l_read_name
            cigar = struct.unpack_from("<%dI" % n_cigar, data, coff)
            for c in cigar:
                op = c & 0xF
                if op in (0, 2, 3, 7, 8):
                    ref_span += c >> 4
            if l_seq == 0:   # pysam infers the query end from the CIGAR when SEQ is '*'
                q_end = sum(c >> 4 for c in cigar if (c & 0xF) in (0, 1, 4, 7, 8))
            for c in cigar:   # leading soft clips (hard clips are skipped)
                op = c & 0xF
                if op == 4:
                    q_start += c >> 4
                elif op != 5:
                    break
            if n_cigar > 1:
                for c in reversed(cigar):  # trailing soft clips
                    op = c & 0xF
                    if op == 4:
                        q_end -= c >> 4
                    elif op != 5:
                        break
        tid.append(ref_id); mtid.append(next_ref); pos.append(p); mpos.append(next_pos)
        tlen.append(tl); qlen.append(q_end - q_start); flag.append(fl); mapq.append(mq)
        rlen.append(l_seq); alen.append(ref_span)
        off += 4 + block_size
        count += 1
        if max_records is not None and count >= max_records:
            break
    return RecordBatch(tid=np.asarray(tid, np.int32), mtid=np.asarray(mtid, np.int32),
                       pos=np.asarray(pos, np.int32), mpos=np.asarray(mpos, np.int32),
                       tlen=np.asarray(tlen, np.int32), qlen=np.asarray(qlen, np.int32),
                       flag=np.asarray(flag, np.uint16), mapq=np.asarray(mapq, np.uint8),
                       references=references, lengths=lengths,
                       rlen=np.asarray(rlen, np.int32), alen=np.asarray(alen, np.int32))


def read_fasta_lengths(path):
    """name -> sequence length, names cut at the first whitespace like
    runBESST:45-74 (ReadInContigseqs) does."""
    out = {}
    name, n = None, 0
    with open(path, "r") as fh:
        for line in fh:
            if line.startswith(">"):
                if name is not None:
                    out[name] = n
                name, n = line[1:].strip().split()[0], 0
            else:
                n += len(line.strip())
    if name is not None:
        out[name] = n
    return out


def stream_bam_native(path, on_window, threads=0, max_records=None, head_records=1000):
    """Decode the file window by window: on_window(columns, first_record, references, lengths) is called for each
    window with a dict of numpy views (tid, mtid, pos, mpos, tlen, qlen, flag, mapq) that are valid only during the
    call -- copy (or upload) what you need.  Returning a true value stops the pass (stats['stopped'] is then True).
    -> stats dict.
    Memory stays bounded for files of any size; this is the hook for overlapping ingest with the sliced upload of
    besst_graph_build."""
    L = load_bamio()
    header = {}
    failure = []

    def view(p, count, dt):
        if count == 0 or not p:
            return np.zeros(0, dtype=dt)
        return np.frombuffer((C.c_char * (count * np.dtype(dt).itemsize)).from_address(p), dtype=dt, count=count)

    def trampoline(_user, handle, cols_p, first):
        try:
            if not header:
                n_ref = int(L.besst_bam_n_refs(handle))
                header["references"] = [L.besst_bam_ref_name(handle, i).decode("ascii") for i in range(n_ref)]
                header["lengths"] = [int(L.besst_bam_ref_length(handle, i)) for i in range(n_ref)]
            c = cols_p.contents
            n = int(c.n)
            cols = {"tid": view(c.tid, n, np.int32), "mtid": view(c.mtid, n, np.int32), "pos": view(c.pos, n, np.int32),
                    "mpos": view(c.mpos, n, np.int32), "tlen": view(c.tlen, n, np.int32), "qlen": view(c.qlen, n, np.int32),
                    "flag": view(c.flag, n, np.uint16), "mapq": view(c.mapq, n, np.uint8)}
            return 1 if on_window(cols, int(first), header["references"], header["lengths"]) else 0
        except BaseException as exc:   # never unwind through the C frames
            failure.append(exc)
            return 1

    cb = WINDOW_FN(trampoline)
    err = C.create_string_buffer(512)
    ptr = L.besst_bam_stream(os.fsencode(path), int(threads), -1 if max_records is None else int(max_records), int(head_records),
                             cb, None, err, len(err))
    if failure:
        if ptr:
            L.besst_bam_close(ptr)
        raise failure[0]
    if not ptr:
        raise IOError("besst_bam_stream: %s" % err.value.decode(errors="replace"))
    st = BamStats()
    L.besst_bam_get_stats(ptr, C.byref(st))
    stopped = bool(L.besst_bam_stopped(ptr))
    L.besst_bam_close(ptr)
    stats = {k: getattr(st, k) for k, _ in BamStats._fields_}
    stats["stopped"] = stopped   # on_window returned a true value before the end of the file
    return stats


_REC_DTYPE = np.dtype([("bs", "<i4"), ("tid", "<i4"), ("pos", "<i4"), ("l_name", "u1"), ("mapq", "u1"), ("bin", "<u2"),
                       ("n_cig", "<u2"), ("flag", "<u2"), ("l_seq", "<i4"), ("mtid", "<i4"), ("mpos", "<i4"), ("tlen", "<i4"),
                       ("name", "S12"), ("cig", "<u4", (2,)), ("seq", "u1", (50,)), ("qual", "u1", (100,))])
BGZF_MAX_INPUT = 0xff00   # htslib's BGZF_BLOCK_SIZE: bytes of input per block


def bgzf_block(data, level=6):
    import zlib
    comp = zlib.compressobj(level, zlib.DEFLATED, -15)
    c = comp.compress(data) + comp.flush()
    bsize = 12 + 6 + len(c) + 8
    hdr = b"\x1f\x8b\x08\x04" + b"\0\0\0\0" + b"\0\xff" + struct.pack("<H", 6) + b"BC" + struct.pack("<HH", 2, bsize - 1)
    return hdr + c + struct.pack("<II", zlib.crc32(data) & 0xffffffff, len(data))


def write_bam_columns(path, batch, level=1, style="htslib", seed=0, chunk_records=1 << 18, threads=0):
    """RecordBatch -> sorted BAM file, vectorised (bench / test tooling: the synthetic libraries as FILES for the ingest
    paths).  Every record is 206 bytes: 100 bases of random sequence, skewed random qualities (a realistic literal / match
    mix: the file compresses about 3x like sequencer output), a two-operation CIGAR whose soft clip carries qlen
    (`(100 - qlen)S qlen M`, or `aM bM` without clipping), read name r<ordinal>.  style 'htslib': no record is split
    across BGZF blocks (bgzf_flush_try); 'packed': blocks are cut every 65280 bytes wherever that falls (htsjdk)."""
    n = len(batch)
    text = b"@HD\tVN:1.0\tSO:coordinate\n"
    head = b"BAM\x01" + struct.pack("<i", len(text)) + text + struct.pack("<i", len(batch.references))
    parts = [head]
    for name, ln in zip(batch.references, batch.lengths):
        nm = name.encode() + b"\0"
        parts.append(struct.pack("<i", len(nm)) + nm + struct.pack("<i", int(ln)))
    head = b"".join(parts)
    rng = np.random.default_rng(seed)
    qual_levels = np.array([2, 11, 25, 37, 37, 37, 37, 37], np.uint8)
    rs = _REC_DTYPE.itemsize
    per_block = BGZF_MAX_INPUT // rs
    from concurrent.futures import ThreadPoolExecutor
    import os as _os
    pool = ThreadPoolExecutor(threads or min(32, _os.cpu_count() or 1))   # zlib releases the GIL: blocks compress in parallel

    def emit(fh, payloads):
        for blk in pool.map(lambda d: bgzf_block(d, level), payloads):
            fh.write(blk)
    with open(path, "wb") as fh:
        for o in range(0, len(head), BGZF_MAX_INPUT):   # htslib flushes after the header: records start a fresh block
            fh.write(bgzf_block(head[o:o + BGZF_MAX_INPUT], level))
        pend = b""
        for r0 in range(0, n, chunk_records):
            r1 = min(n, r0 + chunk_records)
            m = r1 - r0
            rec = np.zeros(m, _REC_DTYPE)
            rec["bs"] = rs - 4
            for f in ("tid", "pos", "mapq", "flag", "mtid", "mpos", "tlen"):
                rec[f] = getattr(batch, f)[r0:r1]
            rec["l_name"], rec["bin"], rec["n_cig"], rec["l_seq"] = 12, 4680, 2, 100
            rec["name"] = np.char.add("r", np.char.zfill(np.arange(r0, r1).astype("U10"), 10)).astype("S12")
            q = np.clip(batch.qlen[r0:r1].astype(np.int64), 0, 100)
            clip = 100 - q
            rec["cig"][:, 0] = np.where(clip > 0, (clip << 4) | 4, ((q // 2) << 4) | 0)
            rec["cig"][:, 1] = np.where(clip > 0, (q << 4) | 0, ((q - q // 2) << 4) | 0)
            rec["seq"] = rng.integers(0, 4, (m, 50), dtype=np.uint8) * 17 % 9 * 16 + (1 << rng.integers(0, 4, (m, 50), dtype=np.uint8))
            rec["qual"] = qual_levels[rng.integers(0, 8, (m, 100), dtype=np.uint8)]
            raw = rec.tobytes()
            if style == "htslib":
                step = per_block * rs
                emit(fh, [raw[o:o + step] for o in range(0, len(raw), step)])
            else:
                raw = pend + raw
                full = len(raw) // BGZF_MAX_INPUT * BGZF_MAX_INPUT
                emit(fh, [raw[o:o + BGZF_MAX_INPUT] for o in range(0, full, BGZF_MAX_INPUT)])
                pend = raw[full:]
        if pend:
            fh.write(bgzf_block(pend, level))
        fh.write(bgzf_block(b"", level))   # EOF marker
    pool.shutdown()
