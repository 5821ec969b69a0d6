"""Record batch: the struct-of-arrays view of a library's BAM records.

This is the data format on the *input* side of the hot path.  The reference
iterates `pysam.AlignedRead` objects one at a time (CreateGraph.py:111,
libmetrics.py:63,257,293); the B200 engine consumes the same per-record fields
as flat, coalescable arrays (one array per field, BAM file order).

Field <-> pysam 0.8.4 attribute <-> BAM fixed-core field:

    tid   rname   refID            int32
    mtid  mrnm    next_refID       int32
    pos   pos     pos (0-based)    int32
    mpos  mpos    next_pos         int32
    tlen  tlen    tlen             int32
    qlen  qlen    l_seq minus leading/trailing soft clips   int32
    flag  (is_reverse=0x10, mate_is_reverse=0x20, is_read1=0x40,
           is_read2=0x80, is_unmapped=0x4, mate_is_unmapped=0x8,
           is_secondary=0x100)                              uint16
    mapq  mapq                                              uint8
    rlen  rlen    l_seq (host only; read-length estimate, libmetrics.py:246-266)
    alen  alen    reference span from CIGAR (host only, same block)
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Optional, Sequence

import numpy as np

FLAG_UNMAPPED = 0x4
FLAG_MATE_UNMAPPED = 0x8
FLAG_REVERSE = 0x10
FLAG_MATE_REVERSE = 0x20
FLAG_READ1 = 0x40
FLAG_READ2 = 0x80
FLAG_SECONDARY = 0x100

_DEVICE_FIELDS = (("tid", np.int32), ("mtid", np.int32), ("pos", np.int32),
                  ("mpos", np.int32), ("tlen", np.int32), ("qlen", np.int32),
                  ("flag", np.uint16), ("mapq", np.uint8))


@dataclass
class RecordBatch:
    """SoA record batch in BAM order plus the header tables."""
    tid: np.ndarray
    mtid: np.ndarray
    pos: np.ndarray
    mpos: np.ndarray
    tlen: np.ndarray
    qlen: np.ndarray
    flag: np.ndarray
    mapq: np.ndarray
    references: Sequence[str] = field(default_factory=list)
    lengths: Sequence[int] = field(default_factory=list)
    rlen: Optional[np.ndarray] = None   # only the first 1000 are ever read
    alen: Optional[np.ndarray] = None
    # optional: flag | mapq << 12 | qlen << 20 in one uint32 column (besst_records.packed): what the graph build uploads and
    # reads instead of the three columns (20 instead of 23 bytes per record).  The native ingest fills it while decoding.
    packed: Optional[np.ndarray] = None

    def __post_init__(self):
        for name, dt in _DEVICE_FIELDS:
            a = np.ascontiguousarray(getattr(self, name), dtype=dt)
            setattr(self, name, a)
        n = self.tid.shape[0]
        for name, _ in _DEVICE_FIELDS:
            if getattr(self, name).shape != (n,):
                raise ValueError("record field %s has shape %s, expected (%d,)"
                                 % (name, getattr(self, name).shape, n))

    def __len__(self):
        return int(self.tid.shape[0])

    def slice(self, lo, hi):
        kw = {name: getattr(self, name)[lo:hi] for name, _ in _DEVICE_FIELDS}
        return RecordBatch(references=self.references, lengths=self.lengths,
                           rlen=None if self.rlen is None else self.rlen[lo:hi],
                           alen=None if self.alen is None else self.alen[lo:hi],
                           packed=None if self.packed is None else self.packed[lo:hi], **kw)

    def with_packed(self):
        """Fill `packed` from flag / mapq / qlen (no-op when it exists or a value does not fit)."""
        if self.packed is None:
            from .abi import pack_record_columns
            self.packed = pack_record_columns(self.flag, self.mapq, self.qlen)
        return self

    def select(self, mask):
        kw = {name: getattr(self, name)[mask] for name, _ in _DEVICE_FIELDS}
        return RecordBatch(references=self.references, lengths=self.lengths,
                           rlen=None if self.rlen is None else self.rlen[mask],
                           alen=None if self.alen is None else self.alen[mask], **kw)

    def device_arrays(self):
        return {name: getattr(self, name) for name, _ in _DEVICE_FIELDS}

    def save(self, path):
        np.savez_compressed(
            path, references=np.array(list(self.references)),
            lengths=np.asarray(self.lengths, dtype=np.int64),
            rlen=np.zeros(0, np.int32) if self.rlen is None else self.rlen[:1000],
            alen=np.zeros(0, np.int32) if self.alen is None else self.alen[:1000],
            **self.device_arrays())

    def save_compact(self, path):
        """Smaller fixture files for coordinate-sorted input: narrow dtypes, positions as differences."""
        np.savez_compressed(
            path, compact=np.array([1]), references=np.array(list(self.references)), lengths=np.asarray(self.lengths, dtype=np.int64),
            rlen=np.zeros(0, np.int32) if self.rlen is None else self.rlen[:1000],
            alen=np.zeros(0, np.int32) if self.alen is None else self.alen[:1000],
            tid=self.tid.astype(np.int32), mtid=self.mtid.astype(np.int32), dpos=np.diff(self.pos, prepend=np.int32(0)).astype(np.int32),
            dmpos=(self.mpos - self.pos).astype(np.int32), tlen=self.tlen, qlen=self.qlen, flag=self.flag, mapq=self.mapq)

    @staticmethod
    def load(path):
        z = np.load(path, allow_pickle=False)
        if "compact" in z.files:
            pos = np.cumsum(z["dpos"].astype(np.int64)).astype(np.int32)
            kw = dict(tid=z["tid"], mtid=z["mtid"], pos=pos, mpos=(pos + z["dmpos"]).astype(np.int32), tlen=z["tlen"], qlen=z["qlen"],
                      flag=z["flag"], mapq=z["mapq"])
            rlen = z["rlen"] if z["rlen"].size else None
            alen = z["alen"] if z["alen"].size else None
            return RecordBatch(references=[str(x) for x in z["references"]], lengths=[int(x) for x in z["lengths"]], rlen=rlen, alen=alen, **kw)
        kw = {name: z[name] for name, _ in _DEVICE_FIELDS}
        rlen = z["rlen"] if z["rlen"].size else None
        alen = z["alen"] if z["alen"].size else None
        return RecordBatch(references=[str(s) for s in z["references"]],
                           lengths=[int(x) for x in z["lengths"]],
                           rlen=rlen, alen=alen, **kw)


def from_alignments(reads, references, lengths):
    """Build a batch from an iterable of pysam-like AlignedRead objects (the
    production ingest path: `for r in pysam.Samfile(...)`)."""
    cols = {k: [] for k in ("tid", "mtid", "pos", "mpos", "tlen", "qlen", "flag", "mapq", "rlen", "alen")}
    for r in reads:
        cols["tid"].append(r.rname)
        cols["mtid"].append(r.mrnm)
        cols["pos"].append(r.pos)
        cols["mpos"].append(r.mpos)
        cols["tlen"].append(r.tlen)
        cols["qlen"].append(r.qlen)
        cols["flag"].append(r.flag)
        cols["mapq"].append(r.mapq)
        cols["rlen"].append(r.rlen)
        cols["alen"].append(r.alen if r.alen is not None else 0)
    return RecordBatch(references=list(references), lengths=[int(x) for x in lengths],
                       rlen=np.asarray(cols.pop("rlen"), np.int32),
                       alen=np.asarray(cols.pop("alen"), np.int32), **cols)


class DeviceRecordBatch(object):
    """Record columns resident in HBM, as the device BAM ingest leaves them (CudaEngine.ingest_bam /
    besst_bam_ingest): the engine's entry points take it wherever they take a RecordBatch (abi.make_records hands
    out `abi_records`, device pointers with on_device = 1).  Host side: the header tables and rlen / alen of the first
    records (all that libmetrics.py:246-266 reads).  The columns belong to the engine and are valid until its next
    ingest_bam / close; to_host() copies them out (tests, the multi-GPU slicing)."""

    def __init__(self, engine, abi_records, references, lengths, rlen, alen, stats):
        self.engine = engine
        self.abi_records = abi_records
        self.references = list(references)
        self.lengths = list(lengths)
        self.rlen, self.alen = rlen, alen
        self.stats = stats

    def __len__(self):
        return int(self.abi_records.n)

    def to_host(self):
        r, n = self.abi_records, len(self)
        cols = {name: self.engine.device_read(getattr(r, name), n, dt) for name, dt in _DEVICE_FIELDS}
        packed = self.engine.device_read(r.packed, n, np.uint32) if r.packed else None
        rlen, alen = np.zeros(n, np.int32), np.zeros(n, np.int32)
        rlen[:self.rlen.shape[0]] = self.rlen
        alen[:self.alen.shape[0]] = self.alen
        return RecordBatch(references=self.references, lengths=self.lengths, rlen=rlen, alen=alen, packed=packed, **cols)


def world():
    """(rank, world size) of the torch.distributed job this process belongs to, (0, 1) outside one."""
    import os
    import sys
    if int(os.environ.get("WORLD_SIZE", "1")) <= 1 and "torch.distributed" not in sys.modules:
        return 0, 1
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return 0, 1
    return dist.get_rank(), dist.get_world_size()


def ingest_mode():
    """How the entry points turn a BAM path into records: 'device' (BGZF inflate + decode on the GPU, columns stay in
    HBM; under a process group every rank ingests ITS part of the file, dist.ingest_bam_distributed) or 'host'
    (libbesst_bamio.so on host threads; under a process group every rank decodes the whole file and slices it).
    BESST_B200_INGEST overrides the default, which is 'device' in a single process and DEFAULT_INGEST_DISTRIBUTED under a
    process group."""
    import os
    mode = os.environ.get("BESST_B200_INGEST")
    if mode is None:
        mode = DEFAULT_INGEST if world()[1] <= 1 else DEFAULT_INGEST_DISTRIBUTED
    if mode not in ("device", "host"):
        raise ValueError("BESST_B200_INGEST must be 'device' or 'host', not %r" % mode)
    return mode


DEFAULT_INGEST = "device"
DEFAULT_INGEST_DISTRIBUTED = "device"


class BatchFile(object):
    """Minimal `pysam.Samfile` look-alike over a decoded RecordBatch: what the
    drop-in entry points read from their `bam_file` argument (`references`,
    `lengths`, `reset`, `fetch`) plus the batch itself, so no per-record Python
    iteration is needed."""

    def __init__(self, batch):
        self.record_batch = batch
        self.references = tuple(batch.references)
        self.lengths = tuple(int(x) for x in batch.lengths)

    def reset(self):
        return None

    def fetch(self, reference=None, *a, **k):
        if reference is not None and reference not in self.references:
            raise ValueError("invalid reference %r" % (reference,))
        return iter(())


_open_cache = {}


def as_file(bam_file, engine=None):
    """What the drop-in entry points do with their `bam_file` argument first: a path becomes a
    BatchFile over the decoded records -- on the device (CudaEngine.ingest_bam) or by the host-thread
    reader (libbesst_bamio.so), see ingest_mode -- cached per path, because runBESST hands the same
    library to get_metrics and then to PE -- anything else passes through."""
    if isinstance(bam_file, (str, bytes)) or hasattr(bam_file, "__fspath__"):
        import os
        key = os.path.abspath(os.fsdecode(bam_file))
        st = os.stat(key)
        stamp = (st.st_size, st.st_mtime_ns)
        hit = _open_cache.get(key)
        if hit is None or hit[0] != stamp:
            _open_cache.clear()   # one library at a time, like the reference's loop (runBESST:143-231)
            if ingest_mode() == "device" and engine is None:
                from .engine import default_engine
                engine = default_engine()   # raises without a GPU: the product path has no CPU fallback
            rank, n_ranks = world()
            if ingest_mode() == "device" and hasattr(engine, "ingest_bam") and n_ranks > 1:
                # every rank its part of the file, already in the BAM-order partition the distributed build takes
                from .dist import host_group, ingest_bam_distributed
                part, info = ingest_bam_distributed(engine, key, rank, n_ranks, group=host_group())
                part.dist_info = dict(info, rank=rank, world=n_ranks, path=key)
                hit = (stamp, BatchFile(part))
            elif ingest_mode() == "device" and hasattr(engine, "ingest_bam"):
                hit = (stamp, BatchFile(engine.ingest_bam(key)))
            else:
                from .bamio import read_bam_native
                hit = (stamp, BatchFile(read_bam_native(key)))
            _open_cache[key] = hit
        return hit[1]
    return bam_file


def as_batch(bam_file):
    """RecordBatch behind a `bam_file` argument: a RecordBatch, a path to a BAM file (decoded by
    libbesst_bamio.so), anything carrying `.record_batch`, or a pysam-like iterable of AlignedRead."""
    if isinstance(bam_file, (RecordBatch, DeviceRecordBatch)) or getattr(bam_file, "dist_info", None):
        return bam_file
    if isinstance(bam_file, (str, bytes)) or hasattr(bam_file, "__fspath__"):   # a path: native threaded ingest
        from .bamio import read_bam_native
        return read_bam_native(bam_file)
    batch = getattr(bam_file, "record_batch", None)
    if batch is not None:
        return batch
    batch = from_alignments(bam_file, bam_file.references, bam_file.lengths)
    if hasattr(bam_file, "reset"):
        bam_file.reset()
    return batch
