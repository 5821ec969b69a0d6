"""ORACLE / TEST INFRASTRUCTURE -- stand-in for `pysam` (absent from this image).

Serves pre-decoded records (a besst_b200.records.RecordBatch) to the
reference's unmodified bytecode with the pysam 0.8.4 attribute names it reads
(CreateGraph.py:111-211,812-833; libmetrics.py:63-84,257-301;
bam_parser.py:22-29).  Never imported by the product path.
"""
import sys
import types


class AlignedRead(object):
    __slots__ = ("rname", "mrnm", "pos", "mpos", "mapq", "qlen", "rlen", "alen", "tlen", "flag")

    def __init__(self, rname, mrnm, pos, mpos, mapq, qlen, rlen, alen, tlen, flag):
        self.rname = rname; self.mrnm = mrnm; self.pos = pos; self.mpos = mpos
        self.mapq = mapq; self.qlen = qlen; self.rlen = rlen; self.alen = alen
        self.tlen = tlen; self.flag = flag

    tid = property(lambda s: s.rname)
    rnext = property(lambda s: s.mrnm)
    is_unmapped = property(lambda s: bool(s.flag & 0x4))
    mate_is_unmapped = property(lambda s: bool(s.flag & 0x8))
    is_reverse = property(lambda s: bool(s.flag & 0x10))
    mate_is_reverse = property(lambda s: bool(s.flag & 0x20))
    is_read1 = property(lambda s: bool(s.flag & 0x40))
    is_read2 = property(lambda s: bool(s.flag & 0x80))
    is_secondary = property(lambda s: bool(s.flag & 0x100))


class Samfile(object):
    """In-memory Samfile over a RecordBatch: references, lengths, fetch,
    reset, iteration (file order, all records)."""

    def __init__(self, batch, mode="rb"):
        self._b = batch
        self.references = tuple(batch.references)
        self.lengths = tuple(int(x) for x in batch.lengths)
        self.nreferences = len(self.references)
        n = len(batch)
        rlen = batch.rlen if batch.rlen is not None else batch.qlen
        alen = batch.alen if batch.alen is not None else batch.qlen
        cols = [batch.tid.tolist(), batch.mtid.tolist(), batch.pos.tolist(), batch.mpos.tolist(),
                batch.mapq.tolist(), batch.qlen.tolist(),
                (rlen.tolist() + batch.qlen[len(rlen):].tolist()) if len(rlen) < n else rlen.tolist(),
                (alen.tolist() + batch.qlen[len(alen):].tolist()) if len(alen) < n else alen.tolist(),
                batch.tlen.tolist(), batch.flag.tolist()]
        self._rows = list(zip(*cols))
        self._it = None
        self.reset()

    def reset(self):
        self._it = iter(self._rows)

    def fetch(self, reference=None, *a, **k):
        if reference is not None and reference not in self.references:
            raise ValueError("invalid reference %r" % (reference,))
        return iter(())

    def getrname(self, tid):
        return self.references[tid]

    def __iter__(self):
        return self

    def __next__(self):
        return AlignedRead(*next(self._it))

    next = __next__

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    def close(self):
        pass


def install():
    mod = types.ModuleType("pysam")
    mod.Samfile = Samfile
    mod.AlignmentFile = Samfile
    mod.AlignedRead = AlignedRead
    sys.modules["pysam"] = mod
    return mod
