"""Pure-Python BGZF/BAM reader -> RecordBatch (host tooling, not the hot path).

In production the records come from pysam (`records.from_alignments`); pysam is
not installed in this image, so tests and fixtures decode the reference's
`testdata/*/mapped.bam` with this reader.  It follows the SAM/BAM spec's 32-byte
fixed core; `qlen`/`alen` follow pysam 0.8.4's `query_alignment_length` /
`reference_length`.  BAM decode stays outside the C ABI
(SURVEY.md 8b, row "BAM decode"); a GPU inflate is a "next" row (8f rank 1).
"""
from __future__ import annotations

import gzip
import struct

import numpy as np

from .records import RecordBatch

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
