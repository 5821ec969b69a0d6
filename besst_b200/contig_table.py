"""Flatten BESST's Contig/Scaffold objects into the per-tid table the kernels
gather from (include/besst_b200.h: besst_contig_row).

Replaces the string-keyed dict lookups the reference does per record
(CreateGraph.py:118-130,170-206,819-829).  Scaffolds get a dense index: large
scaffolds first, in the iteration order of the `Scaffolds` dict, then the
small ones in `small_scaffolds` order.  With that numbering node id
`2*index + (side == 'R')` increases in the node insertion order of G
(InitializeGraph, CreateGraph.py:710-722), which is what makes `edge[0]` of
`G.edges()` the lower node id (GiveScoreOnEdges, CreateGraph.py:498,569-579).
"""
from __future__ import annotations

import numpy as np

from .abi import CONTIG_ROW_DTYPE, CTG_LARGE, CTG_SMALL


def largest_reference_mask(lengths, k=1000):
    """libmetrics.py:231-233: indexes of the k longest references; heapq.nlargest
    with a key is a stable descending sort, so ties go to the lower tid."""
    lengths = np.asarray(lengths, dtype=np.int64)
    order = np.argsort(-lengths, kind="stable")[:k]
    mask = np.zeros(lengths.shape[0], dtype=bool)
    mask[order] = True
    return mask


class ContigTable(object):
    def __init__(self, references, lengths, Contigs, small_contigs, Scaffolds, small_scaffolds):
        self.scaffold_names = list(Scaffolds.keys()) + list(small_scaffolds.keys())
        self.n_large_scaffolds = len(Scaffolds)
        self.n_scaffolds = len(self.scaffold_names)
        index = {name: i for i, name in enumerate(self.scaffold_names)}
        self.scaffold_index = index
        self.scaffold_lengths = np.fromiter((s.s_length for d in (Scaffolds, small_scaffolds) for s in d.values()),
                                            dtype=np.int64, count=self.n_scaffolds)
        n = len(references)
        rows = np.zeros(n, dtype=CONTIG_ROW_DTYPE)
        mask = largest_reference_mask(lengths)
        rows["in_largest"] = mask
        # column-wise: one pass per attribute over the contig objects instead of one structured-row assignment per contig
        get_large, get_small = Contigs.get, small_contigs.get
        objs = [get_large(name) for name in references]
        state = np.fromiter((CTG_LARGE if c is not None else 0 for c in objs), dtype=np.int32, count=n)
        if small_contigs:
            for tid, c in enumerate(objs):
                if c is None:
                    c = get_small(references[tid])
                    if c is not None:
                        objs[tid] = c
                        state[tid] = CTG_SMALL
        present = [c for c in objs if c is not None]
        where = np.nonzero(state)[0]
        m = len(present)
        si = np.fromiter((index[c.scaffold] for c in present), dtype=np.int64, count=m)
        rows["state"] = state
        rows["scaffold"][where] = si
        rows["direction"][where] = np.fromiter((1 if c.direction else 0 for c in present), dtype=np.int32, count=m)
        rows["position"][where] = np.fromiter((int(c.position) for c in present), dtype=np.int64, count=m)
        rows["length"][where] = np.fromiter((int(c.length) for c in present), dtype=np.int64, count=m)
        rows["scaf_length"][where] = self.scaffold_lengths[si]
        self.rows = rows
        self.references = list(references)

    def node(self, node_id):
        """node id -> the reference's node tuple (scaffold name, 'L'|'R')."""
        return (self.scaffold_names[node_id >> 1], "R" if node_id & 1 else "L")

    def node_id(self, node):
        return 2 * self.scaffold_index[node[0]] + (1 if node[1] == "R" else 0)


def first_library_rows(lengths, contig_threshold):
    """Vectorised equivalent of InitializeObjects (CreateGraph.py:729-786) +
    ContigTable for a first library in which every reference is present in the
    FASTA: one single-contig scaffold per reference, large iff length >=
    contig_threshold.  Returns (rows, n_scaffolds, n_large_scaffolds)."""
    lengths = np.asarray(lengths, dtype=np.int64)
    large = lengths >= contig_threshold
    small = (~large) & (lengths > 0)
    rows = np.zeros(lengths.shape[0], dtype=CONTIG_ROW_DTYPE)
    n_large = int(large.sum())
    scaffold = np.zeros(lengths.shape[0], dtype=np.int64)
    scaffold[large] = np.arange(n_large)
    scaffold[small] = n_large + np.arange(int(small.sum()))
    rows["state"] = np.where(large, CTG_LARGE, np.where(small, CTG_SMALL, 0))
    rows["scaffold"] = scaffold
    rows["direction"] = 1
    rows["length"] = lengths
    rows["scaf_length"] = lengths
    rows["in_largest"] = largest_reference_mask(lengths)
    return rows, n_large + int(small.sum()), n_large
