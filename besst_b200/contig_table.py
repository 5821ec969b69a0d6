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
        self.scaffold_lengths = np.zeros(self.n_scaffolds, dtype=np.int64)
        for name, s in Scaffolds.items():
            self.scaffold_lengths[index[name]] = s.s_length
        for name, s in small_scaffolds.items():
            self.scaffold_lengths[index[name]] = s.s_length
        rows = np.zeros(len(references), dtype=CONTIG_ROW_DTYPE)
        mask = largest_reference_mask(lengths)
        rows["in_largest"] = mask
        for tid, name in enumerate(references):
            c = Contigs.get(name)
            if c is not None:
                state = CTG_LARGE
            else:
                c = small_contigs.get(name)
                if c is None:
                    continue
                state = CTG_SMALL
            si = index[c.scaffold]
            rows[tid] = (state, si, 1 if c.direction else 0, int(c.position), int(c.length),
                         int(self.scaffold_lengths[si]), int(mask[tid]), 0)
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
