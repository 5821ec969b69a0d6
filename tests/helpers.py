"""Shared test helpers: library construction, contig tables with multi-contig
scaffolds, and field-by-field comparison of two GraphResults."""
from __future__ import annotations

import numpy as np

from besst_b200 import abi
from besst_b200.contig_table import ContigTable
from besst_b200.objects import contig, scaffold

INT_FIELDS = ["edge_u", "edge_v", "nr_links", "obs_sum", "obs_sq", "first_idx", "row_ptr", "fishy", "flags",
              "obs_u", "obs_v", "aligned_len"]
FLOAT_FIELDS = ["score", "ks", "sd_obs", "sd_model"]
FLOAT_RTOL = 1e-6   # north-star tolerance for gap/score floats; integers are bit-exact


def first_library_objects(references, lengths, contig_threshold):
    """What InitializeObjects (CreateGraph.py:729-786) leaves behind."""
    Contigs, Scaffolds, small_contigs, small_scaffolds = {}, {}, {}, {}
    idx = 1
    for name, n in zip(references, lengths):
        c = contig(name, contig_direction=True, contig_position=0, contig_length=int(n))
        s = scaffold(idx, [c], int(n))
        c.scaffold = idx
        if n >= contig_threshold:
            Contigs[name], Scaffolds[idx] = c, s
        elif n > 0:
            small_contigs[name], small_scaffolds[idx] = c, s
        idx += 1
    return Contigs, Scaffolds, small_contigs, small_scaffolds


def later_library_objects(references, lengths, contig_threshold, seed, join=3, drop_every=17):
    """A state as MS.Algorithm would leave it for a later library: runs of
    consecutive contigs joined into multi-contig scaffolds with random
    directions and gaps, a few contigs removed (repeats)."""
    rng = np.random.default_rng(seed)
    Contigs, Scaffolds, small_contigs, small_scaffolds = {}, {}, {}, {}
    idx, i, n = 1, 0, len(references)
    while i < n:
        k = int(rng.integers(1, join + 1))
        members = []
        pos = 0
        for j in range(i, min(n, i + k)):
            if drop_every and j % drop_every == drop_every - 1:
                continue
            c = contig(references[j], contig_scaffold=idx, contig_direction=bool(rng.integers(0, 2)),
                       contig_position=pos, contig_length=int(lengths[j]))
            pos += int(lengths[j]) + int(rng.integers(1, 400))
            members.append(c)
        i += k
        if not members:
            continue
        s_len = members[-1].position + members[-1].length
        s = scaffold(idx, members, s_len)
        big = s_len >= contig_threshold
        for c in members:
            (Contigs if big else small_contigs)[c.name] = c
        (Scaffolds if big else small_scaffolds)[idx] = s
        idx += 1
    return Contigs, Scaffolds, small_contigs, small_scaffolds


def table_for(batch, objects):
    Contigs, Scaffolds, small_contigs, small_scaffolds = objects
    return ContigTable(batch.references, batch.lengths, Contigs, small_contigs, Scaffolds, small_scaffolds)


def assert_graph_equal(got, want, check_scores=True, label=""):
    assert got.n_edges == want.n_edges, "%s: edge count %d != %d" % (label, got.n_edges, want.n_edges)
    assert got.n_links == want.n_links, "%s: link count %d != %d" % (label, got.n_links, want.n_links)
    for f in INT_FIELDS:
        a, b = getattr(got, f), getattr(want, f)
        if f == "flags" and not check_scores:
            a, b = a & abi.EDGE_LL, b & abi.EDGE_LL
        assert np.array_equal(a, b), "%s: integer field %s differs at %s" % (label, f, np.nonzero(a != b)[0][:5])
    n_cnt = 10
    assert np.array_equal(got.counters[:n_cnt], want.counters[:n_cnt]), "%s: counters %s != %s" % (
        label, got.counters[:n_cnt], want.counters[:n_cnt])
    if check_scores:
        scored = (want.flags & abi.EDGE_SCORED) != 0
        assert np.array_equal(got.gap[scored], want.gap[scored]), "%s: gap differs" % label
        for f in FLOAT_FIELDS:
            a, b = getattr(got, f)[scored], getattr(want, f)[scored]
            assert np.array_equal(np.isnan(a), np.isnan(b)), "%s: NaN pattern of %s differs" % (label, f)
            ok = ~np.isnan(b)
            np.testing.assert_allclose(a[ok], b[ok], rtol=FLOAT_RTOL, atol=0, err_msg="%s: %s" % (label, f))


def max_rel_diff(a, b):
    ok = ~np.isnan(b)
    if not ok.any():
        return 0.0
    d = np.abs(a[ok] - b[ok]) / np.maximum(np.abs(b[ok]), 1e-300)
    return float(d.max())
