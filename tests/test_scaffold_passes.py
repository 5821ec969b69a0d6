"""SURVEY.md 8f rank 2, the consumer's own graph passes: besst_b200.MakeScaffolds.RemoveIsolatedContigs /
RemoveAmbiguousRegionsUsingScore (besst_scaffold_prune_ambiguous: the order-dependent loop in C on the edge list) /
RemoveLoops, and besst_b200.ExtendLargeScaffolds.BetweenScaffolds (besst_paths_between), swapped into the reference's
OWN MakeScaffolds.Algorithm (MakeScaffolds.py:49-130, runBESST:199): the run must leave the very scaffolds, the very
G_prime and the very Information text the unmodified reference leaves, for one library and for a PE -> MP sequence; plus
graphs made to hit the ambivalent-score rule, equal scores and cycles.

CPU only and only where the reference tree exists (its bytecode is the oracle)."""
import contextlib
import io
import os
import re
import sys

import numpy as np
import pytest

import helpers
import test_consumer_boundary as tcb

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "oracle"))
import ref_harness  # noqa: E402

pytestmark = pytest.mark.skipif(not ref_harness.reference_available(), reason="needs the reference tree (consumer bytecode)")

_TIMING = re.compile(r"^[^\n]*(elapsed|time)[^\n]*$", re.M | re.I)   # wall-clock lines


@contextlib.contextmanager
def swapped(MS):
    """the reference's MakeScaffolds with its three cleaning passes and its path search replaced by the drop-ins"""
    from besst_b200 import ExtendLargeScaffolds as ELSB, MakeScaffolds as MSB
    saved = (MS.RemoveIsolatedContigs, MS.RemoveAmbiguousRegionsUsingScore, MS.RemoveLoops, MS.ELS)
    MS.RemoveIsolatedContigs, MS.RemoveAmbiguousRegionsUsingScore, MS.RemoveLoops, MS.ELS = (
        MSB.RemoveIsolatedContigs, MSB.RemoveAmbiguousRegionsUsingScore, MSB.RemoveLoops, ELSB)
    try:
        yield
    finally:
        MS.RemoveIsolatedContigs, MS.RemoveAmbiguousRegionsUsingScore, MS.RemoveLoops, MS.ELS = saved


def _info(p):
    return _TIMING.sub("<t>", p.info.getvalue())


@pytest.mark.parametrize("kind,seed", [("mp", 31), ("pe", 32)])
def test_algorithm_with_the_dropin_passes_one_library(kind, seed):
    ref, MS = tcb._consumer()
    batch, opts = tcb._library(kind, seed)
    a = tcb._Pipeline(ref, MS, use_dropin=False)
    want = a.library(batch, opts, 1)
    b = tcb._Pipeline(ref, MS, use_dropin=False)
    with swapped(MS):
        got = b.library(batch, opts, 1)
    tcb._assert_same_outcome(want, got, kind)
    assert _info(a) == _info(b)
    assert "isolated contigs removed" in _info(b) and "cycles removed from graph" in _info(b)


def test_algorithm_with_the_dropin_passes_two_libraries():
    ref, MS = tcb._consumer()
    lib1, opts1 = tcb._library("pe", 41)
    lib2, opts2 = tcb._library("mp", 41)
    a, b = tcb._Pipeline(ref, MS, use_dropin=False), tcb._Pipeline(ref, MS, use_dropin=False)
    w1 = a.library(lib1, opts1, 1)
    with swapped(MS):
        g1 = b.library(lib1, opts1, 1)
    tcb._assert_same_outcome(w1, g1, "lib1")
    w2 = a.library(lib2, opts2, 2)
    with swapped(MS):
        g2 = b.library(lib2, opts2, 2)
    tcb._assert_same_outcome(w2, g2, "lib2")
    assert _info(a) == _info(b)


class _P(object):
    extend_paths = True
    plots = False


def _random_scored_graph(rng, n_scaf, n_links, score_pool):
    import networkx as nx
    G = nx.Graph()
    for s in range(n_scaf):
        G.add_edge((s, "L"), (s, "R"), nr_links=None)
    for _ in range(n_links):
        a, b = rng.integers(0, n_scaf, 2)
        if a == b:
            continue
        u, v = (int(a), "LR"[int(rng.integers(2))]), (int(b), "LR"[int(rng.integers(2))])
        if not G.has_edge(u, v):
            G.add_edge(u, v, nr_links=int(rng.integers(1, 30)), score=float(score_pool[int(rng.integers(len(score_pool)))]))
    return G


@pytest.mark.parametrize("trial", range(8))
def test_passes_on_graphs_made_to_be_hard(trial):
    """dense graphs: many nodes with several scored edges, scores drawn from a small pool (ties everywhere, zero scores,
    pairs within the 0.8 rule), cycles left over for RemoveLoops"""
    ref, MS = tcb._consumer()
    from besst_b200 import MakeScaffolds as MSB
    pool_rng = np.random.default_rng(100 + trial)
    pool = [0.0, 0.0, 0.3, 0.5, 0.5, 0.79, 0.8, 0.81, 0.9, 1.0] if trial % 2 == 0 else list(pool_rng.random(50)) + [0.0] * 10

    def build():
        # the same construction sequence for both sides: Graph.copy() would re-insert the edges in node order and change the
        # adjacency order (and with it the order in which networkx's cycle_basis reports a cycle's nodes)
        G = _random_scored_graph(np.random.default_rng(200 + trial), 300, 700 if trial < 6 else 330, pool)
        Gp = _random_scored_graph(np.random.default_rng(200 + trial), 300, 700 if trial < 6 else 330, pool)
        Gp.remove_edges_from(list(Gp.edges())[5::17])   # G_prime lost some edges in PE's own filtering
        return G, Gp
    (G1, Gp1), (G2, Gp2) = build(), build()
    i1, i2 = io.StringIO(), io.StringIO()

    def run(G, Gp, info, iso, amb, loops):
        G = iso(G, info)
        amb(G, Gp, info, _P(), "G")
        G = iso(G, info)
        loops(G, Gp, {}, {}, info, _P())
        return G, Gp
    out1 = run(G1, Gp1, i1, MS.RemoveIsolatedContigs, MS.RemoveAmbiguousRegionsUsingScore, MS.RemoveLoops)
    out2 = run(G2, Gp2, i2, MSB.RemoveIsolatedContigs, MSB.RemoveAmbiguousRegionsUsingScore, MSB.RemoveLoops)
    for X, Y in zip(out1, out2):
        assert list(X.nodes()) == list(Y.nodes())
        assert [(u, v, d.get("nr_links"), d.get("score")) for u, v, d in X.edges(data=True)] == \
               [(u, v, d.get("nr_links"), d.get("score")) for u, v, d in Y.edges(data=True)]
    assert i1.getvalue() == i2.getvalue()
    if trial % 2 == 0:
        assert "SCORES AMBVIVALENT" in i1.getvalue()


@pytest.mark.parametrize("seed,first", [(2, "pe"), (9, "mp")])
def test_everything_swapped_at_once(seed, first):
    """the drop-in get_metrics + PE (oracle engine) AND the drop-in passes + path search inside the reference's Algorithm,
    two libraries in sequence, against the reference all the way (16 random pipelines of this kind were run offline)"""
    ref, MS = tcb._consumer()
    rng = np.random.default_rng(seed)
    n_contigs, n_pairs = int(rng.integers(60, 500)), int(rng.integers(20000, 150000))
    second = "mp" if first == "pe" else "pe"
    lib1, opts1 = tcb._library(first, 500 + seed, n_contigs=n_contigs, n_pairs=n_pairs)
    lib2, opts2 = tcb._library(second, 500 + seed, n_contigs=n_contigs, n_pairs=n_pairs)
    a, b = tcb._Pipeline(ref, MS, use_dropin=False), tcb._Pipeline(ref, MS, use_dropin=True)
    w1 = a.library(lib1, opts1, 1)
    with swapped(MS):
        g1 = b.library(lib1, opts1, 1)
    tcb._assert_same_outcome(w1, g1, "lib1")
    w2 = a.library(lib2, opts2, 2)
    with swapped(MS):
        g2 = b.library(lib2, opts2, 2)
    tcb._assert_same_outcome(w2, g2, "lib2")
    assert _info(a) == _info(b)
