"""besst_b200.ExtendLargeScaffolds.BetweenScaffolds (besst_paths_between: the reference's default
heap-driven path search + ScorePaths on a CSR, one search per host thread) against the reference's own
BESST.ExtendLargeScaffolds.BetweenScaffolds, on the G_prime graphs the reference's CreateGraph.PE builds from synthetic
libraries with small contigs between the large ones, on random graphs with equal link counts (heap ties decided by the
path comparison), with the iteration cap hit, with `max_extensions`, with the contamination scoring and with no_score.

CPU only and only where the reference tree exists (its bytecode is the oracle); the searches themselves need no GPU."""
import contextlib
import copy
import io
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "oracle"))
import ref_harness  # noqa: E402

pytestmark = pytest.mark.skipif(not ref_harness.reference_available(), reason="needs the reference tree (its path search is the oracle)")


class _Param(object):
    def __init__(self, **kw):
        self.max_extensions = None
        self.path_threshold = 100000
        self.score_cutoff = 1.5
        self.no_score = False
        self.contamination_ratio = 0
        self.dfs_traversal = True
        self.hit_path_threshold = False
        self.information_file = io.StringIO()
        for k, v in kw.items():
            setattr(self, k, v)


def _end_sets(G_prime, large):
    """the way MakeScaffolds.PROBetweenScaf builds them (:1369-1431): nodes of large scaffolds, isolated ones dropped;
    two identical constructions give two sets with the same pop order"""
    def build():
        keep = [n for n in G_prime.nodes() if large(n)]
        end = set()
        for n in keep:
            nbrs = list(G_prime[n])
            if len(nbrs) == 1 and len(list(G_prime[nbrs[0]])) == 1:
                continue
            end.add(n)
        return end, end.copy()
    return build(), build()


def _compare(G_prime, large, **kw):
    import BESST.ExtendLargeScaffolds as ELS_ref
    from besst_b200 import ExtendLargeScaffolds as ELS
    (end_a, iter_a), (end_b, iter_b) = _end_sets(G_prime, large)
    pa, pb = _Param(**kw), _Param(**kw)
    with contextlib.redirect_stdout(io.StringIO()):
        want = ELS_ref.BetweenScaffolds(G_prime, end_a, iter_a, pa)
        got = ELS.BetweenScaffolds(G_prime, end_b, iter_b, pb, threads=3)
    assert end_a == end_b and iter_a == iter_b
    assert pa.hit_path_threshold == pb.hit_path_threshold
    assert len(got) == len(want)
    for g, w in zip(got, want):
        assert g[1] == w[1] and g[2] == w[2] and g[3] == w[3], (g, w)
        assert g[0] == w[0] and type(g[0]) is type(w[0]), (g[0], w[0])
    assert pa.information_file.getvalue() == pb.information_file.getvalue()
    return want, pa


@pytest.fixture(scope="module")
def reference_graphs():
    """G_prime of the reference's own CreateGraph.PE for a mate-pair library over contigs of which most are smaller
    than the contig threshold: chains of small contigs connect the large ones"""
    from besst_b200 import synth
    ref_harness.load_reference()
    out = {}
    for name, (n_contigs, n_pairs, seed) in {"mp_a": (600, 150000, 21), "mp_b": (1500, 300000, 22)}.items():
        lib = synth.make_library(n_contigs, n_pairs, "rf", 3000.0, 500.0, 0.0, seed=seed)
        r = ref_harness.run_reference(lib.to_batch(), dict(orientation="rf", mean=3000.0, stddev=500.0, readlen=100))
        o = r["objects"]
        out[name] = (o["G_prime"], set(o["Scaffolds"]))
    return out


@pytest.mark.parametrize("name", ["mp_a", "mp_b"])
def test_between_scaffolds_equals_reference_on_pe_graphs(reference_graphs, name):
    G_prime, large_keys = reference_graphs[name]
    want, _ = _compare(G_prime, lambda n: n[0] in large_keys)
    assert len(want) > 10 and max(p[3] for p in want) >= 4      # real multi-contig paths were found
    _compare(G_prime, lambda n: n[0] in large_keys, no_score=True, score_cutoff=0.5)
    _compare(G_prime, lambda n: n[0] in large_keys, contamination_ratio=0.25, score_cutoff=0.3)
    _compare(G_prime, lambda n: n[0] in large_keys, max_extensions=17)


def test_iteration_cap_and_ties(reference_graphs):
    """path_threshold hit in the middle of searches; a graph where every link has the same nr_links, so the heap order is
    decided by (node, path) comparisons"""
    G_prime, large_keys = reference_graphs["mp_b"]
    _, p = _compare(G_prime, lambda n: n[0] in large_keys, path_threshold=7)
    assert p.hit_path_threshold
    import networkx as nx
    rng = np.random.default_rng(5)
    for trial in range(6):
        G = nx.Graph()
        n_scaf = 60
        for s in range(n_scaf):
            G.add_edge((s, "L"), (s, "R"), nr_links=None)
        for _ in range(int(2.2 * n_scaf)):
            a, b = rng.integers(0, n_scaf, 2)
            if a == b:
                continue
            u, v = (int(a), "LR"[int(rng.integers(2))]), (int(b), "LR"[int(rng.integers(2))])
            if not G.has_edge(u, v):
                G.add_edge(u, v, nr_links=5 if trial % 2 == 0 else int(rng.integers(1, 4)))
        large = set(range(0, n_scaf, 4))
        want, _ = _compare(G, lambda n: n[0] in large, score_cutoff=0.0, path_threshold=3000 if trial < 4 else 40)
        assert len(want) > 0


def test_library_entry_point_without_reference_objects():
    """the C entry point on a hand-made graph: two large scaffolds joined through one small contig"""
    import networkx as nx
    from besst_b200 import ExtendLargeScaffolds as ELS
    G = nx.Graph()
    for s in (1, 2, 3):
        G.add_edge((s, "L"), (s, "R"), nr_links=None)
    G.add_edge((1, "R"), (2, "L"), nr_links=7)
    G.add_edge((2, "R"), (3, "L"), nr_links=9)
    end = {(1, "L"), (1, "R"), (3, "L"), (3, "R")}
    iter_nodes = [(1, "R")]   # pop() takes it from the back: one start node

    class OneShot(list):
        pass
    with contextlib.redirect_stdout(io.StringIO()):
        got = ELS.BetweenScaffolds(G, end, OneShot(iter_nodes), _Param(score_cutoff=0.0, max_extensions=1))
    assert got == [[16, 0, [(1, "R"), (2, "L"), (2, "R"), (3, "L")], 4]]
