"""Drop-ins for the graph-cleaning passes at the head of `BESST.MakeScaffolds.Algorithm` (MakeScaffolds.py:49-130; SURVEY.md
8f rank 2): `RemoveIsolatedContigs` (:134-144), `RemoveAmbiguousRegionsUsingScore` (:206-240 with `remove_edges` :156-204)
and `RemoveLoops` (:247-271).  Same signatures, same mutations of G / G_prime, same lines in `Information`.

`RemoveAmbiguousRegionsUsingScore` is the expensive one -- per scored edge the reference sorts and partitions the
neighbour lists of both endpoints in Python -- and it is order dependent, so its loop runs in C on the edge list
(`besst_scaffold_prune_ambiguous`, include/besst_b200.h) in exactly the reference's order: edges sorted by score,
descending, ties in the order `G.edges()` yields them; the removals are then applied to the graphs in bulk.
`RemoveIsolatedContigs` needs no order (a scaffold whose two ends only see each other is isolated whatever happens
elsewhere) and `RemoveLoops` only has work when a connected component has as many edges as nodes, which is decided
without enumerating cycles.

Use: `from besst_b200 import MakeScaffolds as MSB; MS.RemoveIsolatedContigs = MSB.RemoveIsolatedContigs; ...` (or import
these three names in MakeScaffolds.py instead of defining them); tests/test_scaffold_passes.py runs the reference's
Algorithm both ways and compares scaffolds, graphs and the Information text."""
from __future__ import annotations

import numpy as np

from ._lib import load


def _raw_adj(G):
    """the plain dict-of-dicts behind a networkx Graph: `_adj` in networkx >= 2 (where `adj` is a read-only view whose
    item access costs a Python call per lookup), `adj` itself in 1.x"""
    a = getattr(G, "_adj", None)
    return a if a is not None else G.adj


def _edges(adj):
    """(u, v, data) in the order G.edges(data=True) yields them: every node in insertion order, its neighbours in
    insertion order, an edge at its first endpoint only"""
    seen = set()
    for u, nbrs in adj.items():
        for v, d in nbrs.items():
            if v not in seen:
                yield u, v, d
        seen.add(u)


def _n_edges(adj):
    loops = sum(1 for u, nbrs in adj.items() if u in nbrs)
    return (sum(len(nbrs) for nbrs in adj.values()) + loops) // 2


def RemoveIsolatedContigs(G, Information):
    print('Remove isolated nodes.', file=Information)
    counter = 0
    adj = _raw_adj(G)
    doomed = []
    gone = set()
    for node in list(adj):
        if node in gone:
            continue
        nbrs = adj[node]
        if len(nbrs) == 1:
            nbr = next(iter(nbrs))
            if len(adj[nbr]) == 1:
                counter += 1
                doomed.extend((node, nbr))
                gone.add(node)
                gone.add(nbr)
    G.remove_nodes_from(doomed)
    print(str(counter) + ' isolated contigs removed from graph.', file=Information)
    return G


def RemoveAmbiguousRegionsUsingScore(G, G_prime, Information, param, plot):
    if getattr(param, "plots", False):   # the histograms of the decision scores are collected inside the reference's loop
        from BESST import MakeScaffolds as _reference
        return _reference.RemoveAmbiguousRegionsUsingScore(G, G_prime, Information, param, plot)
    nr_edges_before = _n_edges(_raw_adj(G))
    print('Remove edges from node if more than two edges', file=Information)
    us, vs, scores = [], [], []
    for u, v, d in _edges(_raw_adj(G)):
        if 'score' in d:
            if d['nr_links'] is None or not d['nr_links'] > 0:
                continue   # remove_edges only ever looks at neighbours with links (:160)
            us.append(u)
            vs.append(v)
            scores.append(d['score'])
    n = len(us)
    if n:
        nodes = sorted(set(us) | set(vs))
        rank = {x: i for i, x in enumerate(nodes)}
        eu = np.fromiter((rank[x] for x in us), dtype=np.int32, count=n)
        ev = np.fromiter((rank[x] for x in vs), dtype=np.int32, count=n)
        sc = np.asarray(scores, dtype=np.float64)
        order = np.asarray(sorted(range(n), key=scores.__getitem__, reverse=True), dtype=np.int64)
        removed = np.zeros(n, dtype=np.uint8)
        best = np.zeros(2 * n, dtype=np.int64)
        second = np.zeros(2 * n, dtype=np.int64)
        n_amb = load().besst_scaffold_prune_ambiguous(len(nodes), n, eu.ctypes.data, ev.ctypes.data, sc.ctypes.data, order.ctypes.data,
                                                      removed.ctypes.data, best.ctypes.data, second.ctypes.data)
        if n_amb < 0:
            raise ValueError("besst_scaffold_prune_ambiguous rejected its arguments")
        for k in range(n_amb):
            print('SCORES AMBVIVALENT', scores[best[k]], scores[second[k]], file=Information)
        drop = [(us[i], vs[i]) for i in np.nonzero(removed)[0].tolist()]
        G.remove_edges_from(drop)
        if param.extend_paths:
            G_prime.remove_edges_from(drop)   # edges PE's own filtering already took out of G_prime are skipped silently
    nr_edges_after = _n_edges(_raw_adj(G))
    print(' Number of edges in G before:', nr_edges_before, file=Information)
    print(' Number of edges in G after:', nr_edges_after, file=Information)
    try:
        print(' %-age removed edges:', 100 * (1 - (nr_edges_after / float(nr_edges_before))), file=Information)
    except ZeroDivisionError:
        pass
    return ()


def RemoveLoops(G, G_prime, Scaffolds, Contigs, Information, param):
    print('Contigs/scaffolds left:', G.number_of_nodes() / 2, file=Information)
    print('Remove remaining cycles...', file=Information)
    counter = 0
    # a cycle exists iff some connected component has at least as many edges as nodes: decided on arrays, the cycles are
    # enumerated (by networkx, like the reference) only when there are any
    nodes = list(G)
    if nodes:
        import scipy.sparse as sp
        from scipy.sparse.csgraph import connected_components
        rank = {x: i for i, x in enumerate(nodes)}
        pairs = [(rank[u], rank[v]) for u, v, _ in _edges(_raw_adj(G))]
        eu = np.fromiter((a for a, _ in pairs), dtype=np.int64, count=len(pairs))
        ev = np.fromiter((b for _, b in pairs), dtype=np.int64, count=len(pairs))
        m = sp.coo_matrix((np.ones(eu.shape[0], dtype=np.int8), (eu, ev)), shape=(len(nodes), len(nodes)))
        n_comp, label = connected_components(m, directed=False)
        cyclic = bool(eu.shape[0] - len(nodes) + n_comp > 0)
    else:
        cyclic = False
    if cyclic:
        from networkx import algorithms
        for cycle in algorithms.cycles.cycle_basis(G):
            print('A cycle in the scaffold graph: ' + str(cycle) + '\n', file=Information)
            print('A cycle in the scaffold graph: ' + str(cycle), file=Information)
            counter += 1
            for node in cycle:
                if node in G:
                    scaffold_ = node[0]
                    G.remove_nodes_from([(scaffold_, 'L'), (scaffold_, 'R')])
                    if param.extend_paths:
                        G_prime.remove_nodes_from([(scaffold_, 'L'), (scaffold_, 'R')])
    print(str(counter) + ' cycles removed from graph.', file=Information)
    return (G, Contigs, Scaffolds)
