"""The post-loop filters of `CreateGraph.PE` (CreateGraph.py:237-321, 355-404, 690-708) on the CSR edge
list instead of on networkx graphs.

The engine returns ONE edge list; `(G, G_prime)` are two masks over it plus, per graph, the set of
scaffolds that have nodes.  Every filter of the reference becomes a vector operation on those masks --
except `remove_edges_below_threshold` (:355-374), which is order dependent (an edge is dropped only while
both endpoints still have more than four neighbours, in `G.edges()` iteration order): the order is
reproduced from the node insertion order and the edges' first appearance, the sequential loop runs in C
(`besst_csr_prune_dense`).  networkx objects are built once, at the end, for the surviving edges only.

`G.edges()` order (networkx, Python-3 dict order): nodes in insertion order; for a node, its neighbours in
the order the edges were inserted; an edge is reported at its earlier endpoint.  Link edges are inserted in
order of their first accepted link (first_idx), after all nodes exist, so the edge order is
sorted by (rank of the earlier endpoint, first_idx).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import abi


class ObsList(object):
    """Read-only list-like view of one edge's observations (a slice of the engine's result arrays):
    len / iteration / indexing / comparison like the list the reference stores, converted to Python ints
    only when touched.  Opt-in (`param.lazy_observations`): a 1e8-link library otherwise costs seconds of
    `.tolist()` that most edges never need."""
    __slots__ = ("_a", "_b", "_l")

    def __init__(self, a, b=None):
        self._a, self._b, self._l = a, b, None

    def _list(self):
        if self._l is None:
            v = self._a if self._b is None else self._a.astype(np.int64) + self._b
            self._l = v.tolist()
        return self._l

    def __len__(self):
        return int(self._a.shape[0])

    def __iter__(self):
        return iter(self._list())

    def __getitem__(self, i):
        return self._list()[i]

    def __eq__(self, other):
        return self._list() == (other._list() if isinstance(other, ObsList) else other)

    def __repr__(self):
        return repr(self._list())

    def tolist(self):
        return list(self._list())


class CsrGraphs(object):
    """(G, G_prime) as masks over the engine's edge list."""

    def __init__(self, res, table, param):
        self.res, self.table, self.param = res, table, param
        E, S, nl = res.n_edges, table.n_scaffolds, table.n_large_scaffolds
        self.su = (res.edge_u >> 1).astype(np.int64)
        self.sv = (res.edge_v >> 1).astype(np.int64)
        self.nr = res.nr_links.astype(np.int64)
        large = np.arange(S) < nl
        scoring = not param.no_score
        # which scaffolds get nodes in which graph (InitializeGraph calls, CreateGraph.py:85-96)
        if param.no_score:
            self.node_G, self.node_GP = np.zeros(S, bool), np.ones(S, bool)
        elif param.extend_paths:
            self.node_G, self.node_GP = large.copy(), np.ones(S, bool)
        else:
            self.node_G, self.node_GP = large.copy(), np.zeros(S, bool)
        # which link edges CreateEdge inserts into which graph (:170-206)
        ll = (res.flags & abi.EDGE_LL) != 0
        self.edge_G = ll & scoring
        self.edge_GP = np.full(E, bool(param.no_score or param.extend_paths))
        self.n_large = nl

    # -- views ---------------------------------------------------------------------------------------------
    def alive(self, which):
        node, edge = (self.node_G, self.edge_G) if which == "G" else (self.node_GP, self.edge_GP)
        if edge.shape[0] == 0:
            return edge
        return edge & node[self.su] & node[self.sv]

    def number_of_edges(self, which):
        node = self.node_G if which == "G" else self.node_GP
        return int(self.alive(which).sum()) + int(node.sum())   # link edges + one intra-scaffold edge per scaffold

    def number_of_nodes(self, which):
        return 2 * int((self.node_G if which == "G" else self.node_GP).sum())

    # -- node removal (filter_low_coverage_contigs :407-433, RepeatDetector :959-1018) -------------------------
    def remove_scaffold(self, scaffold_name, large):
        i = self.table.scaffold_index[scaffold_name]
        if large:
            self.node_G[i] = False
            if self.param.extend_paths:
                self.node_GP[i] = False
        else:
            self.node_GP[i] = False

    # -- RemoveBugEdges (:690-708) -----------------------------------------------------------------------------
    def remove_bug_edges(self):
        res = self.res
        bad = (res.fishy > 0) & (res.fishy >= self.nr)
        if self.param.extend_paths:
            hit = bad & self.alive("G_prime")
            removed = int(hit.sum())
            self.edge_G &= ~(bad & self.alive("G"))
            self.edge_GP &= ~hit
        else:
            hit = bad & self.alive("G")
            removed = int(hit.sum())
            self.edge_G &= ~hit
        return removed

    # -- link-count profile of G_prime for infer_spurious_link_count_threshold (:336-346) ------------------------
    def link_count_profile(self):
        """[(link_number, edges with at least that many links)] in descending link_number"""
        vals, counts = np.unique(self.nr[self.alive("G_prime")], return_counts=True)
        cum = np.cumsum(counts[::-1])
        return list(zip(vals[::-1].tolist(), cum.tolist()))

    # -- support filters (:291-296, :355-404) ----------------------------------------------------------------------
    def drop_low_support(self, which, edgesupport):
        low = self.alive(which) & (self.nr < edgesupport)
        if which == "G":
            self.edge_G &= ~low
        else:
            self.edge_GP &= ~low
        return int(low.sum())

    def node_rank_GP(self, node_id):
        """position of a node in G_prime's node order: small scaffolds first, then the large ones (:87-95)"""
        s, side = node_id >> 1, node_id & 1
        n_small = self.table.n_scaffolds - self.n_large
        return np.where(s >= self.n_large, 2 * (s - self.n_large) + side, 2 * (n_small + s) + side)

    def prune_dense_regions(self, limit, min_neighbours=4):
        """remove_edges_below_threshold's first loop on G_prime -> number of removed edges"""
        from ._lib import load
        res = self.res
        alive = self.alive("G_prime")
        weak = np.nonzero(alive & (self.nr < limit))[0]
        if weak.shape[0] == 0:
            return 0
        u, v = res.edge_u.astype(np.int64), res.edge_v.astype(np.int64)
        ru, rv = self.node_rank_GP(u[weak]), self.node_rank_GP(v[weak])
        order = np.lexsort((res.first_idx[weak], np.minimum(ru, rv)))
        weak = weak[order]
        degree = np.zeros(2 * self.table.n_scaffolds, dtype=np.int32)
        degree[0::2] = self.node_GP
        degree[1::2] = self.node_GP   # the intra-scaffold edge
        np.add.at(degree, u[alive], 1)
        np.add.at(degree, v[alive], 1)
        wu = np.ascontiguousarray(res.edge_u[weak], dtype=np.uint32)
        wv = np.ascontiguousarray(res.edge_v[weak], dtype=np.uint32)
        dropped = np.zeros(weak.shape[0], dtype=np.uint8)
        n = load().besst_csr_prune_dense(weak.shape[0], wu.ctypes.data, wv.ctypes.data, degree.ctypes.data,
                                        C.c_int32(min_neighbours), dropped.ctypes.data)
        if n < 0:
            raise RuntimeError("besst_csr_prune_dense failed")
        self.edge_GP[weak[dropped != 0]] = False
        return int(n)

    # -- materialisation: networkx graphs for the surviving edges only ----------------------------------------------
    def materialise(self, new_graph, Scaffolds, small_scaffolds, lazy_observations=False):
        """-> (G, G_prime) exactly as the reference leaves them after PE: node order of InitializeGraph minus the
        removed scaffolds, surviving link edges in first-appearance order with CreateEdge's attributes
        (:842-862) and, on G, GiveScoreOnEdges' gap / score (:541-614)."""
        res, table, param = self.res, self.table, self.param
        G, G_prime = new_graph(), new_graph()
        names = table.scaffold_names
        slen = table.scaffold_lengths.tolist()

        def internals(graph):
            # the node / adjacency dicts behind the Graph object (networkx >= 2: _node/_adj; 1.x: node/adj ARE the dicts).
            # Filling them directly is what add_node / add_edge do, minus ~1 us of argument handling per call --
            # seconds for the 1e5..1e6 nodes and edges of a real assembly.
            return (graph._node, graph._adj) if hasattr(graph, "_adj") else (graph.node, graph.adj)

        def add_nodes(graph, idxs):
            node, adj = internals(graph)
            for i in idxs:
                name, length = names[i], slen[i]
                nl_, nr_ = (name, 'L'), (name, 'R')
                node[nl_] = {'length': length}
                node[nr_] = {'length': length}
                d = {'nr_links': None}
                adj[nl_] = {nr_: d}
                adj[nr_] = {nl_: d}
        nl, S = self.n_large, table.n_scaffolds
        add_nodes(G, np.nonzero(self.node_G[:nl])[0].tolist())
        add_nodes(G_prime, (np.nonzero(self.node_GP[nl:])[0] + nl).tolist())
        add_nodes(G_prime, np.nonzero(self.node_GP[:nl])[0].tolist())

        alive_G, alive_GP = self.alive("G"), self.alive("G_prime")
        order = np.argsort(res.first_idx, kind='stable')
        order = order[(alive_G | alive_GP)[order]]
        if lazy_observations:
            total = lambda b, t: ObsList(res.obs_u[b:t], res.obs_v[b:t])   # noqa: E731
            one = lambda a, b, t: ObsList(a[b:t])                          # noqa: E731
        else:
            tot_all = res.obs_u.astype(np.int64) + res.obs_v
            total = lambda b, t: tot_all[b:t].tolist()                     # noqa: E731
            one = lambda a, b, t: a[b:t].tolist()                          # noqa: E731
        scoring = not param.no_score
        _, adj_G = internals(G)
        _, adj_GP = internals(G_prime)
        side = ('L', 'R')
        cols = zip(res.edge_u[order].tolist(), res.edge_v[order].tolist(), res.row_ptr[order].tolist(), res.row_ptr[order + 1].tolist(),
                   res.nr_links[order].tolist(), res.obs_sum[order].tolist(), res.obs_sq[order].tolist(), alive_G[order].tolist(),
                   alive_GP[order].tolist(), res.flags[order].tolist(), res.gap[order].tolist(), res.score[order].tolist())
        for u, v, b, t, nr, obs, obs_sq, in_G, in_GP, flags, gap, score in cols:
            nu = (names[u >> 1], side[u & 1])
            nv = (names[v >> 1], side[v & 1])
            if in_G:
                d = {'nr_links': nr, 'obs': obs, 'obs_sq': obs_sq, 'observations': total(b, t)}
                if scoring:
                    if flags & abi.EDGE_NEGGAP:   # score skipped: the per-scaffold lists stay (:542-544, 848-849)
                        d[nu[0]] = one(res.obs_u, b, t)
                        d[nv[0]] = one(res.obs_v, b, t)
                        d['gap'] = gap
                        d['score'] = 0
                    else:
                        d['gap'] = gap
                        d['score'] = score if score != 0.0 else 0
                adj_G[nv][nu] = d
                adj_G[nu][nv] = d
            if in_GP:
                d = {'nr_links': nr, 'obs': obs, 'obs_sq': obs_sq, 'observations': total(b, t)}
                adj_GP[nv][nu] = d
                adj_GP[nu][nv] = d
        return G, G_prime
