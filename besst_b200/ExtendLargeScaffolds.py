"""Drop-in for the path search between scaffolds, `BESST.ExtendLargeScaffolds.BetweenScaffolds`
(ExtendLargeScaffolds.py:665-712, called from MakeScaffolds.PROBetweenScaf :1367-1450).

Same signature, same side effects on `end` / `iter_nodes` / `param.hit_path_threshold`, same progress lines, same return
value: the list of `[score, bad_link_weight, path, len(path)]` sorted by score.  The search itself -- the reference's
default traversal `find_all_paths_for_start_node_DFS_dynamic_programming_ish` (:526-663) for every start node and
`ScorePaths` (:28-133) for the paths it finds -- runs in `besst_paths_between` (include/besst_b200.h) over a CSR rendering
of G_prime, one search per host thread: the start nodes are independent once `already_visited` / `end` are expressed
through a node's position in the start order, and that order is taken from `iter_nodes.pop()` exactly as the reference's
loop would take it, so the result is the reference's in the same process.  The breadth-first variant (`--bfs_traversal`)
is not restated: its outcome depends on the iteration order of Python sets of tuples."""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import load


def graph_csr(G_prime):
    """networkx G_prime -> (nodes in id order, adj_ptr, adj_node, adj_links).  id = 2 * rank(scaffold key) + (side == 'R')
    with ranks ascending in the key: integer order == tuple order; the contig edge carries -1 (its nr_links is None)."""
    adj = getattr(G_prime, "_adj", None)   # the plain dict-of-dicts (networkx >= 2); `adj` itself in 1.x
    if adj is None:
        adj = G_prime.adj
    keys = sorted({n[0] for n in adj})
    rank = {k: i for i, k in enumerate(keys)}
    n_nodes = 2 * len(keys)
    counts = np.zeros(n_nodes + 1, dtype=np.int64)
    ids, links = [], []
    empty = {}
    for i, k in enumerate(keys):
        for bit, side in enumerate(("L", "R")):
            nbrs = adj.get((k, side), empty)
            counts[2 * i + bit + 1] = len(nbrs)
            for nb, data in nbrs.items():
                ids.append(2 * rank[nb[0]] + (nb[1] == "R"))
                links.append(data.get("nr_links"))
    ptr = np.cumsum(counts)
    nodes = [(k, s) for k in keys for s in ("L", "R")]
    links = np.asarray([-1 if x is None else x for x in links], dtype=np.int64).astype(np.int32)
    return nodes, rank, ptr, np.asarray(ids, dtype=np.int32), links


def BetweenScaffolds(G_prime, end, iter_nodes, param, threads=0):
    print('Entering "find_all_paths_for_start_node" ')
    iter_threshold = param.max_extensions if param.max_extensions else len(end)
    param.hit_path_threshold = False
    n_nodes_g, n_edges_g = G_prime.number_of_nodes(), G_prime.number_of_edges()
    print('iterating until maximum of {0} extensions.'.format(iter_threshold))
    print('Number of nodes:{0}, Number of edges: {1}'.format(n_nodes_g, n_edges_g))
    print('iterating until maximum of {0} extensions.'.format(iter_threshold), file=param.information_file)
    print('Number of nodes:{0}, Number of edges: {1}'.format(n_nodes_g, n_edges_g), file=param.information_file)
    if getattr(param, "dfs_traversal", True) is False:
        raise NotImplementedError("besst_b200.ExtendLargeScaffolds: only the default (heap-driven) traversal is restated; "
                                  "--bfs_traversal depends on the iteration order of Python sets")
    # the start order: the very pops the reference's loop makes (:679-681), `end` loses every start node (:684)
    initial_end = set(end)
    order, iter_count = [], 0
    while len(iter_nodes) > 0 and iter_count <= iter_threshold:
        iter_count += 1
        order.append(iter_nodes.pop())
    for cnter in range(0, len(order), 100):
        print('enter Between scaf node:{0}, scaffold progression {1}%. '.format(cnter, round(cnter / float(iter_threshold) * 100, 1)))
    end.difference_update(order)

    nodes, rank, ptr, adj_node, adj_links = graph_csr(G_prime)

    def nid(n):
        return 2 * rank[n[0]] + (n[1] == "R")
    is_end = np.zeros(len(nodes), dtype=np.uint8)
    for n in initial_end:
        if n[0] in rank:
            is_end[nid(n)] = 1
    order_ids = np.asarray([nid(n) for n in order], dtype=np.int32)
    L = load()
    contamination = 1 if getattr(param, "contamination_ratio", None) else 0
    h = L.besst_paths_between(len(nodes), ptr.ctypes.data, adj_node.ctypes.data if adj_node.size else None,
                              adj_links.ctypes.data if adj_links.size else None, is_end.ctypes.data if is_end.size else None,
                              order_ids.ctypes.data if order_ids.size else None, int(order_ids.shape[0]), int(param.path_threshold),
                              float(param.score_cutoff), 1 if param.no_score else 0, contamination, int(threads))
    if not h:
        raise ValueError("besst_paths_between rejected its arguments")
    try:
        n = int(L.besst_paths_count(h))
        pp, pn, pg, pb, ps = (C.c_void_p() for _ in range(5))
        L.besst_paths_arrays(h, C.byref(pp), C.byref(pn), C.byref(pg), C.byref(pb), C.byref(ps))

        def view(p, count, dt):
            if count == 0 or not p.value:
                return np.zeros(0, dtype=dt)
            return np.frombuffer((C.c_char * (count * np.dtype(dt).itemsize)).from_address(p.value), dtype=dt, count=count).copy()
        path_ptr = view(pp, n + 1, np.int64)
        flat = view(pn, int(path_ptr[-1]) if n else 0, np.int32)
        good, bad = view(pg, n, np.int64), view(pb, n, np.int64)
        if L.besst_paths_hit_threshold(h):
            param.hit_path_threshold = True
    finally:
        L.besst_paths_free(h)
    all_paths = []
    for i in range(n):
        path = [nodes[j] for j in flat[path_ptr[i]:path_ptr[i + 1]].tolist()]
        g, b = int(good[i]), int(bad[i])
        if contamination:
            g = g / 2   # true division, like the reference under Python 3 (:96)
        score = g / float(b) if b != 0 else g
        all_paths.append([score, b, path, len(path)])
    print('Total nr of paths found: {0} with score larger than: {1}'.format(len(all_paths), param.score_cutoff))
    all_paths.sort(key=lambda list_: list_[0])
    if param.hit_path_threshold:
        msg = ('Hit path_threshold of {0} iterations! consider increase --iter <int> parameter to over {0} if speed of BESST is not a problem. '
               'Standard increase is, e.g., 2-10x of current value'.format(param.path_threshold))
        print(msg)
        print(msg, file=param.information_file)
    return all_paths


def WithinScaffolds(G, G_prime, start, end_node, already_visited, max_path_length, param):
    """The per-pair search inside a new scaffold (ExtendLargeScaffolds.py:714-735) is not restated -- one call per pair would
    re-extract the CSR of a G_prime that MakeScaffolds mutates between calls; it is passed through to the reference's, so that
    `from besst_b200 import ExtendLargeScaffolds as ELS` replaces the reference's import as a whole."""
    from BESST import ExtendLargeScaffolds as _reference
    return _reference.WithinScaffolds(G, G_prime, start, end_node, already_visited, max_path_length, param)
