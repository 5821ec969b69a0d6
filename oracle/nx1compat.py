"""ORACLE / TEST INFRASTRUCTURE -- networkx 1.x API on top of networkx 3.x.

The reference pins networkx 1.10 (requirements.txt:2) and uses `G.edge`,
`G.node`, list-returning `edges()/nodes()/neighbors()` and removal while
iterating over `G.edges()` (CreateGraph.py:292-296,363-374,399-403,715-716,
842-862).  Iteration follows Python-3 dict insertion order.
"""
import sys
import types

import networkx as _nx


class Graph(_nx.Graph):
    @property
    def edge(self):
        return self._adj

    @property
    def node(self):
        return self._node

    # networkx 3 implements edges/nodes/degree as cached properties that store the view object in the
    # INSTANCE dict, where it would shadow these methods after the first call: build the views directly
    def edges(self, nbunch=None, data=False, default=None):
        return list(_nx.classes.reportviews.EdgeView(self)(nbunch=nbunch, data=data, default=default))

    def edges_iter(self, nbunch=None, data=False, default=None):
        return iter(self.edges(nbunch, data, default))

    def nodes(self, data=False):
        return list(_nx.classes.reportviews.NodeView(self)(data=data))

    def nodes_iter(self, data=False):
        return iter(self.nodes(data))

    def neighbors(self, n):
        return list(self._adj[n])

    def neighbors_iter(self, n):
        return iter(self._adj[n])


def install():
    mod = types.ModuleType("networkx")
    mod.__dict__.update({k: v for k, v in _nx.__dict__.items() if not k.startswith("__")})
    mod.Graph = Graph
    mod._real = _nx
    sys.modules["networkx"] = mod
    return mod
