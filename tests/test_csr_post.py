"""CPU: the CSR-side post-filters (besst_b200/csr_post.py) against the reference's graph-based logic replayed
with real networkx graphs: the order-dependent high-density pruning (CreateGraph.py:355-374) on a random dense
edge list where the order decides the outcome, and the lazy observation lists."""
import numpy as np
import networkx as nx
import pytest

import helpers
from besst_b200 import abi
from besst_b200.csr_post import CsrGraphs, ObsList


class _Table(object):
    def __init__(self, n_scaffolds, n_large):
        self.n_scaffolds, self.n_large_scaffolds = n_scaffolds, n_large
        self.scaffold_names = [100 + i for i in range(n_scaffolds)]
        self.scaffold_index = {n: i for i, n in enumerate(self.scaffold_names)}
        self.scaffold_lengths = np.full(n_scaffolds, 5000, np.int64)


class _Param(object):
    no_score, extend_paths = False, True


def _random_result(rng, n_scaffolds, n_edges):
    pairs = set()
    while len(pairs) < n_edges:
        a, b = rng.integers(0, 2 * n_scaffolds, 2)
        if a >> 1 != b >> 1:
            pairs.add((min(a, b), max(a, b)))
    pairs = sorted(pairs)
    E = len(pairs)
    nr = rng.integers(1, 12, E).astype(np.int32)
    row_ptr = np.concatenate([[0], np.cumsum(nr)]).astype(np.int64)
    L = int(row_ptr[-1])
    first = rng.permutation(L)[:E].astype(np.int64)   # distinct first-appearance ordinals
    return abi.GraphResult(
        edge_u=np.array([p[0] for p in pairs], np.uint32), edge_v=np.array([p[1] for p in pairs], np.uint32), nr_links=nr,
        obs_sum=nr.astype(np.int64) * 700, obs_sq=nr.astype(np.int64) * 490000, first_idx=first, row_ptr=row_ptr,
        gap=np.zeros(E, np.int32), score=np.zeros(E), ks=np.zeros(E), sd_obs=np.zeros(E), sd_model=np.zeros(E),
        fishy=np.zeros(E, np.int32), flags=np.zeros(E, np.uint8), obs_u=np.full(L, 300, np.int32), obs_v=np.full(L, 400, np.int32),
        aligned_len=np.zeros(0, np.int64), counters=np.zeros(abi.N_COUNTERS, np.int64))


def _reference_prune(res, table, limit, edgesupport):
    """the reference's loop on a real networkx graph built the way PE builds G_prime"""
    G = nx.Graph()
    nl = table.n_large_scaffolds
    for i in list(range(nl, table.n_scaffolds)) + list(range(nl)):   # small scaffolds first, then large (:87-95)
        name = table.scaffold_names[i]
        G.add_edge((name, 'L'), (name, 'R'), nr_links=None)
    node = lambda x: (table.scaffold_names[x >> 1], 'R' if x & 1 else 'L')   # noqa: E731
    for e in np.argsort(res.first_idx, kind='stable').tolist():
        G.add_edge(node(int(res.edge_v[e])), node(int(res.edge_u[e])), nr_links=int(res.nr_links[e]))
    weak = [(a, b) for a, b in G.edges() if G[a][b]['nr_links'] is not None and G[a][b]['nr_links'] < limit]
    removed = 0
    for a, b in weak:
        if len(list(G.neighbors(a))) > 4 and len(list(G.neighbors(b))) > 4:
            G.remove_edge(a, b)
            removed += 1
    low = 0
    for a, b in list(G.edges()):
        if G[a][b]['nr_links'] is not None and G[a][b]['nr_links'] < edgesupport:
            G.remove_edge(a, b)
            low += 1
    return G, removed, low


@pytest.mark.parametrize("seed,n_scaffolds,n_edges", [(1, 40, 260), (2, 25, 300), (3, 200, 900)])
def test_high_density_pruning_is_order_exact(seed, n_scaffolds, n_edges):
    rng = np.random.default_rng(seed)
    res = _random_result(rng, n_scaffolds, n_edges)
    table = _Table(n_scaffolds, n_scaffolds // 2)
    want, removed, low = _reference_prune(res, table, limit=8, edgesupport=3)
    graphs = CsrGraphs(res, table, _Param())
    assert graphs.prune_dense_regions(8) == removed and removed > 10
    assert graphs.drop_low_support("G_prime", 3) == low
    _, GP = graphs.materialise(nx.Graph, None, None)
    assert list(GP.nodes()) == list(want.nodes())
    assert [(a, b, d['nr_links']) for a, b, d in GP.edges(data=True)] == [(a, b, d['nr_links']) for a, b, d in want.edges(data=True)]
    # the outcome does depend on the order: processing the weak edges in CSR order instead removes a different set
    alive = np.ones(res.n_edges, bool)
    deg = np.ones(2 * n_scaffolds, np.int64)
    np.add.at(deg, res.edge_u.astype(np.int64), 1)
    np.add.at(deg, res.edge_v.astype(np.int64), 1)
    for e in np.nonzero(res.nr_links < 8)[0].tolist():
        if deg[res.edge_u[e]] > 4 and deg[res.edge_v[e]] > 4:
            alive[e] = False
            deg[res.edge_u[e]] -= 1
            deg[res.edge_v[e]] -= 1
    assert not np.array_equal(alive, graphs.edge_GP | (res.nr_links < 3))


def test_lazy_observation_lists_are_list_equivalent():
    from oracle_engine import OracleEngine
    from besst_b200 import synth
    batch = synth.make_config("small_mp").to_batch()
    opts = dict(orientation="rf", mean=3000.0, stddev=500.0, readlen=100)
    eager = helpers.run_dropin(batch, opts, OracleEngine())
    lazy = helpers.run_dropin(batch, opts, OracleEngine(), param_overrides=dict(lazy_observations=True))
    helpers.assert_signature_equal(lazy["G"], eager["G"], label="lazy/G")
    helpers.assert_signature_equal(lazy["G_prime"], eager["G_prime"], label="lazy/G_prime")
    o = ObsList(np.array([3, 4, 5], np.int32), np.array([10, 20, 30], np.int32))
    assert len(o) == 3 and list(o) == [13, 24, 35] and o[1] == 24 and o == [13, 24, 35] and [a + b for a, b in zip(o, o)] == [26, 48, 70]
    assert all(type(x) is int for x in o)


def test_PE_with_a_lognormal_library_scores_through_the_lognormal_branch():
    """param.lognormal (libmetrics.py:360-390): PE takes the lognormal scoring branch (CreateGraph.py:485-493,523-531)
    instead of raising; gaps of the long-scaffold edges change, the graph structure does not."""
    import math
    from oracle_engine import OracleEngine
    from besst_b200 import synth
    batch = synth.make_config("small_mp").to_batch()
    opts = dict(orientation="rf", mean=3000.0, stddev=500.0, readlen=100)
    opts.update(threshold=6000.0, minsize=5000.0)   # get_metrics would reset param.lognormal (libmetrics.py:235): skip it
    normal = helpers.run_dropin(batch, opts, OracleEngine(), run_libmetrics=False)
    over = dict(lognormal=True, lognormal_sigma=0.17, lognormal_mean=math.log(3000.0) - 0.17 ** 2 / 2,
                empirical_distribution={x: math.exp(-((x - 3000.0) / 500.0) ** 2 / 2) for x in range(200, 6001)})
    logn = helpers.run_dropin(batch, opts, OracleEngine(), param_overrides=over, run_libmetrics=False)
    assert logn["G_prime"] == normal["G_prime"]
    assert [(e["u"], e["v"], e["nr_links"]) for e in logn["G"]["edges"]] == [(e["u"], e["v"], e["nr_links"]) for e in normal["G"]["edges"]]
    gaps_n = [e.get("gap") for e in normal["G"]["edges"] if e["nr_links"] is not None]
    gaps_l = [e.get("gap") for e in logn["G"]["edges"] if e["nr_links"] is not None]
    assert all(g is not None for g in gaps_l) and sum(a != b for a, b in zip(gaps_n, gaps_l)) > 20
    assert all("score" in e for e in logn["G"]["edges"] if e["nr_links"] is not None)
