// besst_paths.cu -- ELS.BetweenScaffolds (ExtendLargeScaffolds.py:665-712) over a CSR rendering of G_prime: the path search
// of every start node (paths_core.cuh: the reference's default heap-driven traversal, :526-663) and the scoring of the
// found paths (ScorePaths, :28-133), one search per host thread.  SURVEY.md 8f rank 4: "irregular, hard on GPU" -- the
// searches are independent but each is a serial best-first walk with an unbounded frontier, so this rank runs on the
// host cores next to the GPU (like besst_csr_prune_dense); the core is host/device source for a later device port.
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <thread>
#include <vector>

#include "../../include/besst_b200.h"
#include "paths_core.cuh"

struct besst_paths {
    std::vector<int64_t> path_ptr;     // [n_paths + 1]
    std::vector<int32_t> nodes;
    std::vector<int64_t> good, bad;    // link weights of ScorePaths (the caller forms the score with the reference's arithmetic)
    std::vector<int32_t> start_index;  // position of the path's start node in the start order
    int32_t hit_threshold = 0;
    int64_t searches = 0, pops = 0;
};

namespace {

struct Buffers {
    std::vector<int32_t> arena_node, arena_parent, map_key, map_a, map_b, set_key, found;
    std::vector<paths::Entry> heap;
    void size(int scale) {
        arena_node.resize((size_t)1024 << scale); arena_parent.resize(arena_node.size());
        heap.resize((size_t)4096 << scale);
        map_key.resize((size_t)4096 << scale); map_a.resize(map_key.size()); map_b.resize(map_key.size());
        set_key.resize((size_t)8192 << scale);
        found.resize((size_t)512 << scale);
    }
    paths::Scratch fresh() {
        std::fill(map_key.begin(), map_key.end(), -1);
        std::fill(set_key.begin(), set_key.end(), -1);
        paths::Scratch S;
        S.arena_node = arena_node.data(); S.arena_parent = arena_parent.data(); S.arena_cap = (int32_t)arena_node.size(); S.arena_n = 0;
        S.heap = heap.data(); S.heap_cap = (int32_t)heap.size(); S.heap_n = 0;
        S.map_key = map_key.data(); S.map_a = map_a.data(); S.map_b = map_b.data(); S.map_cap = (int32_t)map_key.size(); S.map_n = 0;
        S.set_key = set_key.data(); S.set_cap = (int32_t)set_key.size(); S.set_n = 0;
        S.found_tail = found.data(); S.found_cap = (int32_t)found.size(); S.found_n = 0;
        return S;
    }
};

struct PerStart {
    std::vector<int32_t> nodes;
    std::vector<int32_t> lens;
    std::vector<int64_t> good, bad;
};

}  // namespace

extern "C" besst_paths* besst_paths_between(int64_t n_nodes, const int64_t* adj_ptr, const int32_t* adj_node, const int32_t* adj_links,
                                            const uint8_t* is_end, const int32_t* order, int64_t n_order, int64_t path_threshold,
                                            double score_cutoff, int32_t no_score, int32_t contamination, int32_t n_threads) {
    if (n_nodes < 0 || n_order < 0 || !adj_ptr || (n_order > 0 && !order) || (n_nodes > 0 && !is_end)) return nullptr;
    std::vector<int32_t> order_pos((size_t)n_nodes, INT32_MAX);
    for (int64_t i = 0; i < n_order; ++i) {
        if (order[i] < 0 || order[i] >= n_nodes) return nullptr;
        order_pos[(size_t)order[i]] = (int32_t)i;
    }
    paths::Graph G;
    G.n_nodes = n_nodes; G.adj_ptr = adj_ptr; G.adj_node = adj_node; G.adj_links = adj_links; G.order_pos = order_pos.data(); G.is_end = is_end;
    if (n_threads <= 0) n_threads = (int32_t)std::thread::hardware_concurrency();
    if (n_threads <= 0) n_threads = 1;
    if (n_threads > 256) n_threads = 256;
    if ((int64_t)n_threads > n_order) n_threads = (int32_t)std::max<int64_t>(1, n_order);
    std::vector<PerStart> per((size_t)n_order);
    std::atomic<int64_t> ticket(0), pops(0);
    std::atomic<int> hit(0);
    auto worker = [&]() {
        Buffers B;
        int scale = 0;
        B.size(scale);
        int32_t path[paths::MAX_PATH];
        for (;;) {
            const int64_t i = ticket.fetch_add(1);
            if (i >= n_order) break;
            for (;;) {
                paths::Scratch S = B.fresh();
                int h = 0;
                const int rc = paths::search(G, order[i], (int32_t)i, path_threshold, S, &h);
                if (rc == paths::ST_OVERFLOW && scale < 12) {   // more room and again: the search is deterministic
                    B.size(++scale);
                    continue;
                }
                if (h) hit.store(1);
                pops.fetch_add(S.arena_n);
                PerStart& out = per[(size_t)i];
                for (int32_t f = 0; f < S.found_n; ++f) {
                    int32_t len = 0;
                    for (int32_t c = S.found_tail[f]; c >= 0; c = S.arena_parent[c]) ++len;
                    paths::materialise(S, S.found_tail[f], len, path);
                    int64_t good = 0, bad = 0;
                    paths::connectivity(G, path, len, contamination != 0, &good, &bad);
                    // ScorePaths' filter (:125-131): score >= score_cutoff, and more than the two end nodes unless no_score
                    const double g = contamination ? (double)good / 2.0 : (double)good;
                    const double score = bad != 0 ? g / (double)bad : g;
                    if (!(score >= score_cutoff) || !(no_score || len > 2)) continue;
                    out.nodes.insert(out.nodes.end(), path, path + len);
                    out.lens.push_back(len);
                    out.good.push_back(good);
                    out.bad.push_back(bad);
                }
                break;
            }
        }
    };
    {
        std::vector<std::thread> th;
        for (int t = 1; t < n_threads; ++t) th.emplace_back(worker);
        worker();
        for (auto& x : th) x.join();
    }
    besst_paths* P = new besst_paths();
    P->hit_threshold = hit.load();
    P->searches = n_order;
    P->pops = pops.load();
    P->path_ptr.push_back(0);
    for (int64_t i = 0; i < n_order; ++i) {
        const PerStart& s = per[(size_t)i];
        size_t o = 0;
        for (size_t k = 0; k < s.lens.size(); ++k) {
            P->nodes.insert(P->nodes.end(), s.nodes.begin() + (long)o, s.nodes.begin() + (long)(o + (size_t)s.lens[k]));
            o += (size_t)s.lens[k];
            P->path_ptr.push_back((int64_t)P->nodes.size());
            P->good.push_back(s.good[k]);
            P->bad.push_back(s.bad[k]);
            P->start_index.push_back((int32_t)i);
        }
    }
    return P;
}

extern "C" int64_t besst_paths_count(const besst_paths* p) { return p ? (int64_t)p->good.size() : -1; }
extern "C" int32_t besst_paths_hit_threshold(const besst_paths* p) { return p ? p->hit_threshold : 0; }
extern "C" int64_t besst_paths_pops(const besst_paths* p) { return p ? p->pops : -1; }
extern "C" int besst_paths_arrays(const besst_paths* p, const int64_t** path_ptr, const int32_t** nodes, const int64_t** good,
                                  const int64_t** bad, const int32_t** start_index) {
    if (!p) return BESST_E_INVALID;
    if (path_ptr) *path_ptr = p->path_ptr.data();
    if (nodes) *nodes = p->nodes.data();
    if (good) *good = p->good.data();
    if (bad) *bad = p->bad.data();
    if (start_index) *start_index = p->start_index.data();
    return BESST_OK;
}
extern "C" void besst_paths_free(besst_paths* p) { delete p; }

// ---- RemoveAmbiguousRegionsUsingScore (MakeScaffolds.py:206-240) with its per-node rule remove_edges (:156-204) on the list
// of scored link edges of G.  The pass is sequential and order dependent (what survives at a node depends on what its
// neighbours removed before): the edges arrive in the order G.edges() yields them, `order` is that list sorted by score,
// descending and stable -- the reference's processing order.  Node ids preserve the order of the reference's (scaffold,
// side) tuples (ties between equal scores at a node are broken by the neighbour).
extern "C" int64_t besst_scaffold_prune_ambiguous(int64_t n_nodes, int64_t n_edges, const int32_t* eu, const int32_t* ev, const double* score,
                                                  const int64_t* order, uint8_t* removed, int64_t* amb_best, int64_t* amb_second) {
    if (n_nodes < 0 || n_edges < 0 || (n_edges > 0 && (!eu || !ev || !score || !order || !removed || !amb_best || !amb_second))) return -1;
    std::vector<int64_t> ptr((size_t)n_nodes + 1, 0);
    for (int64_t e = 0; e < n_edges; ++e) {
        if (eu[e] < 0 || eu[e] >= n_nodes || ev[e] < 0 || ev[e] >= n_nodes) return -1;
        ++ptr[(size_t)eu[e] + 1];
        ++ptr[(size_t)ev[e] + 1];
        removed[e] = 0;
    }
    for (int64_t i = 0; i < n_nodes; ++i) ptr[(size_t)i + 1] += ptr[(size_t)i];
    std::vector<int64_t> inc((size_t)(2 * n_edges)), fill(ptr.begin(), ptr.end() - 1);
    for (int64_t e = 0; e < n_edges; ++e) {
        inc[(size_t)fill[(size_t)eu[e]]++] = e;
        inc[(size_t)fill[(size_t)ev[e]]++] = e;
    }
    struct Item { double s; int32_t nbr; int64_t e; };
    std::vector<Item> items;
    int64_t n_amb = 0;
    auto at_node = [&](int32_t node) {
        items.clear();
        for (int64_t a = ptr[(size_t)node]; a < ptr[(size_t)node + 1]; ++a) {
            const int64_t e = inc[(size_t)a];
            if (removed[e]) continue;
            items.push_back(Item{score[e], eu[e] == node ? ev[e] : eu[e], e});
        }
        std::sort(items.begin(), items.end(), [](const Item& a, const Item& b) { return a.s < b.s || (a.s == b.s && a.nbr < b.nbr); });
        // zero-score edges go; of the others all but the best go -- or all of them when the two best are within 0.8 (:178-183)
        for (size_t k = 0; k < items.size(); ++k) {
            if (0 < items[k].s) continue;
            removed[items[k].e] = 1;
        }
        std::vector<Item>& nz = items;
        size_t m = 0;
        for (size_t k = 0; k < items.size(); ++k)
            if (0 < items[k].s) nz[m++] = items[k];   // stable partition in place (m <= k)
        if (m > 1) {
            if (nz[m - 2].s / nz[m - 1].s > 0.8) {
                amb_best[n_amb] = nz[m - 1].e;
                amb_second[n_amb] = nz[m - 2].e;
                ++n_amb;
                for (size_t k = 0; k < m; ++k) removed[nz[k].e] = 1;
            } else {
                for (size_t k = 0; k + 1 < m; ++k) removed[nz[k].e] = 1;
            }
        }
    };
    for (int64_t k = 0; k < n_edges; ++k) {
        const int64_t e = order[k];
        if (e < 0 || e >= n_edges) return -1;
        at_node(eu[e]);
        at_node(ev[e]);
    }
    return n_amb;
}
