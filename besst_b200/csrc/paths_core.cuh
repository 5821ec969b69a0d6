// paths_core.cuh -- the path search between scaffolds on a CSR rendering of G_prime.
//
// Restates, for one start node, ExtendLargeScaffolds.find_all_paths_for_start_node_DFS_dynamic_programming_ish
// (ExtendLargeScaffolds.py:526-663, the default traversal: runBESST:101) and, for the paths it finds, ScorePaths
// (:28-133).  What the reference's Python does with tuples, lists, sets and heapq is done here with integers:
//   node id        2 * rank(scaffold) + (side == 'R'), ranks ascending in the scaffold key: integer order == the order of
//                  the reference's (scaffold, side) tuples, the other end of a contig is id ^ 1
//   path           a chain through an arena (node, parent): `path + [start]` is one new arena cell, shared by all children
//   heap           binary heap over (nr_links, node, path) -- the reference's heapq tuples; ties on (nr_links, node)
//                  (every contig-crossing entry carries 2**16) are decided by comparing the paths lexicographically,
//                  like Python compares the lists
//   ctg_ends_in_path  == the set of the current path (it is rebuilt from the path at every link expansion, :651-653):
//                  membership is a walk up the chain
//   bad_ctgs       ONE set per search: the reference passes the same set object along every entry (:553,:653)
//   head_dict      small open-addressing map node -> (nr_bad_nbrs, bad_link_count)
// The start nodes of BetweenScaffolds (:665-712) are independent once `already_visited` and `end` are expressed through
// the position of a node in the start order (visited: processed earlier; end: in the end set and not processed yet), so
// the searches run in parallel -- one host thread per start node in besst_paths.cu, and the same source is device code.
//
// Everything is written against caller-provided scratch with capacities; running out reports OVERFLOW and the host
// wrapper repeats the search with more room.
#pragma once

#include <stdint.h>

#ifdef __CUDACC__
#define PATHS_HD __host__ __device__ __forceinline__
#else
#define PATHS_HD inline
#endif

namespace paths {

constexpr int MAX_PATH = 128;          // the reference stops extending at 100 nodes (:559)
constexpr int32_t CROSS_LINKS = 1 << 16;   // priority of a contig-crossing entry (:597)
constexpr int ST_OK = 0, ST_OVERFLOW = 1;

struct Graph {
    int64_t n_nodes;
    const int64_t* adj_ptr;     // [n_nodes + 1]
    const int32_t* adj_node;    // neighbours in any order
    const int32_t* adj_links;   // nr_links of the edge; the contig edge (id ^ 1) carries -1
    const int32_t* order_pos;   // [n_nodes] position in the start order, INT32_MAX when the node is no start node
    const uint8_t* is_end;      // [n_nodes] member of the initial `end` set
};

struct Entry {
    int32_t links, node, tail, depth;   // tail: arena cell of the last node of `path` (-1: empty), depth = len(path)
    int32_t nbad, badlinks;
};

struct Scratch {
    int32_t* arena_node;   // [arena_cap]
    int32_t* arena_parent;
    int32_t arena_cap, arena_n;
    Entry* heap;           // [heap_cap]
    int32_t heap_cap, heap_n;
    int32_t* map_key;      // [map_cap] open addressing, -1 empty: head_dict
    int32_t* map_a;
    int32_t* map_b;
    int32_t map_cap, map_n;
    int32_t* set_key;      // [set_cap] bad_ctgs
    int32_t set_cap, set_n;
    int32_t* found_tail;   // [found_cap] arena cells of the found paths, in the order found
    int32_t found_cap, found_n;
};

PATHS_HD uint32_t hash32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}

// path of arena cell `tail` (depth nodes) into out[0 .. depth)
PATHS_HD void materialise(const Scratch& S, int32_t tail, int32_t depth, int32_t* out) {
    for (int32_t i = depth - 1; i >= 0; --i) {
        out[i] = S.arena_node[tail];
        tail = S.arena_parent[tail];
    }
}

PATHS_HD bool in_path(const Scratch& S, int32_t tail, int32_t node) {
    while (tail >= 0) {
        if (S.arena_node[tail] == node) return true;
        tail = S.arena_parent[tail];
    }
    return false;
}

// a < b as the reference's heap tuples (nr_links, (node, path), ...) compare
PATHS_HD bool entry_less(const Scratch& S, const Entry& a, const Entry& b) {
    if (a.links != b.links) return a.links < b.links;
    if (a.node != b.node) return a.node < b.node;
    if (a.tail == b.tail) return false;
    int32_t pa[MAX_PATH], pb[MAX_PATH];
    materialise(S, a.tail, a.depth, pa);
    materialise(S, b.tail, b.depth, pb);
    const int32_t n = a.depth < b.depth ? a.depth : b.depth;
    for (int32_t i = 0; i < n; ++i)
        if (pa[i] != pb[i]) return pa[i] < pb[i];
    return a.depth < b.depth;
}

PATHS_HD bool heap_push(Scratch& S, const Entry& e) {
    if (S.heap_n >= S.heap_cap) return false;
    int32_t i = S.heap_n++;
    while (i > 0) {
        const int32_t p = (i - 1) >> 1;
        if (!entry_less(S, e, S.heap[p])) break;
        S.heap[i] = S.heap[p];
        i = p;
    }
    S.heap[i] = e;
    return true;
}

PATHS_HD Entry heap_pop(Scratch& S) {
    const Entry top = S.heap[0];
    const Entry last = S.heap[--S.heap_n];
    int32_t i = 0;
    for (;;) {
        int32_t c = 2 * i + 1;
        if (c >= S.heap_n) break;
        if (c + 1 < S.heap_n && entry_less(S, S.heap[c + 1], S.heap[c])) ++c;
        if (!entry_less(S, S.heap[c], last)) break;
        S.heap[i] = S.heap[c];
        i = c;
    }
    if (S.heap_n > 0) S.heap[i] = last;
    return top;
}

// head_dict: -> slot of `node` (existing or a new one with *fresh = true), -1 when the table is full
PATHS_HD int32_t map_slot(Scratch& S, int32_t node, bool insert, bool* fresh) {
    uint32_t h = hash32((uint32_t)node) & (uint32_t)(S.map_cap - 1);
    *fresh = false;
    for (int32_t probe = 0; probe < S.map_cap; ++probe) {
        const int32_t k = S.map_key[h];
        if (k == node) return (int32_t)h;
        if (k < 0) {
            if (!insert) return -2;
            if (2 * (S.map_n + 1) > S.map_cap) return -1;
            S.map_key[h] = node;
            ++S.map_n;
            *fresh = true;
            return (int32_t)h;
        }
        h = (h + 1) & (uint32_t)(S.map_cap - 1);
    }
    return -1;
}

// bad_ctgs: 1 present, 0 absent (and inserted when `insert`), -1 full
PATHS_HD int set_test(Scratch& S, int32_t node, bool insert) {
    uint32_t h = hash32((uint32_t)node ^ 0x9e3779b9u) & (uint32_t)(S.set_cap - 1);
    for (int32_t probe = 0; probe < S.set_cap; ++probe) {
        const int32_t k = S.set_key[h];
        if (k == node) return 1;
        if (k < 0) {
            if (!insert) return 0;
            if (2 * (S.set_n + 1) > S.set_cap) return -1;
            S.set_key[h] = node;
            ++S.set_n;
            return 0;
        }
        h = (h + 1) & (uint32_t)(S.set_cap - 1);
    }
    return -1;
}

// One start node (ExtendLargeScaffolds.py:526-663 with is_withing_scaf = 0, max_path_length_allowed = 2**32).
// my_pos: position of `start0` in the start order.  The caller has cleared the scratch (keys = -1, counts = 0).
// -> ST_OK / ST_OVERFLOW; *hit_threshold as param.hit_path_threshold (:561)
PATHS_HD int search(const Graph& G, int32_t start0, int32_t my_pos, int64_t path_threshold, Scratch& S, int* hit_threshold) {
    const int32_t forbidden = start0 ^ 1;
    Entry first;
    first.links = 0; first.node = start0; first.tail = -1; first.depth = 0; first.nbad = 0; first.badlinks = 0;
    if (!heap_push(S, first)) return ST_OVERFLOW;
    int64_t counter = 0;
    int32_t cur_len = 0;   // len(path) as the loop head sees it (:559)
    while (S.heap_n > 0) {
        ++counter;
        if (counter > path_threshold || cur_len > 100) {
            *hit_threshold = 1;
            break;
        }
        const Entry e = heap_pop(S);
        const int32_t start = e.node;
        cur_len = e.depth;
        {
            bool fresh;
            const int32_t s = map_slot(S, start, false, &fresh);
            if (s >= 0 && e.nbad > S.map_a[s] && e.badlinks > S.map_b[s]) continue;   // strictly worse than a path seen before (:566-569)
        }
        const int32_t prev_node = e.tail >= 0 ? S.arena_node[e.tail] : start;
        if (S.arena_n >= S.arena_cap) return ST_OVERFLOW;
        const int32_t cell = S.arena_n++;   // path = path + [start]
        S.arena_node[cell] = start;
        S.arena_parent[cell] = e.tail;
        const int32_t depth = e.depth + 1;
        cur_len = depth;
        if (depth >= MAX_PATH) return ST_OVERFLOW;   // cannot happen: the loop head stops at 101
        if (G.order_pos[start] < my_pos || start == forbidden) continue;      // already_visited / forbidden (:581)
        if (G.is_end[start] && G.order_pos[start] > my_pos) {                 // start in end (:584)
            if (S.found_n >= S.found_cap) return ST_OVERFLOW;
            S.found_tail[S.found_n++] = cell;
            continue;
        }
        if ((prev_node >> 1) != (start >> 1)) {
            // arrived over a link: cross the contig (:596-646)
            const int32_t other = start ^ 1;
            if (other == forbidden) continue;
            int32_t add_nbrs = 0, add_links = 0;
            for (int64_t a = G.adj_ptr[start]; a < G.adj_ptr[start + 1]; ++a) {
                const int32_t nbr = G.adj_node[a];
                if ((nbr >> 1) == (start >> 1)) continue;
                if (in_path(S, cell, nbr)) continue;          // ctg_ends_in_path == the nodes of the path
                const int t = set_test(S, nbr, true);          // bad_ctgs: one set for the whole search
                if (t < 0) return ST_OVERFLOW;
                if (t == 0) {
                    ++add_nbrs;
                    add_links += G.adj_links[a];
                }
            }
            const int32_t tot_a = e.nbad + add_nbrs, tot_b = e.badlinks + add_links;
            bool fresh;
            const int32_t so = map_slot(S, other, true, &fresh);
            if (so < 0) return ST_OVERFLOW;
            if (fresh) {
                S.map_a[so] = tot_a; S.map_b[so] = tot_b;
                bool fresh2;
                const int32_t ss = map_slot(S, start, true, &fresh2);
                if (ss < 0) return ST_OVERFLOW;
                S.map_a[ss] = tot_a; S.map_b[ss] = tot_b;
            } else if (tot_a > S.map_a[so] && tot_b > S.map_b[so]) {
                continue;
            }
            Entry c;
            c.links = CROSS_LINKS; c.node = other; c.tail = cell; c.depth = depth; c.nbad = tot_a; c.badlinks = tot_b;
            if (!heap_push(S, c)) return ST_OVERFLOW;
        } else {
            // arrived over the contig (or the very start): follow every link that does not lead back into the path (:648-655)
            for (int64_t a = G.adj_ptr[start]; a < G.adj_ptr[start + 1]; ++a) {
                const int32_t node = G.adj_node[a];
                if (node == forbidden || in_path(S, cell, node)) continue;
                Entry c;
                c.links = G.adj_links[a]; c.node = node; c.tail = cell; c.depth = depth; c.nbad = e.nbad; c.badlinks = e.badlinks;
                if (!heap_push(S, c)) return ST_OVERFLOW;
            }
        }
    }
    return ST_OK;
}

// ScorePaths' two connectivity measures (ExtendLargeScaffolds.py:31-110).  -> good and bad link weight; the caller
// forms the score (good / bad, or good when bad == 0; the contamination variant halves good first)
PATHS_HD void connectivity(const Graph& G, const int32_t* path, int32_t len, bool contamination, int64_t* good, int64_t* bad) {
    int64_t g = 0, b = 0;
    for (int32_t i = 0; i < len; ++i) {
        const int32_t node = path[i];
        for (int64_t a = G.adj_ptr[node]; a < G.adj_ptr[node + 1]; ++a) {
            const int32_t nbr = G.adj_node[a];
            if ((nbr >> 1) == (node >> 1)) continue;
            const int64_t w = G.adj_links[a];
            // position of nbr in the path with the opposite parity (nodes_odd for an even i, nodes_even for an odd i), and
            // whether it was visited before i
            bool in_opposite = false, visited = false;
            for (int32_t j = 0; j < len; ++j)
                if (path[j] == nbr) {
                    if ((j & 1) != (i & 1)) in_opposite = true;
                    if (j < i) visited = true;
                }
            if (contamination) {
                if (in_opposite) g += w; else b += w;
            } else if ((i & 1) == 0) {
                if (in_opposite) { if (!visited) g += w; }
                else b += w;
            } else {
                if (!in_opposite) b += w;
                else if (!visited) b += w;
            }
        }
    }
    *good = g;
    *bad = b;
}

}  // namespace paths
