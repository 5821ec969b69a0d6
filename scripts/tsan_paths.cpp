#include <stdint.h>
#include <stdio.h>
#include <random>
#include <vector>
#include <set>
#include "../include/besst_b200.h"
int main() {
    std::mt19937 rng(5);
    const int n_scaf = 3000;
    std::vector<std::set<int>> adjset(2 * n_scaf);
    std::vector<std::vector<std::pair<int,int>>> adj(2 * n_scaf);
    for (int s = 0; s < n_scaf; ++s) { adj[2*s].push_back({2*s+1, -1}); adj[2*s+1].push_back({2*s, -1}); }
    for (int k = 0; k < 2 * n_scaf; ++k) {
        int a = rng() % n_scaf, b = a + 1 + rng() % 11; if (b >= n_scaf) continue;
        int u = 2*a + rng()%2, v = 2*b + rng()%2;
        if (adjset[u].count(v)) continue;
        adjset[u].insert(v); adjset[v].insert(u);
        int w = 1 + rng() % 40;
        adj[u].push_back({v, w}); adj[v].push_back({u, w});
    }
    std::vector<int64_t> ptr(2*n_scaf+1, 0); std::vector<int32_t> nodes, links;
    for (int i = 0; i < 2*n_scaf; ++i) { for (auto& e : adj[i]) { nodes.push_back(e.first); links.push_back(e.second); } ptr[i+1] = (int64_t)nodes.size(); }
    std::vector<uint8_t> is_end(2*n_scaf, 0); std::vector<int32_t> order;
    for (int s = 0; s < n_scaf; s += 6) { is_end[2*s] = is_end[2*s+1] = 1; order.push_back(2*s); order.push_back(2*s+1); }
    besst_paths* p = besst_paths_between(2*n_scaf, ptr.data(), nodes.data(), links.data(), is_end.data(), order.data(), (int64_t)order.size(), 5000, 0.0, 0, 0, 8);
    printf("paths %lld pops %lld hit %d\n", (long long)besst_paths_count(p), (long long)besst_paths_pops(p), besst_paths_hit_threshold(p));
    besst_paths_free(p);
    return 0;
}
