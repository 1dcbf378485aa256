// Implementation of graph.hpp (the header carries the design notes and the reference file:line mapping).
#include "graph.hpp"

#include <algorithm>
#include <limits>
#include <map>
#include <queue>
#include <sstream>

#include "base.hpp"

namespace kfrt {

void Graph::add_edge(int i, int j) {
    if (i < 0 || j < 0 || i >= size() || j >= size()) fail("graph: edge out of range");
    if (i == j) {
        nodes_[i].self_loop = true;
        return;
    }
    nodes_[i].nexts.push_back(j);
    nodes_[j].prevs.push_back(i);
}

Graph Graph::reversed() const {
    Graph r(size());
    for (int i = 0; i < size(); ++i) {
        r.nodes_[i].self_loop = nodes_[i].self_loop;
        for (int j : nodes_[i].nexts) r.add_edge(j, i);
    }
    return r;
}

static void put32(std::string& s, int32_t v) { s.append(reinterpret_cast<const char*>(&v), 4); }

std::string Graph::digest() const {
    std::string d;
    put32(d, size());
    for (auto& n : nodes_) {
        std::vector<int> vs = n.nexts;
        std::sort(vs.begin(), vs.end());
        put32(d, n.self_loop ? 1 : 0);
        put32(d, (int32_t)vs.size());
        for (int j : vs) put32(d, j);
    }
    return d;
}

std::string Graph::debug_string() const {
    std::ostringstream os;
    os << "[" << size() << "]{";
    for (int i = 0; i < size(); ++i)
        if (nodes_[i].self_loop) os << "(" << i << ")";
    for (int i = 0; i < size(); ++i)
        for (int j : nodes_[i].nexts) os << "(" << i << "->" << j << ")";
    os << "}";
    return os.str();
}

bool Graph::from_forest_array(const std::vector<int>& father, Graph& out, int& roots) {
    const int n = (int)father.size();
    Graph g(n);
    roots = 0;
    for (int i = 0; i < n; ++i) {
        const int f = father[i];
        if (f < 0 || f >= n) return false;
        if (f == i) ++roots;
        else g.add_edge(f, i);
    }
    // acyclic check: every node must reach a root by following fathers in < n steps
    for (int i = 0; i < n; ++i) {
        int v = i, steps = 0;
        while (father[v] != v) {
            v = father[v];
            if (++steps > n) return false;
        }
    }
    out = g;
    return true;
}

bool Graph::is_rooted_tree(int root) const {
    const int n = size();
    if (n == 0) return true;
    int edges = 0;
    for (int i = 0; i < n; ++i) {
        edges += (int)nodes_[i].nexts.size();
        if (i != root && nodes_[i].prevs.size() != 1) return false;
    }
    if (edges != n - 1 || !nodes_[root].prevs.empty()) return false;
    std::vector<bool> seen(n, false);
    std::queue<int> q;
    q.push(root);
    seen[root] = true;
    int cnt = 0;
    while (!q.empty()) {
        int v = q.front();
        q.pop();
        ++cnt;
        for (int j : nodes_[v].nexts)
            if (!seen[j]) {
                seen[j] = true;
                q.push(j);
            }
    }
    return cnt == n;
}

Graph default_reduce_graph(const Graph& bcast) {
    Graph g = bcast.reversed();
    for (int i = 0; i < g.size(); ++i) g.add_edge(i, i);
    return g;
}

// ---- generators -----------------------------------------------------------------------------------
static void host_stars(const PeerList& peers, Graph& g, std::vector<int>& masters) {
    std::vector<int> parents;
    peers.partition_by_host(masters, parents);
    for (int r = 0; r < (int)peers.size(); ++r)
        if (parents[r] != r) g.add_edge(parents[r], r);
}

Graph gen_tree(const PeerList& peers) {
    Graph g((int)peers.size());
    std::vector<int> masters;
    host_stars(peers, g, masters);
    for (size_t i = 1; i < masters.size(); ++i) g.add_edge(masters[0], masters[i]);
    return g;
}

Graph gen_binary_tree(int k) {
    Graph g(k);
    for (int i = 0; i < k; ++i) {
        if (2 * i + 1 < k) g.add_edge(i, 2 * i + 1);
        if (2 * i + 2 < k) g.add_edge(i, 2 * i + 2);
    }
    return g;
}

Graph gen_binary_tree_star(const PeerList& peers, int offset) {
    Graph g((int)peers.size());
    std::vector<int> masters;
    host_stars(peers, g, masters);
    const int k = (int)masters.size();
    if (k > 1) {
        auto idx = [&](int i) { return (i + offset) % k; };
        for (int i = 0; i < k; ++i) {
            if (2 * i + 1 < k) g.add_edge(masters[idx(i)], masters[idx(2 * i + 1)]);
            if (2 * i + 2 < k) g.add_edge(masters[idx(i)], masters[idx(2 * i + 2)]);
        }
    }
    return g;
}

std::vector<Graph> gen_multi_binary_tree_star(const PeerList& peers) {
    std::vector<int> masters, parents;
    peers.partition_by_host(masters, parents);
    std::vector<Graph> gs;
    for (int i = 0; i < (int)masters.size(); ++i) gs.push_back(gen_binary_tree_star(peers, i));
    return gs;
}

Graph gen_multi_star_one(const PeerList& peers, int root_master) {
    Graph g((int)peers.size());
    std::vector<int> masters;
    host_stars(peers, g, masters);
    const int k = (int)masters.size();
    if (k > 1)
        for (int i = 0; i < k; ++i)
            if (i != root_master) g.add_edge(masters[root_master], masters[i]);
    return g;
}

std::vector<Graph> gen_multi_star(const PeerList& peers) {
    std::vector<int> masters, parents;
    peers.partition_by_host(masters, parents);
    std::vector<Graph> gs;
    for (int i = 0; i < (int)masters.size(); ++i) gs.push_back(gen_multi_star_one(peers, i));
    return gs;
}

Graph gen_star_bcast(int k, int root) {
    Graph g(k);
    for (int i = 0; i < k; ++i)
        if (i != root) g.add_edge(root, i);
    return g;
}

GraphPair gen_circular_pair(int k, int r) {
    GraphPair p{Graph(k), Graph(k)};
    for (int i = 0; i < k; ++i) p.reduce.add_edge(i, i);
    for (int i = 1; i < k; ++i) {
        p.reduce.add_edge((r + i) % k, (r + i + 1) % k);
        p.bcast.add_edge((r + i - 1) % k, (r + i) % k);
    }
    return p;
}

GraphPair gen_cross_circular_pair(int n, const std::vector<int>& vs, int r) {
    GraphPair p{Graph(n), Graph(n)};
    const int k = (int)vs.size();
    for (int i = 0; i < k; ++i) p.reduce.add_edge(vs[i], vs[i]);
    for (int i = 1; i < k; ++i) {
        p.reduce.add_edge(vs[(r + i) % k], vs[(r + i + 1) % k]);
        p.bcast.add_edge(vs[(r + i - 1) % k], vs[(r + i) % k]);
    }
    return p;
}

Graph gen_cross_binary_tree(int n, const std::vector<int>& vs) {
    Graph g(n);
    const int k = (int)vs.size();
    for (int i = 0; i < k; ++i) {
        if (2 * i + 1 < k) g.add_edge(vs[i], vs[2 * i + 1]);
        if (2 * i + 2 < k) g.add_edge(vs[i], vs[2 * i + 2]);
    }
    return g;
}

// ---- strategies -----------------------------------------------------------------------------------
static const std::vector<std::pair<std::string, Strategy>>& strategy_table() {
    static const std::vector<std::pair<std::string, Strategy>> t = {
        {"STAR", Strategy::Star}, {"MULTI_STAR", Strategy::MultiStar}, {"RING", Strategy::Ring}, {"CLIQUE", Strategy::Clique},
        {"TREE", Strategy::Tree}, {"BINARY_TREE", Strategy::BinaryTree}, {"BINARY_TREE_STAR", Strategy::BinaryTreeStar},
        {"MULTI_BINARY_TREE_STAR", Strategy::MultiBinaryTreeStar}, {"AUTO", Strategy::Auto}};
    return t;
}
Strategy parse_strategy(const std::string& name) {
    std::string n;
    for (char c : name) n.push_back((char)toupper(c));
    for (auto& kv : strategy_table())
        if (kv.first == n) return kv.second;
    fail("unknown all-reduce strategy '" + name + "'");
}
std::string strategy_name(Strategy s) {
    for (auto& kv : strategy_table())
        if (kv.second == s) return kv.first;
    return "?";
}
std::vector<std::string> strategy_names() {
    std::vector<std::string> v;
    for (auto& kv : strategy_table()) v.push_back(kv.first);
    return v;
}
Strategy auto_select(const PeerList& peers) { return peers.host_count() <= 1 ? Strategy::Star : Strategy::BinaryTreeStar; }

static GraphPair simple(const Graph& bcast) { return GraphPair{default_reduce_graph(bcast), bcast}; }

std::vector<GraphPair> gen_global_strategies(const PeerList& peers, Strategy s) {
    const int k = (int)peers.size();
    if (s == Strategy::Auto) s = auto_select(peers);
    std::vector<GraphPair> sl;
    switch (s) {
    case Strategy::Star: sl.push_back(simple(gen_star_bcast(k, 0))); break;
    case Strategy::MultiStar:
        for (auto& g : gen_multi_star(peers)) sl.push_back(simple(g));
        break;
    case Strategy::Clique:
        for (int r = 0; r < k; ++r) sl.push_back(simple(gen_star_bcast(k, r)));
        break;
    case Strategy::Ring:
        for (int r = 0; r < k; ++r) sl.push_back(gen_circular_pair(k, r));
        break;
    case Strategy::Tree: sl.push_back(simple(gen_tree(peers))); break;
    case Strategy::BinaryTree: sl.push_back(simple(gen_binary_tree(k))); break;
    case Strategy::BinaryTreeStar: sl.push_back(simple(gen_binary_tree_star(peers))); break;
    case Strategy::MultiBinaryTreeStar:
        for (auto& g : gen_multi_binary_tree_star(peers)) sl.push_back(simple(g));
        break;
    default: fail("bad strategy");
    }
    if (sl.empty()) sl.push_back(simple(Graph(k)));
    return sl;
}

std::vector<GraphPair> gen_local_strategies(const PeerList& peers) {
    std::vector<int> masters, parents;
    peers.partition_by_host(masters, parents);
    Graph g;
    int roots = 0;
    if (!Graph::from_forest_array(parents, g, roots) || roots != (int)masters.size()) fail("local strategy: bad host forest");
    return {simple(g)};
}

std::vector<GraphPair> gen_cross_strategies(const PeerList& peers, Strategy s) {
    const int n = (int)peers.size();
    std::vector<int> masters, parents;
    peers.partition_by_host(masters, parents);
    std::vector<GraphPair> sl;
    if (s == Strategy::Ring) {
        for (int r = 0; r < (int)masters.size(); ++r) sl.push_back(gen_cross_circular_pair(n, masters, r));
    } else {
        sl.push_back(simple(gen_cross_binary_tree(n, masters)));
    }
    return sl;
}

std::string digest_strategies(const std::vector<GraphPair>& sl) {
    std::string d;
    for (auto& p : sl) d += p.reduce.digest() + p.bcast.digest();
    return d;
}

// ---- MST ------------------------------------------------------------------------------------------
std::vector<std::pair<int, int>> minimum_spanning_tree(const std::vector<double>& w, int n) {
    if ((int)w.size() != n * n) fail("mst: weight matrix must be n x n");
    auto W = [&](int i, int j) { return (w[(size_t)i * n + j] + w[(size_t)j * n + i]) / 2; };
    std::vector<std::pair<int, int>> edges;
    if (n <= 1) return edges;
    std::vector<bool> used(n, false);
    std::vector<double> best(n);
    std::vector<int> from(n, 0);
    used[0] = true;
    for (int i = 0; i < n; ++i) best[i] = W(0, i);
    for (int it = 1; it < n; ++it) {
        int k = -1;
        for (int j = 0; j < n; ++j)
            if (!used[j] && (k < 0 || best[j] < best[k])) k = j;
        used[k] = true;
        edges.emplace_back(from[k], k);
        for (int j = 0; j < n; ++j)
            if (!used[j] && W(k, j) < best[j]) {
                best[j] = W(k, j);
                from[j] = k;
            }
    }
    return edges;
}

std::vector<int> mst_forest_array(const std::vector<double>& w, int n, int root) {
    auto edges = minimum_spanning_tree(w, n);
    std::vector<std::vector<int>> adj(n);
    for (auto& e : edges) {
        adj[e.first].push_back(e.second);
        adj[e.second].push_back(e.first);
    }
    std::vector<int> father(n, -1);
    std::queue<int> q;
    father[root] = root;
    q.push(root);
    while (!q.empty()) {
        int v = q.front();
        q.pop();
        for (int u : adj[v])
            if (father[u] < 0) {
                father[u] = v;
                q.push(u);
            }
    }
    return father;
}

} // namespace kfrt
