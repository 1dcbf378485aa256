// kungfu_b200 runtime — communication graphs and topology generators for the CPU collective backend.
//
// Reference: srcs/go/plan/graph/graph.go:18-154 (digraph with a self-loop flag per node; a reduce graph is the reversed
// broadcast graph plus self-loops), srcs/go/plan/topology.go:5-160 (generators), srcs/go/plan/subgraph/subgraph.go:5-31
// (ring / binary tree over host masters), srcs/cpp/include/kungfu/mst.hpp:9-58 (Prim on symmetrised weights).
#pragma once
#include <cstdint>
#include <string>
#include <utility>
#include <vector>

#include "plan.hpp"

namespace kfrt {

class Graph {
public:
    explicit Graph(int n = 0) : nodes_(n) {}
    int size() const { return (int)nodes_.size(); }
    void add_edge(int i, int j); // i == j marks a self loop
    const std::vector<int>& prevs(int i) const { return nodes_[i].prevs; }
    const std::vector<int>& nexts(int i) const { return nodes_[i].nexts; }
    bool is_self_loop(int i) const { return nodes_[i].self_loop; }
    bool is_isolated(int i) const { return nodes_[i].prevs.empty() && nodes_[i].nexts.empty(); }
    Graph reversed() const;
    std::string digest() const; // canonical byte string
    std::string debug_string() const;
    // from a father[] array (father[i] == i marks a root). Returns false if not a forest. `roots` = number of roots.
    static bool from_forest_array(const std::vector<int>& father, Graph& out, int& roots);
    // true iff the graph is a tree rooted at `root` reaching all nodes (ignoring self loops)
    bool is_rooted_tree(int root) const;

private:
    struct Node {
        bool self_loop = false;
        std::vector<int> prevs, nexts;
    };
    std::vector<Node> nodes_;
};

// (reduce graph, broadcast graph)
struct GraphPair {
    Graph reduce, bcast;
};

Graph default_reduce_graph(const Graph& bcast); // reverse + self loops

Graph gen_tree(const PeerList& peers);               // host stars + star over masters
Graph gen_binary_tree(int k);
Graph gen_binary_tree_star(const PeerList& peers, int offset = 0);
std::vector<Graph> gen_multi_binary_tree_star(const PeerList& peers);
Graph gen_multi_star_one(const PeerList& peers, int root_master);
std::vector<Graph> gen_multi_star(const PeerList& peers);
Graph gen_star_bcast(int k, int root);
GraphPair gen_circular_pair(int k, int r);
// cross-host variants over masters only (other nodes isolated)
GraphPair gen_cross_circular_pair(int n, const std::vector<int>& masters, int r);
Graph gen_cross_binary_tree(int n, const std::vector<int>& masters);

enum class Strategy : int {
    Star = 0, Ring = 1, Clique = 2, Tree = 3, BinaryTree = 4, BinaryTreeStar = 5, MultiBinaryTreeStar = 6, MultiStar = 7, Auto = 8
};
Strategy parse_strategy(const std::string& name); // throws on unknown
std::string strategy_name(Strategy s);
std::vector<std::string> strategy_names();
Strategy auto_select(const PeerList& peers); // STAR on one host, BINARY_TREE_STAR otherwise

std::vector<GraphPair> gen_global_strategies(const PeerList& peers, Strategy s);
std::vector<GraphPair> gen_local_strategies(const PeerList& peers);
std::vector<GraphPair> gen_cross_strategies(const PeerList& peers, Strategy s);
std::string digest_strategies(const std::vector<GraphPair>& sl);

// Prim's minimum spanning tree on the symmetrised n x n weight matrix; returns the n-1 edges (u, v)
std::vector<std::pair<int, int>> minimum_spanning_tree(const std::vector<double>& weights, int n);
// father[] array of the MST rooted at `root`
std::vector<int> mst_forest_array(const std::vector<double>& weights, int n, int root = 0);

} // namespace kfrt
