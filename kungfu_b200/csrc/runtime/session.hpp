// kungfu_b200 runtime — Session: the CPU collective engine for one cluster version.
//
// Reference: srcs/go/kungfu/session/{session.go:21-321, allreduce.go, allgather.go, strategy.go, shard.go,
// adaptation.go, monitoring.go, adaptiveStrategies.go}. A session is immutable for one membership; an all-reduce is cut
// into ~1 MiB chunks, each chunk is reduced along one (reduce graph, broadcast graph) pair chosen by hashing the chunk
// index (or the tensor name), all chunks progress concurrently. Every rank ends with bit-identical results because
// each chunk has exactly one reduction order (at its root).
//
// On a B200 box this engine carries (a) every GPU-less test/"gloo"-style configuration, (b) small host-side
// collectives of the control plane (barrier, consensus, int max, handle exchange); tensors in HBM never use it.
#pragma once
#include <atomic>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "graph.hpp"
#include "plan.hpp"
#include "reduce.hpp"
#include "transport.hpp"

namespace kfrt {

struct Workspace {
    const void* send = nullptr;
    void* recv = nullptr;
    size_t count = 0;
    int dtype = kF32;
    int op = kSum;
    std::string name;
    bool inplace() const { return send == recv; }
    size_t bytes() const { return count * dtype_size(dtype); }
};

struct StrategyStat {
    std::mutex mu;
    double throughput = 0; // bytes/s of the last window
    double ref_throughput = 0;
    uint64_t acc_size = 0;
    double first_begin = 0, last_end = 0;
    void update(double begin, double end, uint64_t size);
    void reset();
};

class Session {
public:
    Session(Strategy strategy, const PeerID& self, const PeerList& peers, Client* client, CollectiveEndpoint* coll);

    int rank() const { return rank_; }
    int size() const { return (int)peers_.size(); }
    int local_rank() const { return local_rank_; }
    int local_size() const { return local_size_; }
    int host_count() const { return host_count_; }
    const PeerList& peers() const { return peers_; }

    void barrier();
    bool consensus(const void* data, size_t len, const std::string& name);
    void all_reduce(const Workspace& w);
    void monitored_all_reduce(const Workspace& w, const std::vector<int32_t>& tree); // tree may be empty
    void cross_all_reduce(const Workspace& w);
    void subset_all_reduce(const Workspace& w, const std::vector<int32_t>& forest); // independent all-reduce per tree
    void reduce(const Workspace& w);          // result on rank 0
    void broadcast(const Workspace& w);       // from rank 0
    void local_reduce(const Workspace& w);
    void local_broadcast(const Workspace& w);
    void gather(const Workspace& w);          // recv holds size()*count on rank 0
    void all_gather(const Workspace& w);      // recv holds size()*count everywhere

    void set_tree(const std::vector<int32_t>& father); // consensus-checked switch of the global strategy
    int num_strategies() const { return (int)global_.size(); }
    std::vector<double> peer_latencies();     // seconds, 0 for self, <0 unreachable
    std::vector<double> egress_rates();

    void calc_stats();
    void log_stats();
    void print_strategy_stats();
    bool check_interference();
    double last_throughput();
    uint64_t next_seq() { return seq_++; } // session-scoped op numbering: identical on survivors and newcomers

private:
    struct Strat {
        GraphPair g;
        std::shared_ptr<StrategyStat> stat;
    };
    using StratList = std::vector<Strat>;
    static StratList wrap(const std::vector<GraphPair>& gs);
    void run_graphs(const Workspace& w, const Graph* const* graphs, int ngraphs);
    void run_strategies(const Workspace& w, const StratList& sl, bool monitored);
    uint64_t strategy_hash(size_t chunk, const std::string& name) const;
    void forward(const Workspace& w) const;

    PeerID self_;
    PeerList peers_;
    int rank_ = 0, local_rank_ = 0, local_size_ = 1, host_count_ = 1;
    Client* client_;
    CollectiveEndpoint* coll_;
    StratList local_, global_, cross_;
    std::mutex mu_; // guards strategy switches
    bool hash_by_name_ = false;
    std::vector<double> stat_log_;
    std::atomic<uint64_t> seq_{0};
};

} // namespace kfrt
