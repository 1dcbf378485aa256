#include "session.hpp"

#include <atomic>

#include <cstring>

namespace kfrt {

void StrategyStat::update(double begin, double end, uint64_t size) {
    std::lock_guard<std::mutex> g(mu);
    if (first_begin == 0) first_begin = begin;
    if (end > last_end) last_end = end;
    acc_size += size;
}
void StrategyStat::reset() {
    acc_size = 0;
    first_begin = 0;
    last_end = 0;
}

Session::StratList Session::wrap(const std::vector<GraphPair>& gs) {
    StratList sl;
    for (auto& g : gs) sl.push_back(Strat{g, std::make_shared<StrategyStat>()});
    return sl;
}

Session::Session(Strategy strategy, const PeerID& self, const PeerList& peers, Client* client, CollectiveEndpoint* coll)
    : self_(self), peers_(peers), client_(client), coll_(coll) {
    rank_ = peers_.rank_of(self_);
    if (rank_ < 0) fail("session: self " + self_.str() + " is not in the peer list " + peers_.str());
    local_rank_ = peers_.local_rank_of(self_);
    local_size_ = peers_.local_size_of(self_);
    host_count_ = peers_.host_count();
    if (strategy == Strategy::Auto) strategy = auto_select(peers_);
    local_ = wrap(gen_local_strategies(peers_));
    global_ = wrap(gen_global_strategies(peers_, strategy));
    cross_ = wrap(gen_cross_strategies(peers_, strategy));
    hash_by_name_ = Config::get().strategy_hash == "NAME";
}

uint64_t Session::strategy_hash(size_t chunk, const std::string& name) const {
    if (!hash_by_name_) return chunk;
    uint64_t h = 0;
    for (unsigned char c : name) h += (uint64_t)c * c;
    return h;
}

void Session::forward(const Workspace& w) const {
    if (!w.inplace() && w.bytes()) memcpy(w.recv, w.send, w.bytes());
}

// One chunk along a sequence of graphs. In a graph where this rank has a self loop it is an *accumulating* node:
// it receives from all predecessors (any order), folds each arrival into recv, then forwards the partial result. In a
// graph without a self loop it is a *relay*: it receives at most one message straight into recv and passes it on.
void Session::run_graphs(const Workspace& w, const Graph* const* graphs, int ngraphs) {
    if (w.count == 0) return;
    bool isolated = true;
    for (int i = 0; i < ngraphs; ++i)
        if (!graphs[i]->is_isolated(rank_)) isolated = false;
    if (isolated) {
        forward(w);
        return;
    }
    const size_t bytes = w.bytes();
    int recv_count = 0;
    std::mutex mu;
    auto effective = [&]() -> const void* { return (recv_count > 0 || w.inplace()) ? w.recv : w.send; };

    for (int gi = 0; gi < ngraphs; ++gi) {
        const Graph& g = *graphs[gi];
        const std::vector<int>& prevs = g.prevs(rank_);
        const std::vector<int>& nexts = g.nexts(rank_);
        if (g.is_self_loop(rank_)) {
            std::string err = parallel_for((int)prevs.size(), [&](int i) {
                Bytes m = coll_->recv(peers_[prevs[i]], w.name);
                if (m.size() != bytes)
                    fail("all-reduce chunk size mismatch for " + w.name + ": " + peers_[prevs[i]].str() + " sent " + std::to_string(m.size()) +
                         " bytes, this rank has " + std::to_string(bytes) + " (do all workers pass tensors of the same shape and dtype?)");
                std::lock_guard<std::mutex> lk(mu);
                transform2(w.recv, effective(), m.data(), w.count, w.dtype, w.op);
                ++recv_count;
            });
            if (!err.empty()) fail(err);
            const void* src = effective();
            err = parallel_for((int)nexts.size(), [&](int i) { client_->send(peers_[nexts[i]], w.name, src, bytes, kConnCollective, kNoFlag); });
            if (!err.empty()) fail(err);
        } else {
            if (prevs.size() > 1) KF_LOGE("more than one relay input at node %d", rank_);
            if (prevs.empty() && recv_count == 0) {
                forward(w);
            } else {
                for (int p : prevs) {
                    coll_->recv_into(peers_[p], w.name, w.recv, bytes);
                    ++recv_count;
                }
            }
            const void* src = effective();
            std::string err = parallel_for((int)nexts.size(), [&](int i) { client_->send(peers_[nexts[i]], w.name, src, bytes, kConnCollective, kWaitRecvBuf); });
            if (!err.empty()) fail(err);
        }
    }
}

void Session::run_strategies(const Workspace& w, const StratList& sl, bool monitored) {
    if (sl.empty()) fail("no strategy");
    const size_t bytes = w.bytes();
    if (bytes == 0) return;
    const size_t chunk = Config::get().chunk_bytes;
    const size_t k = (bytes + chunk - 1) / chunk;
    const size_t esz = dtype_size(w.dtype);
    // Chunks block on their peers (chunk i here waits for chunk i there) and fan out into nested parallel_for calls, so an
    // unbounded chunk count per op could exhaust any thread budget (a 1 GiB tensor = 1024 chunks, each holding a worker while
    // its recv/send tasks queue behind them). `kWindow` lanes pull chunk indices from a shared counter instead: chunks start in
    // increasing order on every rank, so the lowest unfinished chunk is always in flight everywhere and the op makes progress
    // with at most kWindow blocked workers (the reference relies on unbounded goroutines, session.go:304-317).
    constexpr size_t kWindow = 64;
    std::atomic<size_t> next{0};
    auto run_chunk = [&](size_t i) {
        auto be = even_partition(w.count, k, i);
        Workspace c = w;
        c.count = be.second - be.first;
        c.send = static_cast<const char*>(w.send) + be.first * esz;
        c.recv = static_cast<char*>(w.recv) + be.first * esz;
        // chunk i of the op, NOT its element range: two ranks that disagree on the tensor size must still rendezvous, so that the
        // receiver can report "chunk size mismatch" instead of both sides waiting forever for a name the other never uses
        c.name = "part::" + w.name + "#" + std::to_string(i);
        const Strat& s = sl[strategy_hash(i, w.name) % sl.size()];
        const Graph* gs[2] = {&s.g.reduce, &s.g.bcast};
        const double t0 = monitored ? now_s() : 0;
        run_graphs(c, gs, 2);
        if (monitored) s.stat->update(t0, now_s(), c.bytes());
    };
    std::string err = parallel_for((int)std::min(k, kWindow), [&](int) {
        for (size_t i = next.fetch_add(1); i < k; i = next.fetch_add(1)) run_chunk(i);
    });
    if (!err.empty()) fail(err);
}

void Session::all_reduce(const Workspace& w) {
    KF_TRACE_SCOPE("Session::all_reduce");
    StallGuard sg("AllReduce(" + w.name + ")");
    StratList sl;
    {
        std::lock_guard<std::mutex> g(mu_);
        sl = global_;
    }
    run_strategies(w, sl, false);
}

void Session::monitored_all_reduce(const Workspace& w, const std::vector<int32_t>& tree) {
    StallGuard sg("MonitoredAllReduce(" + w.name + ")");
    StratList sl;
    if (!tree.empty()) {
        if ((int)tree.size() != size()) fail("monitored_all_reduce: tree must have one entry per peer");
        std::vector<int> father(tree.begin(), tree.end());
        Graph bg;
        int roots = 0;
        if (!Graph::from_forest_array(father, bg, roots) || roots != 1) fail("monitored_all_reduce: not a rooted tree");
        std::lock_guard<std::mutex> g(mu_);
        // reuse the active stat object when the tree is the active strategy so that calc_stats sees the traffic
        if (global_.size() == 1 && global_[0].g.bcast.digest() == bg.digest()) sl = global_;
        else sl = wrap({GraphPair{default_reduce_graph(bg), bg}});
    } else {
        std::lock_guard<std::mutex> g(mu_);
        sl = global_;
    }
    run_strategies(w, sl, true);
}

// All-reduce inside every tree of a forest (father[i] == i marks a root; a root without children keeps its own data).
// Reference: `SubsetAllReduce` / `subset_all_reduce(t, topology)` (srcs/go/kungfu/session/allreduce.go, ops/collective.py).
// Peers outside every multi-node tree need not call it at all: no message is addressed to them.
void Session::subset_all_reduce(const Workspace& w, const std::vector<int32_t>& forest) {
    StallGuard sg("SubsetAllReduce(" + w.name + ")");
    if ((int)forest.size() != size()) fail("subset_all_reduce: forest must have one entry per peer");
    std::vector<int> father(forest.begin(), forest.end());
    Graph bg;
    int roots = 0;
    if (!Graph::from_forest_array(father, bg, roots) || roots < 1) fail("subset_all_reduce: not a forest");
    run_strategies(w, wrap({GraphPair{default_reduce_graph(bg), bg}}), false);
}

void Session::cross_all_reduce(const Workspace& w) {
    StallGuard sg("CrossAllReduce(" + w.name + ")");
    run_strategies(w, cross_, false);
}

void Session::reduce(const Workspace& w) {
    StallGuard sg("Reduce(" + w.name + ")");
    const Graph* gs[1] = {&global_[0].g.reduce};
    run_graphs(w, gs, 1);
}
void Session::broadcast(const Workspace& w) {
    KF_TRACE_SCOPE("Session::broadcast");
    StallGuard sg("Broadcast(" + w.name + ")");
    const Graph* gs[1] = {&global_[0].g.bcast};
    run_graphs(w, gs, 1);
}
void Session::local_reduce(const Workspace& w) {
    const Graph* gs[1] = {&local_[0].g.reduce};
    run_graphs(w, gs, 1);
}
void Session::local_broadcast(const Workspace& w) {
    const Graph* gs[1] = {&local_[0].g.bcast};
    run_graphs(w, gs, 1);
}

void Session::barrier() {
    KF_TRACE_SCOPE("Session::barrier");
    StallGuard sg("Barrier");
    const int n = size();
    std::vector<uint8_t> a(n, 0), b(n, 0);
    Workspace w;
    w.send = a.data();
    w.recv = b.data();
    w.count = (size_t)n;
    w.dtype = kU8;
    w.op = kSum;
    w.name = "kungfu::barrier::" + std::to_string(seq_++);
    all_reduce(w);
}

bool Session::consensus(const void* data, size_t len, const std::string& name) {
    StallGuard sg("Consensus(" + name + ")");
    const uint64_t id = seq_++;
    {
        int64_t n = (int64_t)len, lo = 0, hi = 0;
        Workspace w1{&n, &lo, 1, kI64, kMin, "kungfu::consensus::len::min::" + name + std::to_string(id)};
        Workspace w2{&n, &hi, 1, kI64, kMax, "kungfu::consensus::len::max::" + name + std::to_string(id)};
        all_reduce(w1);
        all_reduce(w2);
        if (lo != hi) return false;
    }
    if (len == 0) return true;
    std::vector<uint8_t> lo(len), hi(len);
    Workspace w1{data, lo.data(), len, kU8, kMin, "kungfu::consensus::min::" + name + std::to_string(id)};
    Workspace w2{data, hi.data(), len, kU8, kMax, "kungfu::consensus::max::" + name + std::to_string(id)};
    all_reduce(w1);
    all_reduce(w2);
    return lo == hi;
}

void Session::gather(const Workspace& w) {
    StallGuard sg("Gather(" + w.name + ")");
    const size_t bytes = w.bytes();
    if (rank_ != 0) {
        client_->send(peers_[0], w.name, w.send, bytes, kConnCollective, kWaitRecvBuf);
        return;
    }
    char* out = static_cast<char*>(w.recv);
    std::string err = parallel_for(size(), [&](int r) {
        if (r == rank_) {
            if (bytes) memcpy(out + (size_t)r * bytes, w.send, bytes);
        } else {
            coll_->recv_into(peers_[r], w.name, out + (size_t)r * bytes, bytes);
        }
    });
    if (!err.empty()) fail(err);
}

void Session::all_gather(const Workspace& w) {
    KF_TRACE_SCOPE("Session::all_gather");
    StallGuard sg("AllGather(" + w.name + ")");
    const size_t bytes = w.bytes();
    char* out = static_cast<char*>(w.recv);
    std::string e1, e2;
    // sends and receives proceed concurrently (direct all-to-all)
    std::string err = parallel_for(2, [&](int phase) {
        if (phase == 0) {
            e1 = parallel_for(size(), [&](int r) {
                if (r != rank_) client_->send(peers_[r], w.name, w.send, bytes, kConnCollective, kWaitRecvBuf);
            });
        } else {
            e2 = parallel_for(size(), [&](int r) {
                if (r == rank_) {
                    if (bytes) memcpy(out + (size_t)r * bytes, w.send, bytes);
                } else {
                    coll_->recv_into(peers_[r], w.name, out + (size_t)r * bytes, bytes);
                }
            });
        }
    });
    if (!err.empty()) fail(err);
    if (!e1.empty()) fail(e1);
    if (!e2.empty()) fail(e2);
}

void Session::set_tree(const std::vector<int32_t>& father_i32) {
    if ((int)father_i32.size() != size()) fail("set_tree: need one father per peer");
    std::vector<int> father(father_i32.begin(), father_i32.end());
    Graph bg;
    int roots = 0;
    if (!Graph::from_forest_array(father, bg, roots) || roots != 1) fail("set_tree: not a rooted tree");
    StratList sl = wrap({GraphPair{default_reduce_graph(bg), bg}});
    barrier();
    const std::string d = bg.digest();
    if (!consensus(d.data(), d.size(), "kungfu::SetStrategy")) fail("set_tree: peers disagree on the new strategy");
    {
        std::lock_guard<std::mutex> g(mu_);
        global_ = sl;
    }
    barrier();
}

std::vector<double> Session::peer_latencies() {
    std::vector<double> out(size(), 0.0);
    parallel_for(size(), [&](int r) {
        if (r != rank_) out[r] = client_->ping(peers_[r]);
    });
    return out;
}

std::vector<double> Session::egress_rates() { return NetMonitor::get().egress_rates(peers_); }

// ---- strategy statistics / interference detection ---------------------------------------------------
void Session::calc_stats() {
    std::lock_guard<std::mutex> g(mu_);
    if (global_.size() != 1) {
        KF_LOGE("calc_stats should only be called with one active communication strategy");
        return;
    }
    StrategyStat& st = *global_[0].stat;
    std::lock_guard<std::mutex> g2(st.mu);
    if (st.acc_size == 0) return;
    const double dt = st.last_end - st.first_begin;
    st.throughput = dt > 0 ? (double)st.acc_size / dt : 0.0;
    st.reset();
}
void Session::log_stats() {
    std::lock_guard<std::mutex> g(mu_);
    if (global_.size() != 1) {
        KF_LOGE("log_stats should only be called with one active communication strategy");
        return;
    }
    stat_log_.push_back(global_[0].stat->throughput);
}
void Session::print_strategy_stats() {
    std::lock_guard<std::mutex> g(mu_);
    KF_LOGI("available strategies: %d", (int)global_.size());
    for (size_t i = 0; i < stat_log_.size(); ++i) KF_LOGI("global step #%d, master[0], throughput=%.3f MiB/s", (int)i, stat_log_[i] / (1 << 20));
}
double Session::last_throughput() {
    std::lock_guard<std::mutex> g(mu_);
    return global_.empty() ? 0.0 : global_[0].stat->throughput;
}
bool Session::check_interference() {
    double thr = 0, ref = 0;
    {
        std::lock_guard<std::mutex> g(mu_);
        if (global_.size() != 1) {
            KF_LOGE("check_interference should only be called with one active communication strategy");
            return false;
        }
        StrategyStat& st = *global_[0].stat;
        if (st.ref_throughput == 0) {
            st.ref_throughput = st.throughput; // first call: take the reference window
            return false;
        }
        thr = st.throughput;
        ref = st.ref_throughput;
    }
    int32_t vote = thr < 0.8 * ref ? 1 : 0, total = 0; // int32: an int8 sum wraps beyond 127 peers
    Workspace w{&vote, &total, 1, kI32, kSum, "StratMon::" + std::to_string(seq_++)};
    all_reduce(w);
    return total > (int32_t)(size() / 2);
}

} // namespace kfrt
