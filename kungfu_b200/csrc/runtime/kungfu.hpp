// kungfu.hpp — header-only C++ API for applications: `kungfu::Peer` over the runtime's C ABI (capi.h, lib/libkungfu-comm.so).
//
//   g++ -std=c++17 -I kungfu_b200/csrc/runtime app.cpp -L kungfu_b200/lib -lkungfu-comm -Wl,-rpath,$PWD/kungfu_b200/lib -pthread
//
// Reference: srcs/cpp/include/kungfu/peer.hpp:14-167 (`kungfu::Peer`: metadata, Save / Request, Barrier, Consensus, Reduce,
// AllReduce, CrossAllReduce, MonitoredAllReduce, AllGather, Broadcast, LocalBroadcast, Gather, AllGatherTransform, SetTree,
// GetPeerLatencies, CheckInterference, GetEgressRates, ResizeCluster[FromURL], ProposeNewSize, stats), dtype.hpp (`type_encoder`),
// mst.hpp (`MinimumSpanningTree`). The library's own classes (kfrt::Peer, kfrt::Session) are not exported from the shared object; this
// header is the supported C++ surface. Every call returns 0 on success; `kungfu::last_error()` has the text otherwise. A call with a
// `DoneCallback` returns at once and runs the callback on a runtime thread when the operation has completed.
#pragma once
#include <cstdint>
#include <functional>
#include <limits>
#include <string>
#include <type_traits>
#include <vector>

#include "capi.h"

namespace kungfu {

using DoneCallback = std::function<void()>;

inline std::string last_error() { return kungfu_last_error(); }

// C++ type -> dtype code of the wire protocol
template <typename T> struct type_encoder;
#define KF_ENCODE(T, code)                                  \
    template <> struct type_encoder<T> {                    \
        static constexpr int value() { return code; }       \
    }
KF_ENCODE(float, KungFu_F32);
KF_ENCODE(double, KungFu_F64);
KF_ENCODE(int8_t, KungFu_I8);
KF_ENCODE(uint8_t, KungFu_U8);
KF_ENCODE(int16_t, KungFu_I16);
KF_ENCODE(uint16_t, KungFu_U16);
KF_ENCODE(int32_t, KungFu_I32);
KF_ENCODE(uint32_t, KungFu_U32);
KF_ENCODE(int64_t, KungFu_I64);
KF_ENCODE(uint64_t, KungFu_U64);
#undef KF_ENCODE

inline size_t dtype_size(int dtype) {
    switch (dtype) {
    case KungFu_U8: case KungFu_I8: return 1;
    case KungFu_BF16: case KungFu_F16: case KungFu_I16: case KungFu_U16: return 2;
    case KungFu_F32: case KungFu_I32: case KungFu_U32: return 4;
    default: return 8;
    }
}

namespace detail {
// a std::function travels through the C ABI as a heap cell that the trampoline frees after the call
inline void trampoline(void* p) {
    auto* f = static_cast<DoneCallback*>(p);
    (*f)();
    delete f;
}
inline void* box(const DoneCallback& done) { return new DoneCallback(done); }
// operations whose C entry point has no callback form complete inline; the callback then rides on a no-op so that it still arrives on a
// runtime thread, after the result is in place
inline int then(int rc, const DoneCallback& done) {
    if (rc != 0) return rc;
    return kungfu_noop(trampoline, box(done));
}
} // namespace detail

class Peer {
public:
    Peer() : ok_(kungfu_init() == 0) {}                                                   // from the launcher's environment
    Peer(int rank, int size, int base_port = 0) : ok_(kungfu_init_single_machine(rank, size, base_port) == 0) {}
    ~Peer() { kungfu_finalize(); }
    Peer(const Peer&) = delete;
    Peer& operator=(const Peer&) = delete;
    bool Ok() const { return ok_; }

    // metadata
    bool Detached() const { return kungfu_detached() != 0; }
    uint64_t Uid() const { return kungfu_uid(); }
    int Rank() const { return kungfu_rank(); }
    int Size() const { return kungfu_size(); }
    int LocalRank() const { return kungfu_local_rank(); }
    int LocalSize() const { return kungfu_local_size(); }
    int HostCount() const { return kungfu_host_count(); }
    int ClusterVersion() const { return kungfu_cluster_version(); }
    int Noop(const DoneCallback& done) { return kungfu_noop(detail::trampoline, detail::box(done)); }

    // blob store and peer-to-peer requests
    int Save(const char* name, const void* buf, int count, int dtype) { return kungfu_save(name, buf, count * dtype_size(dtype)); }
    int Save(const char* version, const char* name, const void* buf, int count, int dtype) {
        return kungfu_save_version(version, name, buf, count * dtype_size(dtype));
    }
    int Save(const char* name, const void* buf, int count, int dtype, const DoneCallback& done) { return detail::then(Save(name, buf, count, dtype), done); }
    // `*ok` (optional) tells whether the target had the blob; a missing blob is not an error
    int Request(int rank, const char* name, void* buf, int count, int dtype, bool* ok = nullptr) { return Request(rank, "", name, buf, count, dtype, ok); }
    int Request(int rank, const char* version, const char* name, void* buf, int count, int dtype, bool* ok = nullptr) {
        int found = 0;
        const int rc = kungfu_request(rank, version, name, buf, count * dtype_size(dtype), &found);
        if (ok) *ok = found != 0;
        return rc;
    }

    // collectives
    int Barrier() { return kungfu_barrier(); }
    int Barrier(const DoneCallback& done) { return detail::then(kungfu_barrier(), done); }
    int Consensus(const void* buf, int count, int dtype, bool* ok, const char* name) {
        int agreed = 0;
        const int rc = kungfu_consensus(buf, count * dtype_size(dtype), name, &agreed);
        *ok = agreed != 0;
        return rc;
    }
    int Reduce(const void* send, void* recv, int count, int dtype, int op, const char* name) { return kungfu_reduce(send, recv, count, dtype, op, name); }
    int AllReduce(const void* send, void* recv, int count, int dtype, int op, const char* name) {
        return kungfu_all_reduce(send, recv, count, dtype, op, name, nullptr, nullptr);
    }
    int AllReduce(const void* send, void* recv, int count, int dtype, int op, const char* name, const DoneCallback& done) {
        return kungfu_all_reduce(send, recv, count, dtype, op, name, detail::trampoline, detail::box(done));
    }
    int CrossAllReduce(const void* send, void* recv, int count, int dtype, int op, const char* name) {
        return kungfu_cross_all_reduce(send, recv, count, dtype, op, name);
    }
    int MonitoredAllReduce(const void* send, void* recv, int count, int dtype, int op, const int32_t* tree, const char* name) {
        return kungfu_monitored_all_reduce(send, recv, count, dtype, op, tree, tree ? Size() : 0, name);
    }
    int SubsetAllReduce(const void* send, void* recv, int count, int dtype, int op, const int32_t* forest, const char* name) {
        return kungfu_subset_all_reduce(send, recv, count, dtype, op, forest, Size(), name);
    }
    int AllGather(const void* send, int count, int dtype, void* recv, const char* name) { return kungfu_all_gather(send, recv, count, dtype, name); }
    int Gather(const void* send, int count, int dtype, void* recv, const char* name) { return kungfu_gather(send, recv, count, dtype, name); }
    int Broadcast(const void* send, void* recv, int count, int dtype, const char* name) {
        return kungfu_broadcast(send, recv, count, dtype, name, nullptr, nullptr);
    }
    int Broadcast(const void* send, void* recv, int count, int dtype, const char* name, const DoneCallback& done) {
        return kungfu_broadcast(send, recv, count, dtype, name, detail::trampoline, detail::box(done));
    }
    int LocalReduce(const void* send, void* recv, int count, int dtype, int op, const char* name) { return kungfu_local_reduce(send, recv, count, dtype, op, name); }
    int LocalBroadcast(const void* send, void* recv, int count, int dtype, const char* name) { return kungfu_local_broadcast(send, recv, count, dtype, name); }

    // typed forms
    template <typename T> int AllReduce(const T* send, T* recv, int count, int op, const char* name) {
        return AllReduce(send, recv, count, type_encoder<T>::value(), op, name);
    }
    template <typename T> int AllReduce(const T* send, T* recv, int count, int op, const char* name, const DoneCallback& done) {
        return AllReduce(send, recv, count, type_encoder<T>::value(), op, name, done);
    }
    template <typename T> int Broadcast(const T* send, T* recv, int count, const char* name) { return Broadcast(send, recv, count, type_encoder<T>::value(), name); }
    template <typename T> int AllGather(const T* send, int count, T* recv, const char* name) { return AllGather(send, count, type_encoder<T>::value(), recv, name); }

    // Everybody contributes `input_count` elements; `f(all inputs in rank order, Size() * input_count, output, output_count)` is evaluated
    // on rank 0 and its output is what every rank receives — ONE evaluation, so the result is the same everywhere even if `f` is not
    // deterministic (used for: measured latencies -> spanning tree -> SetTree).
    template <typename T1, typename T2, typename F>
    int AllGatherTransform(const T1* input, int input_count, T2* output, int output_count, const char* name, const F& f) {
        std::vector<T1> all(Rank() == 0 ? (size_t)Size() * input_count : 0);
        const std::string n(name);
        if (int rc = Gather(input, input_count, type_encoder<T1>::value(), all.data(), (n + "::gather").c_str())) return rc;
        if (Rank() == 0) f(all.data(), Size() * input_count, output, output_count);
        return Broadcast(output, output, output_count, (n + "::bcast").c_str());
    }

    // adaptation and monitoring
    int SetTree(const int32_t* tree) { return kungfu_set_tree(tree, Size()); }
    int GetPeerLatencies(float* out, int n) { return kungfu_get_peer_latencies(out, n); }
    int GetEgressRates(float* out) { return kungfu_get_egress_rates(out, Size()); }
    int CheckInterference() {
        int r = 0;
        return kungfu_check_interference(&r) == 0 ? r : -1;
    }
    void CalcStats() { kungfu_calc_stats(); }
    void LogStats() { kungfu_log_stats(); }
    void PrintStategyStats() { kungfu_print_strategy_stats(); } // spelled as in the reference

    // elastic control
    int ResizeCluster(uint32_t new_size, bool* changed, bool* detached) {
        int c = 0, d = 0;
        const int rc = kungfu_resize((int)new_size, &c, &d);
        *changed = c != 0, *detached = d != 0;
        return rc;
    }
    int ResizeClusterFromURL(bool* changed, bool* detached) {
        int c = 0, d = 0;
        const int rc = kungfu_resize_from_url(&c, &d);
        *changed = c != 0, *detached = d != 0;
        return rc;
    }
    int ProposeNewSize(int new_size) { return kungfu_propose_new_size(new_size); }
    // shrink-and-continue: after a collective failed because a worker died, adopt the cluster the launcher announced
    int Recover(double timeout_s, bool* changed, bool* detached) {
        int c = 0, d = 0;
        const int rc = kungfu_recover(timeout_s, &c, &d);
        *changed = c != 0, *detached = d != 0;
        return rc;
    }
    bool FailurePending() {
        int p = 0;
        return kungfu_failure_pending(&p) == 0 && p != 0;
    }

private:
    bool ok_;
};

// Minimum spanning tree of a dense, possibly asymmetric weight matrix (the cost of an edge is the mean of its two directions), as a
// father array rooted at `root`: father[root] == root. This is the form SetTree takes.
template <typename Weight> std::vector<int32_t> minimum_spanning_tree(int n, const Weight* weights, int root = 0) {
    std::vector<int32_t> father(n, root);
    std::vector<char> in_tree(n, 0);
    std::vector<Weight> dist(n, std::numeric_limits<Weight>::max());
    auto cost = [&](int a, int b) { return (weights[a * n + b] + weights[b * n + a]) / 2; };
    dist[root] = Weight();
    for (int round = 0; round < n; ++round) {
        int next = -1;
        for (int v = 0; v < n; ++v)
            if (!in_tree[v] && (next < 0 || dist[v] < dist[next])) next = v;
        in_tree[next] = 1;
        for (int v = 0; v < n; ++v)
            if (!in_tree[v] && cost(next, v) < dist[v]) dist[v] = cost(next, v), father[v] = next;
    }
    father[root] = root;
    return father;
}

} // namespace kungfu
