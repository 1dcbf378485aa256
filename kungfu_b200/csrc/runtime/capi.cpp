// Implementation of capi.h (reference mapping there: libkungfu-comm cgo exports, srcs/go/libkungfu-comm/*.go).
#include "capi.h"

#include <memory>
#include <mutex>
#include <string>

#include "peer.hpp"

using namespace kfrt;

namespace {
std::unique_ptr<Peer> g_peer;
std::mutex g_mu;
thread_local std::string g_err;

template <typename F> int guard(F&& f) {
    try {
        f();
        return 0;
    } catch (const std::exception& e) {
        g_err = e.what();
        KF_LOGE("%s", e.what());
        return 1;
    }
}
Peer& peer() {
    if (!g_peer) fail("kungfu runtime is not initialised");
    return *g_peer;
}
Workspace mk(const void* s, void* r, size_t n, int dt, int op, const char* name) {
    Workspace w;
    w.send = s;
    w.recv = r;
    w.count = n;
    w.dtype = dt;
    w.op = op;
    w.name = name ? name : "";
    return w;
}
template <typename F> int run(F fn, kungfu_callback_t done, void* arg) {
    if (!done) return guard(fn);
    Pool::get().submit([fn, done, arg] {
        guard(fn);
        done(arg);
    });
    return 0;
}
} // namespace

extern "C" {

#define KF_API __attribute__((visibility("default")))

KF_API int kungfu_init(void) {
    return guard([] {
        std::lock_guard<std::mutex> g(g_mu);
        if (g_peer) return;
        g_peer.reset(new Peer(PeerConfig::from_env()));
        g_peer->start();
    });
}
KF_API int kungfu_init_single_machine(int rank, int size, int base_port) {
    return guard([&] {
        std::lock_guard<std::mutex> g(g_mu);
        g_peer.reset(new Peer(PeerConfig::single_machine(rank, size, (uint16_t)base_port)));
        g_peer->start();
    });
}
KF_API int kungfu_finalize(void) {
    return guard([] {
        std::lock_guard<std::mutex> g(g_mu);
        if (g_peer) g_peer->close();
        g_peer.reset();
    });
}
KF_API const char* kungfu_last_error(void) { return g_err.c_str(); }

KF_API int kungfu_rank(void) { return g_peer ? g_peer->session()->rank() : 0; }
KF_API int kungfu_size(void) { return g_peer ? g_peer->session()->size() : 1; }
KF_API int kungfu_local_rank(void) { return g_peer ? g_peer->session()->local_rank() : 0; }
KF_API int kungfu_local_size(void) { return g_peer ? g_peer->session()->local_size() : 1; }
KF_API int kungfu_host_count(void) { return g_peer ? g_peer->session()->host_count() : 1; }
KF_API uint64_t kungfu_uid(void) { return g_peer ? g_peer->uid() : 0; }
KF_API int kungfu_detached(void) { return g_peer && g_peer->detached() ? 1 : 0; }
KF_API int kungfu_cluster_version(void) { return g_peer ? g_peer->cluster_version() : 0; }

KF_API int kungfu_barrier(void) {
    return guard([] { peer().session()->barrier(); });
}
KF_API int kungfu_consensus(const void* buf, size_t len, const char* name, int* agreed) {
    return guard([&] { *agreed = peer().session()->consensus(buf, len, name ? name : "") ? 1 : 0; });
}
KF_API int kungfu_all_reduce(const void* s, void* r, size_t n, int dt, int op, const char* name, kungfu_callback_t done, void* arg) {
    auto sess = peer().session();
    Workspace w = mk(s, r, n, dt, op, name);
    return run([sess, w] { sess->all_reduce(w); }, done, arg);
}
KF_API int kungfu_monitored_all_reduce(const void* s, void* r, size_t n, int dt, int op, const int32_t* tree, int tree_len, const char* name) {
    return guard([&] { peer().session()->monitored_all_reduce(mk(s, r, n, dt, op, name), std::vector<int32_t>(tree, tree + (tree ? tree_len : 0))); });
}
KF_API int kungfu_subset_all_reduce(const void* s, void* r, size_t n, int dt, int op, const int32_t* forest, int forest_len, const char* name) {
    return guard([&] { peer().session()->subset_all_reduce(mk(s, r, n, dt, op, name), std::vector<int32_t>(forest, forest + (forest ? forest_len : 0))); });
}
KF_API int kungfu_noop(kungfu_callback_t done, void* arg) {
    return run([] {}, done, arg); // completes on the runtime's executor: orders a callback after everything queued before it
}
KF_API int kungfu_cross_all_reduce(const void* s, void* r, size_t n, int dt, int op, const char* name) {
    return guard([&] { peer().session()->cross_all_reduce(mk(s, r, n, dt, op, name)); });
}
KF_API int kungfu_reduce(const void* s, void* r, size_t n, int dt, int op, const char* name) {
    return guard([&] { peer().session()->reduce(mk(s, r, n, dt, op, name)); });
}
KF_API int kungfu_broadcast(const void* s, void* r, size_t n, int dt, const char* name, kungfu_callback_t done, void* arg) {
    auto sess = peer().session();
    Workspace w = mk(s, r, n, dt, kSum, name);
    return run([sess, w] { sess->broadcast(w); }, done, arg);
}
KF_API int kungfu_local_reduce(const void* s, void* r, size_t n, int dt, int op, const char* name) {
    return guard([&] { peer().session()->local_reduce(mk(s, r, n, dt, op, name)); });
}
KF_API int kungfu_local_broadcast(const void* s, void* r, size_t n, int dt, const char* name) {
    return guard([&] { peer().session()->local_broadcast(mk(s, r, n, dt, kSum, name)); });
}
KF_API int kungfu_gather(const void* s, void* r, size_t n, int dt, const char* name) {
    return guard([&] { peer().session()->gather(mk(s, r, n, dt, kSum, name)); });
}
KF_API int kungfu_all_gather(const void* s, void* r, size_t n, int dt, const char* name) {
    return guard([&] { peer().session()->all_gather(mk(s, r, n, dt, kSum, name)); });
}
KF_API int kungfu_all_reduce_int_max(int64_t* x) {
    return guard([&] {
        auto sess = peer().session();
        int64_t y = *x;
        sess->all_reduce(mk(x, &y, 1, kI64, kMax, ("kungfu::int_max::" + std::to_string(sess->next_seq())).c_str()));
        *x = y;
    });
}

KF_API int kungfu_save(const char* name, const void* buf, size_t len) {
    return guard([&] { peer().save(name, buf, len); });
}
KF_API int kungfu_save_version(const char* version, const char* name, const void* buf, size_t len) {
    return guard([&] { peer().save_version(version, name, buf, len); });
}
KF_API int kungfu_request(int rank, const char* version, const char* name, void* buf, size_t len, int* ok) {
    return guard([&] { *ok = peer().request(rank, version ? version : "", name, buf, len) ? 1 : 0; });
}

KF_API int kungfu_resize(int new_size, int* changed, int* detached) {
    return guard([&] {
        auto r = peer().resize_cluster(new_size);
        *changed = r.first;
        *detached = r.second;
    });
}
KF_API int kungfu_resize_from_url(int* changed, int* detached) {
    return guard([&] {
        auto r = peer().resize_cluster_from_url();
        *changed = r.first;
        *detached = r.second;
    });
}
KF_API int kungfu_recover(double timeout_s, int* changed, int* detached) {
    return guard([&] {
        auto r = peer().recover(timeout_s);
        *changed = r.first;
        *detached = r.second;
    });
}
KF_API int kungfu_failure_pending(int* pending) {
    return guard([&] { *pending = peer().failure_pending() ? 1 : 0; });
}
KF_API int kungfu_propose_new_size(int new_size) {
    return guard([&] { peer().propose_new_size(new_size); });
}
KF_API int kungfu_set_tree(const int32_t* father, int n) {
    return guard([&] { peer().session()->set_tree(std::vector<int32_t>(father, father + n)); });
}
KF_API int kungfu_get_peer_latencies(float* out, int n) {
    return guard([&] {
        auto v = peer().session()->peer_latencies();
        for (int i = 0; i < n && i < (int)v.size(); ++i) out[i] = (float)v[i];
    });
}
KF_API int kungfu_get_egress_rates(float* out, int n) {
    return guard([&] {
        auto v = peer().session()->egress_rates();
        for (int i = 0; i < n && i < (int)v.size(); ++i) out[i] = (float)v[i];
    });
}
KF_API int kungfu_check_interference(int* result) {
    return guard([&] { *result = peer().session()->check_interference() ? 1 : 0; });
}
KF_API int kungfu_calc_stats(void) {
    return guard([] { peer().session()->calc_stats(); });
}
KF_API int kungfu_log_stats(void) {
    return guard([] { peer().session()->log_stats(); });
}
KF_API int kungfu_print_strategy_stats(void) {
    return guard([] { peer().session()->print_strategy_stats(); });
}

} // extern "C"
