// kungfu_b200 runtime — Peer: the process-level singleton of a worker.
//
// Reference: srcs/go/kungfu/peer/{peer.go:27-276, legacy.go:18-39, p2p.go:15-41, router.go:14-73} and the env bootstrap
// srcs/go/kungfu/env/config.go:24-80. Owns the server, the client, the endpoints, the current Session and the
// current Cluster + version. Elastic protocol (`resize_cluster_from_url`): fetch the desired cluster from the config
// server until all peers agree on it (consensus on the digest), notify every runner with Stage{version+1, cluster}, bump
// the version; survivors rebuild their session (new token = new version, fresh connections, barrier with the NEW
// membership), detached peers stop themselves.
#pragma once
#include <atomic>
#include <condition_variable>
#include <deque>
#include <functional>
#include <memory>
#include <mutex>
#include <string>

#include "plan.hpp"
#include "session.hpp"
#include "transport.hpp"

namespace kfrt {

struct PeerConfig {
    PeerID self;
    PeerID parent;
    PeerList init_peers;
    PeerList init_runners;
    Strategy strategy = Strategy::Auto;
    int init_cluster_version = 0;
    std::string config_server;
    bool single = false;
    std::string launcher = "single"; // kungfu-run | torchrun | single | explicit

    static PeerConfig from_env();                         // KUNGFU_* env, else torchrun env, else single process
    static PeerConfig single_machine(int rank, int size, uint16_t base_port = 0); // in-process multi-worker (tests)
};

class Peer {
public:
    explicit Peer(const PeerConfig& cfg);
    ~Peer();
    void start();
    void close();

    std::shared_ptr<Session> session();
    bool update();

    uint64_t uid() const;
    bool detached() const { return detached_; }
    bool single() const { return cfg_.single; }
    int cluster_version() const { return cluster_version_.load(); }
    const PeerConfig& config() const { return cfg_; }
    Cluster current_cluster();
    PeerID self() const { return cfg_.self; }

    // elastic control; each returns (changed, detached)
    std::pair<bool, bool> resize_cluster(int new_size);
    std::pair<bool, bool> resize_cluster_from_url();
    void propose_new_size(int new_size);      // rank 0 PUTs the resized cluster to the config server
    std::pair<bool, bool> propose(const Cluster& c);
    // Shrink-and-continue after a worker died: the runner that saw the process go sends every survivor a "lost" control
    // message carrying Stage{version + 1, cluster without the dead worker}; it aborts the collectives in flight (they throw)
    // and is kept here. `recover()` waits for that stage (at most timeout_s), adopts it WITHOUT a consensus round on the old
    // session (which the dead peer can no longer answer) and meets the other survivors at the barrier of the new session.
    // Returns (changed, detached). The reference only restarts the whole job from a checkpoint (runner/monitored.go:48-73).
    std::pair<bool, bool> recover(double timeout_s);
    bool failure_pending();
    // called (on the server's reader thread) the moment a loss is announced: the GPU plane uses it to abort device-side waits
    void set_lost_callback(std::function<void()> f);

    // p2p blob store
    void save(const std::string& name, const void* data, size_t len);
    void save_version(const std::string& version, const std::string& name, const void* data, size_t len);
    bool request(int rank, const std::string& version, const std::string& name, void* buf, size_t len);

    Client& client() { return *client_; }
    CollectiveEndpoint& collective() { return coll_; }

private:
    bool update_to(const PeerList& pl);
    bool fetch_cluster(Cluster& out);
    bool trusted_source(const PeerID& src);
    void on_control(const PeerID& src, const std::string& name, const Bytes& payload);

    PeerConfig cfg_;
    std::mutex mu_;
    std::unique_ptr<Client> client_;
    CollectiveEndpoint coll_;
    std::unique_ptr<P2PEndpoint> p2p_;
    std::unique_ptr<Server> server_;
    std::unique_ptr<class HttpServer> metrics_;
    std::shared_ptr<Session> session_;
    Cluster cluster_;
    // Announced by the runners, not yet adopted. Own mutex: `mu_` is held for the whole of update_to() — including its barrier, which is
    // exactly what a further loss must be able to abort — so the handler of a `lost` message may not wait for it.
    std::mutex lost_mu_;
    std::condition_variable lost_cv_;
    std::deque<Stage> lost_stages_;
    std::function<void()> on_lost_;
    std::atomic<int> rank_hint_{-1}; // rank in the current session, readable without the lock (watchdog thread)
    bool reports_waiting_ = false;
    std::atomic<int> cluster_version_{0}; // written under mu_; read without it by the loss handler (which must never wait for mu_)
    bool updated_ = false;
    bool detached_ = false;
    bool started_ = false;
};

} // namespace kfrt
