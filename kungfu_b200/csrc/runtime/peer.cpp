#include "peer.hpp"

#include <unistd.h>

#include <cstdlib>

#include "http.hpp"

namespace kfrt {

// ---- bootstrap --------------------------------------------------------------------------------------
PeerConfig PeerConfig::from_env() {
    PeerConfig c;
    const std::string self = getenv_str(envkeys::kSelfSpec);
    if (!self.empty()) { // launched by kungfu-run
        c.launcher = "kungfu-run";
        c.self = PeerID::parse(self);
        c.init_peers = PeerList::parse(getenv_str(envkeys::kInitPeers));
        c.init_runners = PeerList::parse(getenv_str(envkeys::kInitRunners));
        const std::string parent = getenv_str(envkeys::kParentID);
        if (!parent.empty()) c.parent = PeerID::parse(parent);
        c.init_cluster_version = atoi(getenv_str(envkeys::kInitClusterVersion, "0").c_str());
        c.strategy = parse_strategy(getenv_str(envkeys::kAllReduceStrategy, "AUTO"));
        c.config_server = getenv_str(envkeys::kConfigServer);
        return c;
    }
    const std::string rank = getenv_str("RANK"), world = getenv_str("WORLD_SIZE");
    if (!rank.empty() && !world.empty() && atoi(world.c_str()) > 1) { // torchrun, one node
        c.launcher = "torchrun";
        const int r = atoi(rank.c_str()), n = atoi(world.c_str());
        const uint32_t ip = resolve_ipv4(getenv_str("MASTER_ADDR", "127.0.0.1"));
        const int master_port = atoi(getenv_str("MASTER_PORT", "29500").c_str());
        int base = atoi(getenv_str("KUNGFU_PORT_BASE", "0").c_str());
        // stay below the kernel's ephemeral port range (32768+): listening up there collides now and then with the source
        // port of somebody's outgoing connection
        if (base <= 0) base = 10000 + (master_port * 7 + 131) % 20000;
        for (int i = 0; i < n; ++i) c.init_peers.push_back(PeerID{ip, (uint16_t)(base + i)});
        c.self = c.init_peers[r];
        c.strategy = parse_strategy(getenv_str(envkeys::kAllReduceStrategy, "AUTO"));
        c.config_server = getenv_str(envkeys::kConfigServer);
        return c;
    }
    c.single = true;
    c.self = PeerID{parse_ipv4("127.0.0.1"), kDefaultPortBegin};
    c.init_peers.push_back(c.self);
    return c;
}

PeerConfig PeerConfig::single_machine(int rank, int size, uint16_t base_port) {
    PeerConfig c;
    c.launcher = "explicit";
    const uint32_t ip = parse_ipv4("127.0.0.1");
    if (base_port == 0) base_port = kDefaultPortBegin;
    for (int i = 0; i < size; ++i) c.init_peers.push_back(PeerID{ip, (uint16_t)(base_port + i)});
    c.self = c.init_peers.at(rank);
    c.single = size == 1;
    c.strategy = parse_strategy(getenv_str(envkeys::kAllReduceStrategy, "AUTO"));
    return c;
}

// ---- lifecycle --------------------------------------------------------------------------------------
Peer::Peer(const PeerConfig& cfg) : cfg_(cfg) {
    client_.reset(new Client(cfg_.self));
    p2p_.reset(new P2PEndpoint(client_.get(), cfg_.self));
    server_.reset(new Server(cfg_.self, &coll_, p2p_.get(), Config::get().use_unix_sock));
    cluster_.runners = cfg_.init_runners;
    cluster_.workers = cfg_.init_peers;
    cluster_version_ = cfg_.init_cluster_version;
}

Peer::~Peer() {
    try {
        close();
    } catch (...) {
    }
}

void Peer::start() {
    if (started_) return;
    started_ = true;
    if (!cfg_.single) {
        server_->set_control_handler([this](const PeerID& src, const std::string& name, const Bytes& payload) { on_control(src, name, payload); });
        server_->start();
        if (Config::get().enable_monitoring) {
            metrics_.reset(new HttpServer());
            try {
                const std::string bind_to = (cfg_.self.ipv4 >> 24) == 127 ? format_ipv4(cfg_.self.ipv4) : std::string("0.0.0.0");
                const int port = metrics_->start(bind_to, cfg_.self.port + 10000, [](const HttpRequest& r) {
                    HttpResponse resp;
                    if (r.path == "/metrics") resp.body = NetMonitor::get().metrics_text();
                    else resp.status = 404;
                    return resp;
                });
                KF_LOGI("peer %s started, monitoring endpoint http://%s:%d/metrics", cfg_.self.str().c_str(), format_ipv4(cfg_.self.ipv4).c_str(), port);
            } catch (const std::exception& e) {
                KF_LOGW("metrics endpoint unavailable: %s", e.what());
                metrics_.reset();
            }
        }
    }
    // under a launcher that watches for hung workers: tell it whenever this worker is WAITING inside a collective, so that the one that
    // keeps everybody waiting can be told apart from its victims
    const std::string hw = getenv_str("KUNGFU_HANG_WATCH", "");
    if (!hw.empty() && !cfg_.single) {
        reports_waiting_ = true;
        set_stall_reporter([this, hw](double) {
            const int r = rank_hint_.load();
            if (r >= 0) http_request("POST", hw, "{\"key\":\"waiting:" + std::to_string(r) + "\"}", nullptr);
        });
    }
    update();
}

void Peer::close() {
    if (!started_) return;
    started_ = false;
    if (reports_waiting_) {
        set_stall_reporter(nullptr);
        reports_waiting_ = false;
    }
    coll_.shutdown("peer closed");
    p2p_->shutdown();
    if (metrics_) metrics_->stop();
    if (!cfg_.single) server_->stop();
    client_->close_all();
}

void Peer::on_control(const PeerID& src, const std::string& name, const Bytes& payload) {
    if (name == "exit") { // fault injection / orderly removal (reference: rchannel/handler/control.go:17-21)
        if (!trusted_source(src) && !getenv_bool("KUNGFU_ALLOW_REMOTE_EXIT", false)) {
            KF_LOGW("ignoring exit request from %s: not a host of this job", src.str().c_str());
            return;
        }
        KF_LOGW("exit requested by %s", src.str().c_str());
        _exit(0);
    }
    if (name == "lost") { // a runner saw one of the workers die: abort what is in flight, keep the stage for recover()
        if (!trusted_source(src)) {
            KF_LOGW("ignoring membership notice from %s: not a host of this job", src.str().c_str());
            return;
        }
        Stage st;
        try {
            st = Stage::from_json(std::string(payload.begin(), payload.end()));
        } catch (const std::exception& e) {
            KF_LOGE("bad stage in a lost-peer notice from %s: %s", src.str().c_str(), e.what());
            return;
        }
        {
            std::lock_guard<std::mutex> g(lost_mu_); // NOT mu_: a regroup in progress holds that one, and this message must abort it
            if (st.version <= cluster_version_.load()) return; // old news
            for (auto& q : lost_stages_)
                if (q.version == st.version) return; // every runner announces it
            lost_stages_.push_back(st);
            KF_LOGW("peer lost (announced by %s): v%d has %d workers: %s", src.str().c_str(), st.version, (int)st.cluster.workers.size(),
                    st.cluster.workers.str().c_str());
        }
        coll_.shutdown("a peer of this cluster was lost");
        client_->interrupt();
        NetMonitor::get().add_counter("kungfu_peer_failures_total", 1);
        lost_cv_.notify_all();
        std::function<void()> cb;
        {
            std::lock_guard<std::mutex> g(lost_mu_);
            cb = on_lost_;
        }
        if (cb) {
            try {
                cb();
            } catch (const std::exception& e) {
                KF_LOGW("lost-peer callback: %s", e.what());
            }
        }
    }
}

void Peer::set_lost_callback(std::function<void()> f) {
    std::lock_guard<std::mutex> g(lost_mu_);
    on_lost_ = std::move(f);
}

bool Peer::trusted_source(const PeerID& src) {
    // only hosts that take part in the job (workers, runners) or this machine may change a worker's fate
    if ((src.ipv4 >> 24) == 127 || src.ipv4 == cfg_.self.ipv4) return true;
    for (auto& p : cfg_.init_peers)
        if (p.ipv4 == src.ipv4) return true;
    for (auto& p : cfg_.init_runners)
        if (p.ipv4 == src.ipv4) return true;
    // hosts that joined later: the current cluster — but never wait for a regroup in progress (it holds mu_ and may be what this
    // message is about to abort)
    std::unique_lock<std::mutex> g(mu_, std::try_to_lock);
    if (!g.owns_lock()) return false;
    for (auto& p : cluster_.workers)
        if (p.ipv4 == src.ipv4) return true;
    for (auto& p : cluster_.runners)
        if (p.ipv4 == src.ipv4) return true;
    return false;
}

bool Peer::failure_pending() {
    std::lock_guard<std::mutex> g(lost_mu_);
    return !lost_stages_.empty();
}

std::pair<bool, bool> Peer::recover(double timeout_s) {
    StallGuard sg("recover");
    const auto wait_ms = std::chrono::milliseconds((long)(timeout_s * 1000));
    bool changed = false;
    for (int attempt = 0;; ++attempt) {
        Stage st;
        {
            std::unique_lock<std::mutex> lk(lost_mu_);
            if (!lost_cv_.wait_for(lk, wait_ms, [&] { return !lost_stages_.empty(); })) {
                if (changed) break;
                fail("recover: no runner announced a membership change within " + std::to_string(timeout_s) + " s");
            }
            // several workers may have died, and their announcements travel on different threads: adopt the NEWEST stage, whatever the
            // order of arrival (versions are strictly increasing)
            st = lost_stages_.front();
            for (auto& q : lost_stages_)
                if (q.version > st.version) st = q;
            lost_stages_.clear();
        }
        {
            std::lock_guard<std::mutex> g(mu_);
            if (st.version <= cluster_version_.load()) {
                if (changed) break;
                return {false, false};
            }
            KF_LOGW("recovering: v%d -> v%d, %d -> %d workers", cluster_version_.load(), st.version, (int)cluster_.workers.size(),
                    (int)st.cluster.workers.size());
            cluster_ = st.cluster;
            cluster_version_ = st.version;
            updated_ = false;
            session_.reset();
            changed = true;
        }
        if (!st.cluster.workers.contains(cfg_.self)) {
            detached_ = true;
            return {true, true};
        }
        try {
            update(); // meets the other survivors at the barrier of the new session (holds mu_ meanwhile)
            std::lock_guard<std::mutex> g(lost_mu_);
            if (lost_stages_.empty()) return {true, false};
            // another loss was announced while we were regrouping: go round again with it
        } catch (const std::exception& e) {
            // the barrier of this stage broke: one more member died meanwhile (its announcement shut the mailboxes and interrupted the
            // dials). Take the next stage if it is there or comes within the time-out; otherwise the error goes to the caller.
            KF_LOGW("recover: regrouping as v%d failed (%s); waiting for a newer membership", st.version, e.what());
            if (attempt >= 8) throw;
            std::unique_lock<std::mutex> lk(lost_mu_);
            if (!lost_cv_.wait_for(lk, wait_ms, [&] { return !lost_stages_.empty(); })) throw;
        }
    }
    return {true, false};
}

uint64_t Peer::uid() const {
    const uint64_t hi = cfg_.self.ipv4;
    const uint64_t lo = ((uint64_t)cfg_.self.port << 16) | (uint64_t)(uint16_t)cfg_.init_cluster_version;
    return (hi << 32) | lo;
}

Cluster Peer::current_cluster() {
    std::lock_guard<std::mutex> g(mu_);
    return cluster_;
}

std::shared_ptr<Session> Peer::session() {
    std::lock_guard<std::mutex> g(mu_);
    if (!session_) update_to(cluster_.workers);
    if (!session_) fail("this peer is not part of the current cluster");
    return session_;
}

bool Peer::update() {
    std::lock_guard<std::mutex> g(mu_);
    return update_to(cluster_.workers);
}

bool Peer::update_to(const PeerList& pl) {
    StallGuard sg("updateTo(" + pl.str() + ")");
    if (updated_) {
        server_->set_token((uint32_t)cluster_version_);
        return true;
    }
    // open the mailboxes for the new epoch BEFORE peers of that epoch can connect (they are refused until the token moves)
    coll_.reopen(cluster_version_.load());
    server_->set_token((uint32_t)cluster_version_);
    KF_LOGD("updateTo v%d of %d peers: %s", cluster_version_.load(), (int)pl.size(), pl.str().c_str());
    client_->reset_connections(pl, (uint32_t)cluster_version_);
    if (!pl.contains(cfg_.self)) return false;
    auto sess = std::make_shared<Session>(cfg_.strategy, cfg_.self, pl, client_.get(), &coll_);
    sess->barrier(); // meets the barrier of every other member of the new cluster version
    session_ = sess;
    rank_hint_ = sess->rank();
    updated_ = true;
    NetMonitor::get().set_gauge("kungfu_cluster_version", cluster_version_.load());
    NetMonitor::get().set_gauge("kungfu_cluster_size", (double)pl.size());
    NetMonitor::get().set_gauge("kungfu_rank", sess->rank());
    return true;
}

// ---- elastic ------------------------------------------------------------------------------------------
bool Peer::fetch_cluster(Cluster& out) {
    if (cfg_.config_server.empty()) return false;
    std::string body;
    const int st = http_request("GET", cfg_.config_server, "", &body);
    if (st != 200) return false;
    try {
        out = Cluster::from_json(body);
        return true;
    } catch (const std::exception& e) {
        KF_LOGW("bad cluster config from %s: %s", cfg_.config_server.c_str(), e.what());
        return false;
    }
}

void Peer::propose_new_size(int new_size) {
    Cluster c;
    {
        std::lock_guard<std::mutex> g(mu_);
        c = cluster_.resized(new_size);
    }
    if (cfg_.config_server.empty()) fail("propose_new_size: no config server (KUNGFU_CONFIG_SERVER)");
    std::string resp;
    const int st = http_request("PUT", cfg_.config_server, c.to_json(), &resp);
    if (st != 200) KF_LOGW("propose_new_size(%d): config server answered %d %s", new_size, st, resp.c_str());
}

std::pair<bool, bool> Peer::propose(const Cluster& c) {
    StallGuard sg("propose");
    {
        std::lock_guard<std::mutex> g(mu_);
        if (cluster_ == c) return {false, false};
    }
    const std::string d = c.digest();
    if (!session()->consensus(d.data(), d.size(), "kungfu::propose")) {
        KF_LOGE("diverging proposal among peers; I proposed %s", c.workers.str().c_str());
        return {false, false};
    }
    Stage st;
    {
        std::lock_guard<std::mutex> g(mu_);
        st.version = cluster_version_.load() + 1;
    }
    st.cluster = c;
    const std::string payload = st.to_json();
    const double wait_s = Config::get().wait_runner_timeout_s;
    std::string err = parallel_for((int)c.runners.size(), [&](int i) {
        const PeerID& r = c.runners[i];
        if (!client_->wait_peer(r, wait_s)) fail("runner " + r.str() + " is not reachable");
        client_->send(r, "update", payload.data(), payload.size(), kConnControl, kNoFlag);
    });
    if (!err.empty()) fail("propose: " + err);
    std::lock_guard<std::mutex> g(mu_);
    if (cluster_.workers.disjoint(c.workers)) KF_LOGE("full update detected: %s -> %s! State will be lost.", cluster_.workers.str().c_str(), c.workers.str().c_str());
    else if (!c.workers.empty() && !cluster_.workers.contains(c.workers[0])) KF_LOGE("the new root is a new worker! State will be lost.");
    cluster_ = c;
    cluster_version_ += 1;
    updated_ = false;
    session_.reset();
    const bool keep = c.workers.contains(cfg_.self);
    return {true, !keep};
}

std::pair<bool, bool> Peer::resize_cluster(int new_size) {
    auto sess = session();
    if (sess->rank() == 0) {
        try {
            propose_new_size(new_size);
        } catch (const std::exception& e) {
            KF_LOGW("resize_cluster: %s", e.what());
        }
    }
    // Everybody is in this call (it is collective), so one barrier AFTER the proposal has been stored makes every worker read the new
    // configuration on its first try. Without it the others usually read the old one, the digests differ and each of them goes round
    // the 50 ms retry loop of resize_cluster_from_url — measured with native workers: 8 -> 4 took 256 ms, with the barrier 3 ms.
    if (sess->size() > 1) sess->barrier();
    return resize_cluster_from_url();
}

std::pair<bool, bool> Peer::resize_cluster_from_url() {
    Cluster c;
    for (int i = 0;; ++i) {
        if (!fetch_cluster(c)) {
            if (i == 0) KF_LOGD("no cluster config available, using the current one");
            std::lock_guard<std::mutex> g(mu_);
            c = cluster_;
        }
        const std::string d = c.digest();
        if (session()->consensus(d.data(), d.size(), "kungfu::resize")) break;
        KF_LOGW("peers read different cluster configs (attempt %d), retrying", i + 1);
        sleep_ms(50);
    }
    auto res = propose(c);
    if (res.second) detached_ = true;
    else update();
    return res;
}

// ---- p2p ----------------------------------------------------------------------------------------------
void Peer::save(const std::string& name, const void* data, size_t len) { p2p_->save(name, data, len); }
void Peer::save_version(const std::string& version, const std::string& name, const void* data, size_t len) {
    p2p_->save_version(version, name, data, len);
}
bool Peer::request(int rank, const std::string& version, const std::string& name, void* buf, size_t len) {
    auto sess = session();
    if (rank < 0 || rank >= sess->size()) fail("request: rank out of range");
    return p2p_->request(sess->peers()[rank], version, name, buf, len);
}

} // namespace kfrt
