// Implementation of transport.hpp (the header carries the design notes and the reference file:line mapping).
#include "transport.hpp"

#include <arpa/inet.h>
#include <errno.h>
#include <fcntl.h>
#include <netinet/in.h>
#include <netinet/tcp.h>
#include <poll.h>
#include <signal.h>
#include <sys/socket.h>
#include <sys/un.h>
#include <unistd.h>

#include <cstring>
#include <sstream>

namespace kfrt {

// ---------------------------------------------------------------------------------------------------
// Pool
// ---------------------------------------------------------------------------------------------------
Pool& Pool::get() {
    static Pool* p = new Pool();
    return *p;
}
void Pool::submit(std::function<void()> fn) {
    std::unique_lock<std::mutex> lk(mu_);
    q_.push_back(std::move(fn));
    // Tasks block on each other (a chunk's reduce step waits for the peer's task of the same chunk), so a queued task must
    // never sit behind a running one: every queued task needs its own waiting worker. `idle_` still counts a worker that
    // has been notified but has not yet woken up, hence the comparison with the queue length and not with zero — with
    // `idle_ == 0` as the test, two quick submits both "used" the same idle worker and the second task starved
    // (intermittent all-reduce deadlock with >= 7 chunks in flight).
    // No cap that could leave a task queued behind blocked workers: the per-op chunk window (Session::run_strategies) bounds
    // how many workers one collective can hold, and a queued task that nobody can run is a deadlock, not back-pressure.
    if ((int)q_.size() > idle_) {
        ++nthreads_;
        std::thread([this] { worker(); }).detach();
    } else {
        cv_.notify_one();
    }
}
void Pool::worker() {
    std::unique_lock<std::mutex> lk(mu_);
    while (true) {
        if (q_.empty()) {
            ++idle_;
            const bool got = cv_.wait_for(lk, std::chrono::seconds(30), [this] { return !q_.empty(); });
            --idle_;
            if (!got) {
                --nthreads_;
                return;
            }
        }
        auto fn = std::move(q_.front());
        q_.pop_front();
        lk.unlock();
        try {
            fn();
        } catch (const std::exception& e) {
            KF_LOGE("pool task failed: %s", e.what());
        }
        lk.lock();
    }
}

std::string parallel_for(int n, const std::function<void(int)>& fn) {
    if (n <= 0) return "";
    std::vector<std::string> errs(n);
    if (n == 1) {
        try {
            fn(0);
        } catch (const std::exception& e) {
            errs[0] = e.what();
        }
        return merge_errors(errs, "parallel");
    }
    std::mutex mu;
    std::condition_variable cv;
    int left = n;
    for (int i = 0; i < n; ++i) {
        Pool::get().submit([&, i] {
            try {
                fn(i);
            } catch (const std::exception& e) {
                errs[i] = e.what();
            }
            std::lock_guard<std::mutex> g(mu);
            if (--left == 0) cv.notify_all();
        });
    }
    std::unique_lock<std::mutex> lk(mu);
    cv.wait(lk, [&] { return left == 0; });
    return merge_errors(errs, "parallel");
}

// ---------------------------------------------------------------------------------------------------
// Socket
// ---------------------------------------------------------------------------------------------------
Socket& Socket::operator=(Socket&& o) noexcept {
    if (this != &o) {
        close();
        fd_ = o.fd_;
        o.fd_ = -1;
    }
    return *this;
}
void Socket::close() {
    if (fd_ >= 0) {
        ::close(fd_);
        fd_ = -1;
    }
}
void Socket::shutdown_rw() {
    if (fd_ >= 0) ::shutdown(fd_, SHUT_RDWR);
}
bool Socket::send_all(const void* p, size_t n) {
    const char* c = static_cast<const char*>(p);
    while (n > 0) {
        ssize_t k = ::send(fd_, c, n, MSG_NOSIGNAL);
        if (k < 0) {
            if (errno == EINTR) continue;
            return false;
        }
        c += k;
        n -= (size_t)k;
    }
    return true;
}
bool Socket::recv_all(void* p, size_t n) {
    char* c = static_cast<char*>(p);
    while (n > 0) {
        ssize_t k = ::recv(fd_, c, n, 0);
        if (k < 0) {
            if (errno == EINTR) continue;
            return false;
        }
        if (k == 0) return false;
        c += k;
        n -= (size_t)k;
    }
    return true;
}
bool Socket::discard(size_t n) {
    char tmp[4096];
    while (n > 0) {
        size_t k = n < sizeof(tmp) ? n : sizeof(tmp);
        if (!recv_all(tmp, k)) return false;
        n -= k;
    }
    return true;
}
Socket Socket::connect_tcp(uint32_t ipv4, uint16_t port, int timeout_ms) {
    int fd = ::socket(AF_INET, SOCK_STREAM | SOCK_CLOEXEC, 0);
    if (fd < 0) return Socket();
    sockaddr_in a{};
    a.sin_family = AF_INET;
    a.sin_port = htons(port);
    a.sin_addr.s_addr = htonl(ipv4);
    // the automatically chosen source port of this connection may be a port some worker wants to LISTEN on later (default
    // runner port 38080 and torchrun-style jobs sit in the ephemeral range): with SO_REUSEADDR on both sides the listener's bind
    // still succeeds
    int reuse = 1;
    setsockopt(fd, SOL_SOCKET, SO_REUSEADDR, &reuse, sizeof(reuse));
    int fl = fcntl(fd, F_GETFL, 0);
    fcntl(fd, F_SETFL, fl | O_NONBLOCK);
    int r = ::connect(fd, reinterpret_cast<sockaddr*>(&a), sizeof(a));
    if (r != 0 && errno == EINPROGRESS) {
        pollfd p{fd, POLLOUT, 0};
        if (::poll(&p, 1, timeout_ms) == 1) {
            int err = 0;
            socklen_t l = sizeof(err);
            getsockopt(fd, SOL_SOCKET, SO_ERROR, &err, &l);
            r = err == 0 ? 0 : -1;
        } else {
            r = -1;
        }
    }
    if (r != 0) {
        ::close(fd);
        return Socket();
    }
    fcntl(fd, F_SETFL, fl);
    int one = 1;
    setsockopt(fd, IPPROTO_TCP, TCP_NODELAY, &one, sizeof(one));
    return Socket(fd);
}
Socket Socket::connect_unix(const std::string& path) {
    int fd = ::socket(AF_UNIX, SOCK_STREAM | SOCK_CLOEXEC, 0);
    if (fd < 0) return Socket();
    sockaddr_un a{};
    a.sun_family = AF_UNIX;
    strncpy(a.sun_path, path.c_str(), sizeof(a.sun_path) - 1);
    if (::connect(fd, reinterpret_cast<sockaddr*>(&a), sizeof(a)) != 0) {
        ::close(fd);
        return Socket();
    }
    return Socket(fd);
}

// ---------------------------------------------------------------------------------------------------
// wire format
// ---------------------------------------------------------------------------------------------------
namespace {
#pragma pack(push, 1)
struct ConnHeader {
    uint16_t type;
    uint16_t src_port;
    uint32_t src_ipv4;
};
struct ConnAck {
    uint32_t token;
};
struct MsgHeader {
    uint32_t name_len;
    uint32_t flags;
    uint64_t data_len;
};
#pragma pack(pop)
constexpr uint32_t kMaxNameLen = 4096;

bool send_message(Socket& s, const std::string& name, const void* data, size_t len, uint32_t flags) {
    MsgHeader h{(uint32_t)name.size(), flags, (uint64_t)len};
    // one buffer for header + name keeps small messages in one segment
    std::string head(reinterpret_cast<const char*>(&h), sizeof(h));
    head += name;
    if (!s.send_all(head.data(), head.size())) return false;
    return len == 0 || s.send_all(data, len);
}
} // namespace

// ---------------------------------------------------------------------------------------------------
// NetMonitor
// ---------------------------------------------------------------------------------------------------
NetMonitor& NetMonitor::get() {
    static NetMonitor* m = new NetMonitor();
    return *m;
}
void NetMonitor::egress(const PeerID& to, size_t bytes) {
    std::lock_guard<std::mutex> g(mu_);
    eg_[to.key()] += bytes;
}
void NetMonitor::ingress(const PeerID& from, size_t bytes) {
    std::lock_guard<std::mutex> g(mu_);
    in_[from.key()] += bytes;
}
std::vector<double> NetMonitor::egress_rates(const PeerList& peers) {
    std::lock_guard<std::mutex> g(mu_);
    const double t = now_s();
    std::vector<double> out;
    for (auto& p : peers) {
        const uint64_t tot = eg_[p.key()];
        auto& l = last_[p.key()];
        const double dt = t - l.second;
        out.push_back(l.second > 0 && dt > 0 ? (double)(tot - l.first) / dt : 0.0);
        l = {tot, t};
    }
    return out;
}
std::map<uint64_t, uint64_t> NetMonitor::egress_totals() {
    std::lock_guard<std::mutex> g(mu_);
    return eg_;
}
std::string NetMonitor::metrics_text() {
    std::lock_guard<std::mutex> g(mu_);
    std::ostringstream os;
    for (auto& kv : eg_) {
        PeerID p{(uint32_t)(kv.first >> 16), (uint16_t)(kv.first & 0xffff)};
        os << "egress_total_bytes{peer=\"" << p.str() << "\"} " << kv.second << "\n";
    }
    for (auto& kv : in_) {
        PeerID p{(uint32_t)(kv.first >> 16), (uint16_t)(kv.first & 0xffff)};
        os << "ingress_total_bytes{peer=\"" << p.str() << "\"} " << kv.second << "\n";
    }
    os.precision(12);
    for (auto& kv : gauges_) os << kv.first << " " << kv.second << "\n";
    for (auto& kv : counters_) os << kv.first << " " << kv.second << "\n";
    return os.str();
}
void NetMonitor::set_gauge(const std::string& name, double value) {
    std::lock_guard<std::mutex> g(mu_);
    gauges_[name] = value;
}
void NetMonitor::add_counter(const std::string& name, double delta) {
    std::lock_guard<std::mutex> g(mu_);
    counters_[name] += delta;
}
void NetMonitor::reset() {
    std::lock_guard<std::mutex> g(mu_);
    eg_.clear();
    in_.clear();
    last_.clear();
}

// ---------------------------------------------------------------------------------------------------
// Client
// ---------------------------------------------------------------------------------------------------
Socket Client::dial(const PeerID& to, ConnType type, bool retry) {
    const Config& cfg = Config::get();
    // Collective / control traffic waits for its peer for as long as the configured budget says (peers of a new epoch come up at their
    // own pace). A peer-to-peer request is asynchronous by nature (model pulls of PairAveraging): its target is either there or it is
    // not worth waiting 100 s for — 15 attempts (3 s with the default period) cover a restart blip.
    int tries = retry ? (type == kConnPeerToPeer ? std::min(cfg.conn_retry_count, 15) : cfg.conn_retry_count) : 1;
    if (retry && max_tries_.load() > 0) tries = std::min(tries, max_tries_.load());
    const uint64_t gen = abort_gen_.load();
    for (int attempt = 0; attempt < tries; ++attempt) {
        if (abort_gen_.load() != gen) break; // interrupt(): the peer is known to be lost
        // One attempt = one retry period. A peer that ANSWERS but is still in another epoch (it has not adopted the new cluster version
        // yet — after a resize the members switch within milliseconds of each other) is asked again after 1, 2, 4, ... ms inside the
        // period instead of once per period: a shrink 8 -> 4 cost a full 200 ms whenever one survivor dialled before another had
        // switched (measured, 2 runs in 3), now it costs the few ms the slowest member needs.
        int waited = 0, step = 1;
        for (;;) {
            bool other_epoch = false;
            Socket s;
            if (cfg.use_unix_sock && to.colocated_with(self_)) s = Socket::connect_unix(to.sock_file());
            if (!s.valid()) s = Socket::connect_tcp(to.ipv4, to.port, 2000);
            if (s.valid()) {
                ConnHeader h{(uint16_t)type, self_.port, self_.ipv4};
                ConnAck ack{};
                if (s.send_all(&h, sizeof(h)) && s.recv_all(&ack, sizeof(ack))) {
                    // stale-epoch fencing: collective traffic only flows between peers of the same cluster version
                    if (type != kConnCollective || ack.token == token_.load()) return s;
                    KF_LOGD("token mismatch dialing %s: mine %u theirs %u", to.str().c_str(), token_.load(), ack.token);
                    other_epoch = true;
                }
            }
            if (!other_epoch || !retry || waited >= cfg.conn_retry_period_ms || abort_gen_.load() != gen) break;
            step = std::min(step, cfg.conn_retry_period_ms - waited);
            sleep_ms(step);
            waited += step;
            step *= 2;
        }
        if (attempt + 1 < tries && waited < cfg.conn_retry_period_ms) sleep_ms(cfg.conn_retry_period_ms - waited);
    }
    return Socket();
}

std::shared_ptr<Client::Conn> Client::get_conn(const PeerID& to, ConnType type) {
    std::lock_guard<std::mutex> g(mu_);
    auto& c = pool_[{to.key(), (int)type}];
    if (!c) c = std::make_shared<Conn>();
    return c;
}

void Client::send(const PeerID& to, const std::string& name, const void* data, size_t len, ConnType type, uint32_t flags) {
    if (name.size() > kMaxNameLen) fail("message name too long");
    auto c = get_conn(to, type);
    std::lock_guard<std::mutex> g(c->mu);
    for (int attempt = 0; attempt < 2; ++attempt) {
        if (!c->sock.valid()) {
            c->sock = dial(to, type, true);
            if (!c->sock.valid()) fail("cannot connect to " + to.str() + " (" + std::to_string((int)type) + ")");
        }
        if (send_message(c->sock, name, data, len, flags)) {
            NetMonitor::get().egress(to, len);
            return;
        }
        c->sock.close(); // peer restarted? redial once
    }
    fail("send to " + to.str() + " failed: " + strerror(errno));
}

double Client::ping(const PeerID& to, int timeout_ms) {
    const double t0 = now_s();
    Socket s;
    if (Config::get().use_unix_sock && to.colocated_with(self_)) s = Socket::connect_unix(to.sock_file());
    if (!s.valid()) s = Socket::connect_tcp(to.ipv4, to.port, timeout_ms);
    if (!s.valid()) return -1;
    ConnHeader h{(uint16_t)kConnPing, self_.port, self_.ipv4};
    ConnAck ack{};
    if (!s.send_all(&h, sizeof(h)) || !s.recv_all(&ack, sizeof(ack))) return -1;
    const char payload[8] = {'p', 'i', 'n', 'g', 0, 0, 0, 0};
    if (!send_message(s, "ping", payload, sizeof(payload), kNoFlag)) return -1;
    MsgHeader mh{};
    pollfd p{s.fd(), POLLIN, 0};
    if (::poll(&p, 1, timeout_ms) != 1) return -1;
    if (!s.recv_all(&mh, sizeof(mh)) || mh.name_len > kMaxNameLen) return -1;
    if (!s.discard(mh.name_len + mh.data_len)) return -1;
    return now_s() - t0;
}

bool Client::wait_peer(const PeerID& to, double timeout_s) {
    const double t0 = now_s();
    while (true) {
        if (ping(to, 500) >= 0) return true;
        if (now_s() - t0 > timeout_s) return false;
        sleep_ms(Config::get().conn_retry_period_ms);
    }
}

void Client::reset_connections(const PeerList& keeps, uint32_t token) {
    token_ = token;
    std::lock_guard<std::mutex> g(mu_);
    // collective connections carry the old token: all of them are re-dialled lazily; connections to departed
    // peers are dropped for every type
    for (auto it = pool_.begin(); it != pool_.end();) {
        PeerID p{(uint32_t)(it->first.first >> 16), (uint16_t)(it->first.first & 0xffff)};
        const bool keep = keeps.contains(p) && it->first.second != (int)kConnCollective;
        if (!keep) {
            // keep the connection object alive past the erase: the guard below unlocks a mutex that lives INSIDE it (TSan: heap-use-after-
            // free when the pool held the last reference)
            std::shared_ptr<Conn> conn = it->second;
            {
                std::lock_guard<std::mutex> g2(conn->mu);
                conn->sock.close();
            }
            it = pool_.erase(it);
        } else {
            ++it;
        }
    }
}

void Client::close_all() {
    std::lock_guard<std::mutex> g(mu_);
    for (auto& kv : pool_) {
        std::lock_guard<std::mutex> g2(kv.second->mu);
        kv.second->sock.close();
    }
    pool_.clear();
}

// ---------------------------------------------------------------------------------------------------
// BufferPool
// ---------------------------------------------------------------------------------------------------
BufferPool& BufferPool::get() {
    static BufferPool* p = new BufferPool(); // leaked on purpose: buffers may be returned during static destruction
    return *p;
}
int BufferPool::bucket_of(size_t n) {
    if (n > ((size_t)1 << kMaxShift)) return -1;
    int s = kMinShift;
    while (((size_t)1 << s) < n) ++s;
    return s - kMinShift;
}
void* BufferPool::take(size_t n) {
    const int b = bucket_of(n);
    if (b < 0) return ::operator new(n);
    {
        std::lock_guard<std::mutex> g(mu_);
        if (!free_[b].empty()) {
            void* p = free_[b].back();
            free_[b].pop_back();
            retained_ -= (size_t)1 << (b + kMinShift);
            ++hits_;
            return p;
        }
    }
    return ::operator new((size_t)1 << (b + kMinShift));
}
void BufferPool::give(void* p, size_t n) {
    if (!p) return;
    const int b = bucket_of(n);
    if (b >= 0) {
        const size_t sz = (size_t)1 << (b + kMinShift);
        std::lock_guard<std::mutex> g(mu_);
        if (retained_ + sz <= kMaxRetained) {
            free_[b].push_back(p);
            retained_ += sz;
            return;
        }
    }
    ::operator delete(p);
}
size_t BufferPool::retained_bytes() {
    std::lock_guard<std::mutex> g(mu_);
    return retained_;
}
size_t BufferPool::hits() {
    std::lock_guard<std::mutex> g(mu_);
    return hits_;
}

// ---------------------------------------------------------------------------------------------------
// CollectiveEndpoint
// ---------------------------------------------------------------------------------------------------
std::string CollectiveEndpoint::key(const PeerID& src, const std::string& name) {
    std::string k(reinterpret_cast<const char*>(&src.ipv4), 4);
    k.append(reinterpret_cast<const char*>(&src.port), 2);
    k += name;
    return k;
}

void CollectiveEndpoint::wait_done(std::unique_lock<std::mutex>& lk, Waiter& w, const std::string& what) {
    StallGuard sg(what, /*fine=*/true);
    w.cv.wait(lk, [&] { return w.done; });
    if (w.failed) fail(w.err);
}

void CollectiveEndpoint::recv_into(const PeerID& src, const std::string& name, void* buf, size_t len) {
    const std::string k = key(src, name);
    std::unique_lock<std::mutex> lk(mu_);
    if (closed_) fail("collective endpoint closed: " + close_reason_);
    Slot& slot = slots_[k];
    if (!slot.ready.empty()) {
        Bytes b = std::move(slot.ready.front());
        slot.ready.pop_front();
        if (slot.ready.empty() && !slot.waiter) slots_.erase(k);
        lk.unlock();
        if (b.size() != len) fail("recv_into(" + name + "): size mismatch " + std::to_string(b.size()) + " != " + std::to_string(len));
        if (len) memcpy(buf, b.data(), len);
        return;
    }
    if (slot.waiter) fail("two concurrent receives for the same (peer, name): " + name);
    Waiter w;
    w.buf = buf;
    w.len = len;
    w.into = true;
    slot.waiter = &w;
    wait_done(lk, w, "recv_into(" + name + " from " + src.str() + ")");
}

Bytes CollectiveEndpoint::recv(const PeerID& src, const std::string& name) {
    const std::string k = key(src, name);
    std::unique_lock<std::mutex> lk(mu_);
    if (closed_) fail("collective endpoint closed: " + close_reason_);
    Slot& slot = slots_[k];
    if (!slot.ready.empty()) {
        Bytes b = std::move(slot.ready.front());
        slot.ready.pop_front();
        if (slot.ready.empty() && !slot.waiter) slots_.erase(k);
        return b;
    }
    if (slot.waiter) fail("two concurrent receives for the same (peer, name): " + name);
    Waiter w;
    slot.waiter = &w;
    wait_done(lk, w, "recv(" + name + " from " + src.str() + ")");
    return std::move(w.data);
}

bool CollectiveEndpoint::on_message(const PeerID& src, const std::string& name, uint32_t, uint64_t len, Socket& s, uint32_t conn_epoch) {
    const std::string k = key(src, name);
    Waiter* w = nullptr;
    {
        std::lock_guard<std::mutex> g(mu_);
        if (closed_ || (epoch_ >= 0 && (int64_t)conn_epoch != epoch_)) {
            KF_LOGD("dropping %s from %s: %s", name.c_str(), src.str().c_str(), closed_ ? "endpoint closed" : "stale epoch");
            return s.discard(len);
        }
        auto it = slots_.find(k);
        if (it != slots_.end() && it->second.waiter) {
            w = it->second.waiter;
            it->second.waiter = nullptr;
            if (it->second.ready.empty()) slots_.erase(it);
        }
    }
    if (w) {
        bool ok = true;
        std::string err;
        if (w->into) {
            if (w->len != len) {
                ok = s.discard(len);
                err = "recv_into(" + name + "): size mismatch " + std::to_string(len) + " != " + std::to_string(w->len);
            } else {
                ok = s.recv_all(w->buf, len); // zero-copy into the registered buffer
                if (!ok) err = "connection lost while receiving " + name;
            }
        } else {
            w->data.resize(len);
            ok = len == 0 || s.recv_all(w->data.data(), len);
            if (!ok) err = "connection lost while receiving " + name;
        }
        std::lock_guard<std::mutex> g(mu_);
        w->failed = !err.empty();
        w->err = err;
        w->done = true;
        w->cv.notify_all();
        return ok;
    }
    Bytes b(len);
    if (len && !s.recv_all(b.data(), len)) return false;
    std::lock_guard<std::mutex> g(mu_);
    if (closed_ || (epoch_ >= 0 && (int64_t)conn_epoch != epoch_)) return true; // membership changed while we were reading
    Slot& slot = slots_[k];
    if (slot.waiter) { // a receiver registered while we were reading
        Waiter* w2 = slot.waiter;
        slot.waiter = nullptr;
        if (w2->into) {
            if (w2->len != b.size()) {
                w2->failed = true;
                w2->err = "recv_into(" + name + "): size mismatch";
            } else if (len) {
                memcpy(w2->buf, b.data(), len);
            }
        } else {
            w2->data = std::move(b);
        }
        w2->done = true;
        w2->cv.notify_all();
        if (slot.ready.empty()) slots_.erase(k);
    } else {
        slot.ready.push_back(std::move(b));
    }
    return true;
}

void CollectiveEndpoint::shutdown(const std::string& why) {
    std::lock_guard<std::mutex> g(mu_);
    closed_ = true;
    close_reason_ = why;
    for (auto& kv : slots_)
        if (kv.second.waiter) {
            kv.second.waiter->failed = true;
            kv.second.waiter->err = "collective endpoint closed: " + why;
            kv.second.waiter->done = true;
            kv.second.waiter->cv.notify_all();
            kv.second.waiter = nullptr;
        }
    slots_.clear();
}
void CollectiveEndpoint::reopen(int64_t epoch) {
    std::lock_guard<std::mutex> g(mu_);
    if (closed_) slots_.clear(); // nothing parked before a shutdown belongs to the session that comes next
    closed_ = false;
    if (epoch >= 0) epoch_ = epoch;
    close_reason_.clear();
}
size_t CollectiveEndpoint::parked() const {
    std::lock_guard<std::mutex> g(mu_);
    size_t n = 0;
    for (auto& kv : slots_) n += kv.second.ready.size();
    return n;
}

// ---------------------------------------------------------------------------------------------------
// P2PEndpoint
// ---------------------------------------------------------------------------------------------------
static std::string p2p_key(const PeerID& p, const std::string& name) { return p.str() + "/" + name; }

bool P2PEndpoint::request(const PeerID& target, const std::string& version, const std::string& name, void* buf, size_t len) {
    if (target == self_) { // local fast path
        auto b = version.empty() ? store_.get(name) : vstore_.get(version, name);
        if (!b) return false;
        std::shared_lock<std::shared_mutex> r(b->mu);
        if (b->data.size() != len) return false;
        memcpy(buf, b->data.data(), len);
        return true;
    }
    Waiter w{buf, len};
    const std::string k = p2p_key(target, name);
    {
        std::lock_guard<std::mutex> g(mu_);
        if (closed_) fail("p2p endpoint closed");
        waiters_[k].push_back(&w);
    }
    try {
        client_->send(target, name, version.data(), version.size(), kConnPeerToPeer, kNoFlag);
    } catch (...) {
        std::lock_guard<std::mutex> g(mu_);
        auto& q = waiters_[k];
        for (auto it = q.begin(); it != q.end(); ++it)
            if (*it == &w) {
                q.erase(it);
                break;
            }
        throw;
    }
    StallGuard sg("request(" + name + " from " + target.str() + ")");
    std::unique_lock<std::mutex> lk(mu_);
    w.cv.wait(lk, [&] { return w.done; });
    return w.ok;
}

void P2PEndpoint::save(const std::string& name, const void* data, size_t len) { store_.save(name, data, len); }
void P2PEndpoint::save_version(const std::string& version, const std::string& name, const void* data, size_t len) {
    vstore_.save(version, name, data, len);
}

void P2PEndpoint::respond(const PeerID& to, const std::string& name, const std::string& version) {
    auto b = version.empty() ? store_.get(name) : vstore_.get(version, name);
    try {
        if (!b) {
            client_->send(to, name, nullptr, 0, kConnPeerToPeer, kIsResponse | kRequestFailed);
            return;
        }
        std::shared_lock<std::shared_mutex> r(b->mu); // readers never block the owner for longer than one send
        client_->send(to, name, b->data.data(), b->data.size(), kConnPeerToPeer, kIsResponse);
    } catch (const std::exception& e) {
        KF_LOGW("p2p response to %s failed: %s", to.str().c_str(), e.what());
    }
}

bool P2PEndpoint::on_message(const PeerID& src, const std::string& name, uint32_t flags, uint64_t len, Socket& s) {
    if (flags & kIsResponse) {
        Waiter* w = nullptr;
        {
            std::lock_guard<std::mutex> g(mu_);
            auto it = waiters_.find(p2p_key(src, name));
            if (it != waiters_.end() && !it->second.empty()) {
                w = it->second.front();
                it->second.pop_front();
                if (it->second.empty()) waiters_.erase(it);
            }
        }
        bool ok = true, good = false;
        if (w && !(flags & kRequestFailed) && w->len == len) {
            ok = s.recv_all(w->buf, len);
            good = ok;
        } else {
            ok = s.discard(len);
        }
        if (w) {
            std::lock_guard<std::mutex> g(mu_);
            w->ok = good;
            w->done = true;
            w->cv.notify_all();
        }
        return ok;
    }
    if (len > kMaxNameLen) { // a request carries a version label, nothing more: never allocate what a stranger announces
        KF_LOGW("dropping connection from %s: %llu-byte version label in a p2p request", src.str().c_str(), (unsigned long long)len);
        return false;
    }
    std::string version(len, '\0');
    if (len && !s.recv_all(&version[0], len)) return false;
    Pool::get().submit([this, src, name, version] { respond(src, name, version); });
    return true;
}

void P2PEndpoint::shutdown() {
    std::lock_guard<std::mutex> g(mu_);
    closed_ = true;
    for (auto& kv : waiters_)
        for (auto* w : kv.second) {
            w->ok = false;
            w->done = true;
            w->cv.notify_all();
        }
    waiters_.clear();
}

// ---------------------------------------------------------------------------------------------------
// Server
// ---------------------------------------------------------------------------------------------------
Server::Server(const PeerID& self, CollectiveEndpoint* coll, P2PEndpoint* p2p, bool use_unix)
    : self_(self), coll_(coll), p2p_(p2p), use_unix_(use_unix) {}

Server::~Server() { stop(); }

void Server::set_control_handler(ControlHandler h) {
    std::lock_guard<std::mutex> g(mu_);
    control_ = std::move(h);
}

void Server::start() {
    if (running_) return;
    signal(SIGPIPE, SIG_IGN);
    tcp_fd_ = ::socket(AF_INET, SOCK_STREAM | SOCK_CLOEXEC, 0);
    if (tcp_fd_ < 0) fail("socket() failed");
    int one = 1;
    setsockopt(tcp_fd_, SOL_SOCKET, SO_REUSEADDR, &one, sizeof(one));
    sockaddr_in a{};
    a.sin_family = AF_INET;
    a.sin_port = htons(self_.port);
    // Listen on the peer's own address when it is a loopback alias (127.x.y.z): several "hosts" can then share one machine
    // (127.0.0.1, 127.0.0.2, ... each with the same port numbering), which is how multi-host runs are tested without a cluster
    // (the reference needs docker-compose for that). Real addresses keep INADDR_ANY: behind NAT the advertised address need
    // not be a local interface.
    const bool loopback_alias = (self_.ipv4 >> 24) == 127;
    a.sin_addr.s_addr = htonl(loopback_alias ? self_.ipv4 : INADDR_ANY);
    // A port that is still held by a process on its way out (the previous owner of this slot in an elastic job, the metrics endpoint of a
    // worker that is being reaped) frees itself within moments: wait for it briefly instead of failing the worker.
    int bound = -1;
    for (int attempt = 0; attempt < 30; ++attempt) {
        bound = ::bind(tcp_fd_, reinterpret_cast<sockaddr*>(&a), sizeof(a));
        if (bound == 0 || errno != EADDRINUSE) break;
        if (attempt == 0) KF_LOGW("port %d is in use, waiting for it", (int)self_.port);
        sleep_ms(100);
    }
    if (bound != 0 || ::listen(tcp_fd_, 256) != 0) {
        std::string e = strerror(errno);
        ::close(tcp_fd_);
        tcp_fd_ = -1;
        fail("cannot listen on port " + std::to_string(self_.port) + ": " + e);
    }
    if (use_unix_) {
        const std::string path = self_.sock_file();
        ::unlink(path.c_str());
        unix_fd_ = ::socket(AF_UNIX, SOCK_STREAM | SOCK_CLOEXEC, 0);
        sockaddr_un u{};
        u.sun_family = AF_UNIX;
        strncpy(u.sun_path, path.c_str(), sizeof(u.sun_path) - 1);
        if (unix_fd_ < 0 || ::bind(unix_fd_, reinterpret_cast<sockaddr*>(&u), sizeof(u)) != 0 || ::listen(unix_fd_, 256) != 0) {
            KF_LOGW("unix socket %s unavailable (%s); TCP only", path.c_str(), strerror(errno));
            if (unix_fd_ >= 0) ::close(unix_fd_);
            unix_fd_ = -1;
        }
    }
    running_ = true;
    acceptors_.emplace_back([this] { accept_loop(tcp_fd_); });
    if (unix_fd_ >= 0) acceptors_.emplace_back([this] { accept_loop(unix_fd_); });
}

void Server::stop() {
    if (!running_.exchange(false)) return;
    if (tcp_fd_ >= 0) ::shutdown(tcp_fd_, SHUT_RDWR);
    if (unix_fd_ >= 0) ::shutdown(unix_fd_, SHUT_RDWR);
    for (auto& t : acceptors_)
        if (t.joinable()) t.join();
    acceptors_.clear();
    if (tcp_fd_ >= 0) ::close(tcp_fd_);
    if (unix_fd_ >= 0) {
        ::close(unix_fd_);
        ::unlink(self_.sock_file().c_str());
    }
    tcp_fd_ = unix_fd_ = -1;
    {
        std::lock_guard<std::mutex> g(conns_mu_);
        for (auto& kv : live_) ::shutdown(kv.first, SHUT_RDWR);
    }
    for (int i = 0; i < 200 && active_.load() > 0; ++i) sleep_ms(10);
}

void Server::accept_loop(int lfd) {
    while (running_) {
        pollfd p{lfd, POLLIN, 0};
        int pr = ::poll(&p, 1, 200);
        if (pr <= 0) continue;
        int fd = ::accept4(lfd, nullptr, nullptr, SOCK_CLOEXEC);
        if (fd < 0) {
            if (!running_) break;
            continue;
        }
        int one = 1;
        setsockopt(fd, IPPROTO_TCP, TCP_NODELAY, &one, sizeof(one)); // harmless failure on unix sockets
        ++active_;
        std::thread([this, fd] {
            {
                std::lock_guard<std::mutex> g(conns_mu_);
                live_[fd] = nullptr;
            }
            try {
                serve(Socket(fd));
            } catch (const std::exception& e) {
                KF_LOGW("connection handler: %s", e.what());
            }
            {
                std::lock_guard<std::mutex> g(conns_mu_);
                live_.erase(fd);
            }
            --active_;
        }).detach();
    }
}

void Server::serve(Socket s) {
    ConnHeader h{};
    if (!s.recv_all(&h, sizeof(h))) return;
    ConnAck ack{token_.load()};
    if (!s.send_all(&ack, sizeof(ack))) return;
    const PeerID src{h.src_ipv4, h.src_port};
    const ConnType type = (ConnType)h.type;
    while (running_) {
        MsgHeader mh{};
        if (!s.recv_all(&mh, sizeof(mh))) return;
        if (mh.name_len > kMaxNameLen) return;
        std::string name(mh.name_len, '\0');
        if (mh.name_len && !s.recv_all(&name[0], mh.name_len)) return;
        // Lengths come from the wire: never allocate what an unauthenticated sender asks for. Control / ping messages are
        // small by construction; data messages into a registered buffer are checked against that buffer's size, everything
        // else is bounded by KUNGFU_MAX_MESSAGE_MB (default 4096).
        static const uint64_t kMaxData = (uint64_t)std::max(1.0, parse_duration_s(getenv_str("KUNGFU_MAX_MESSAGE_MB", "4096"), 4096)) << 20;
        const uint64_t cap = type == kConnControl ? (16ull << 20) : type == kConnPing ? (1ull << 20) : kMaxData;
        if (mh.data_len > cap) {
            KF_LOGW("dropping connection from %s: %llu-byte message on a type-%d connection exceeds the limit", src.str().c_str(),
                    (unsigned long long)mh.data_len, (int)type);
            return;
        }
        NetMonitor::get().ingress(src, mh.data_len);
        switch (type) {
        case kConnCollective:
            if (!coll_ || !coll_->on_message(src, name, mh.flags, mh.data_len, s, ack.token)) return;
            break;
        case kConnPeerToPeer:
            if (!p2p_ || !p2p_->on_message(src, name, mh.flags, mh.data_len, s)) return;
            break;
        case kConnControl: {
            Bytes payload(mh.data_len);
            if (mh.data_len && !s.recv_all(payload.data(), mh.data_len)) return;
            ControlHandler hnd;
            {
                std::lock_guard<std::mutex> g(mu_);
                hnd = control_;
            }
            if (hnd) hnd(src, name, payload);
            break;
        }
        case kConnPing: {
            Bytes payload(mh.data_len);
            if (mh.data_len && !s.recv_all(payload.data(), mh.data_len)) return;
            if (!send_message(s, name, payload.data(), payload.size(), kIsResponse)) return;
            break;
        }
        default: return;
        }
    }
}

} // namespace kfrt
