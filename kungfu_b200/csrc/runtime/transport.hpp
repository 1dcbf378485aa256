// kungfu_b200 runtime — control-plane / CPU-backend transport.
//
// Reference: srcs/go/rchannel/{connection,client,server,handler} — framed simplex connections with a typed handshake
// (8-byte header {type u16, src port u16, src ipv4 u32} answered by a 4-byte ACK {token u32}; token = cluster version, a
// mismatching token on a collective connection is refused = stale-epoch fencing, connection/connection.go:81-88),
// messages {name, flags, payload}, a per-(peer, type) connection pool, TCP + Unix-socket listeners, named rendezvous for
// collectives with zero-copy receive into registered buffers, request/response P2P backed by a blob store, control
// and ping endpoints.
//
// Design differences (ours): readers never block on a full queue (arrivals are parked in an unbounded per-name
// mailbox), blocking work runs on an elastic thread pool instead of goroutines, payload length is 64-bit.
#pragma once
#include <atomic>
#include <condition_variable>
#include <deque>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "base.hpp"
#include "plan.hpp"
#include "store.hpp"

namespace kfrt {

enum ConnType : uint16_t { kConnPing = 0, kConnControl = 1, kConnCollective = 2, kConnPeerToPeer = 3 };
enum MsgFlags : uint32_t { kNoFlag = 0, kWaitRecvBuf = 1, kIsResponse = 2, kRequestFailed = 4 };

// ---- elastic thread pool: a task never waits behind blocked tasks (goroutine-like) -------------------
class Pool {
public:
    static Pool& get();
    void submit(std::function<void()> fn);
    int threads() const { return nthreads_; }

private:
    Pool() = default;
    void worker();
    std::mutex mu_;
    std::condition_variable cv_;
    std::deque<std::function<void()>> q_;
    int idle_ = 0;
    std::atomic<int> nthreads_{0};
};

// Run fn(i) for i in [0, n) concurrently and wait; returns merged error text ("" on success).
std::string parallel_for(int n, const std::function<void(int)>& fn);

// ---- socket -------------------------------------------------------------------------------------------
class Socket {
public:
    Socket() = default;
    explicit Socket(int fd) : fd_(fd) {}
    ~Socket() { close(); }
    Socket(Socket&& o) noexcept : fd_(o.fd_) { o.fd_ = -1; }
    Socket& operator=(Socket&& o) noexcept;
    Socket(const Socket&) = delete;
    int fd() const { return fd_; }
    bool valid() const { return fd_ >= 0; }
    void close();
    void shutdown_rw();
    bool send_all(const void* p, size_t n);
    bool recv_all(void* p, size_t n);
    bool discard(size_t n);
    static Socket connect_tcp(uint32_t ipv4, uint16_t port, int timeout_ms);
    static Socket connect_unix(const std::string& path);

private:
    int fd_ = -1;
};

// ---- monitoring counters (reference: srcs/go/monitor/counters.go:13-164) -------------------------------
class NetMonitor {
public:
    static NetMonitor& get();
    void egress(const PeerID& to, size_t bytes);
    void ingress(const PeerID& from, size_t bytes);
    std::vector<double> egress_rates(const PeerList& peers); // bytes/s since the previous call per peer
    std::map<uint64_t, uint64_t> egress_totals();
    std::string metrics_text(); // prometheus-style
    void reset();
    // application metrics published next to the network counters: gauges are overwritten, counters accumulate. `name` may
    // carry labels, e.g. gpu_collective_bytes_total{op="all_reduce"}
    void set_gauge(const std::string& name, double value);
    void add_counter(const std::string& name, double delta);

private:
    std::mutex mu_;
    std::map<uint64_t, uint64_t> eg_, in_;
    std::map<uint64_t, std::pair<uint64_t, double>> last_; // (total, time) at the previous rate query
    std::map<std::string, double> gauges_, counters_;
};

// ---- client ---------------------------------------------------------------------------------------------
class Client {
public:
    explicit Client(const PeerID& self) : self_(self) {}
    void set_token(uint32_t t) { token_ = t; }
    uint32_t token() const { return token_; }
    // throws kfrt::Error when the peer cannot be reached within the retry budget
    void send(const PeerID& to, const std::string& name, const void* data, size_t len, ConnType type, uint32_t flags);
    double ping(const PeerID& to, int timeout_ms = 2000); // RTT seconds, <0 on failure
    bool wait_peer(const PeerID& to, double timeout_s);   // polls ping until reachable
    void reset_connections(const PeerList& keeps, uint32_t token);
    void close_all();
    // abandon every dial that is being retried right now (a peer is known to be gone: do not sit out the retry budget)
    void interrupt() { ++abort_gen_; }
    // cap the number of dial attempts of this client (0 = the configured budget): notifications to peers that may be dead must not
    // sit out 500 x 200 ms
    void set_max_dial_tries(int n) { max_tries_ = n; }

private:
    struct Conn {
        std::mutex mu;
        Socket sock;
    };
    std::shared_ptr<Conn> get_conn(const PeerID& to, ConnType type);
    Socket dial(const PeerID& to, ConnType type, bool retry);
    PeerID self_;
    std::atomic<uint32_t> token_{0};
    std::atomic<uint64_t> abort_gen_{0};
    std::atomic<int> max_tries_{0};
    std::mutex mu_;
    std::map<std::pair<uint64_t, int>, std::shared_ptr<Conn>> pool_;
};

// ---- message buffers ------------------------------------------------------------------------------------
// Received chunks that no receiver is waiting for yet are parked in `Bytes`. Their storage comes from a size-bucketed free
// list (reference: the sync.Pool buffer pool of srcs/go/rchannel/connection/buffer_pool.go) and is NOT zero-filled: a plain
// std::vector<uint8_t>(1 MiB) costs a malloc, a memset and — once glibc has trimmed the heap — a page fault per 4 KiB,
// for every chunk of every all-reduce.
class BufferPool {
public:
    static BufferPool& get();
    void* take(size_t n);            // >= n bytes, uninitialised
    void give(void* p, size_t n);    // n as passed to take()
    size_t retained_bytes();
    size_t hits();

private:
    static int bucket_of(size_t n);  // -1: not pooled (tiny or huge)
    static constexpr int kMinShift = 9, kMaxShift = 26; // 512 B .. 64 MiB, powers of two
    static constexpr size_t kMaxRetained = 256u << 20;
    std::mutex mu_;
    std::vector<void*> free_[kMaxShift - kMinShift + 1];
    size_t retained_ = 0, hits_ = 0;
};

template <typename T> struct PoolAllocator {
    using value_type = T;
    PoolAllocator() = default;
    template <typename U> PoolAllocator(const PoolAllocator<U>&) {}
    T* allocate(size_t n) { return static_cast<T*>(BufferPool::get().take(n * sizeof(T))); }
    void deallocate(T* p, size_t n) { BufferPool::get().give(p, n * sizeof(T)); }
    template <typename U> void construct(U* p) { ::new (static_cast<void*>(p)) U; } // default-init: bytes stay uninitialised
    template <typename U, typename... A> void construct(U* p, A&&... a) { ::new (static_cast<void*>(p)) U(std::forward<A>(a)...); }
    template <typename U> bool operator==(const PoolAllocator<U>&) const { return true; }
    template <typename U> bool operator!=(const PoolAllocator<U>&) const { return false; }
};

// ---- endpoints ------------------------------------------------------------------------------------------
using Bytes = std::vector<uint8_t, PoolAllocator<uint8_t>>;

class CollectiveEndpoint {
public:
    // blocking receive of the message (src, name) into `buf` (exactly len bytes). Throws on shutdown / size mismatch.
    void recv_into(const PeerID& src, const std::string& name, void* buf, size_t len);
    Bytes recv(const PeerID& src, const std::string& name);
    // called by the server's reader thread; must consume exactly `len` payload bytes from `s`. `conn_epoch` is the cluster
    // version the connection was accepted under: traffic of another epoch (a peer that has not noticed a membership change
    // yet) is consumed and dropped instead of being parked under a name the new session may reuse.
    bool on_message(const PeerID& src, const std::string& name, uint32_t flags, uint64_t len, Socket& s, uint32_t conn_epoch = 0);
    void shutdown(const std::string& why); // wakes every waiter with an error; later receives fail until reopen()
    void reopen(int64_t epoch = -1);       // forgets parked messages; epoch >= 0 turns the fence on
    size_t parked() const;

private:
    struct Waiter {
        void* buf = nullptr;
        size_t len = 0;
        bool into = false;
        bool done = false;
        bool failed = false;
        std::string err;
        Bytes data;
        std::condition_variable cv;
    };
    struct Slot {
        std::deque<Bytes> ready;
        Waiter* waiter = nullptr;
    };
    static std::string key(const PeerID& src, const std::string& name);
    void wait_done(std::unique_lock<std::mutex>& lk, Waiter& w, const std::string& what);
    mutable std::mutex mu_;
    std::unordered_map<std::string, Slot> slots_;
    bool closed_ = false;
    int64_t epoch_ = -1;
    std::string close_reason_;
};

class P2PEndpoint {
public:
    P2PEndpoint(Client* client, const PeerID& self) : client_(client), self_(self) {}
    // request blob `name` (optionally a version) from `target` into buf; returns false if the target has no such blob
    bool request(const PeerID& target, const std::string& version, const std::string& name, void* buf, size_t len);
    void save(const std::string& name, const void* data, size_t len);
    void save_version(const std::string& version, const std::string& name, const void* data, size_t len);
    bool on_message(const PeerID& src, const std::string& name, uint32_t flags, uint64_t len, Socket& s);
    Store& store() { return store_; }
    VersionedStore& versioned() { return vstore_; }
    void shutdown();

private:
    struct Waiter {
        void* buf;
        size_t len;
        bool done = false, ok = false;
        std::condition_variable cv;
    };
    void respond(const PeerID& to, const std::string& name, const std::string& version);
    Client* client_;
    PeerID self_;
    Store store_;
    VersionedStore vstore_{3};
    std::mutex mu_;
    std::map<std::string, std::deque<Waiter*>> waiters_;
    bool closed_ = false;
};

// ---- server ---------------------------------------------------------------------------------------------
class Server {
public:
    using ControlHandler = std::function<void(const PeerID& src, const std::string& name, const Bytes& payload)>;
    Server(const PeerID& self, CollectiveEndpoint* coll, P2PEndpoint* p2p, bool use_unix);
    ~Server();
    void start(); // throws if the port cannot be bound
    void stop();
    void set_token(uint32_t t) { token_ = t; }
    void set_control_handler(ControlHandler h);

private:
    void accept_loop(int lfd);
    void serve(Socket s);
    PeerID self_;
    CollectiveEndpoint* coll_;
    P2PEndpoint* p2p_;
    bool use_unix_;
    std::atomic<uint32_t> token_{0};
    std::atomic<bool> running_{false};
    int tcp_fd_ = -1, unix_fd_ = -1;
    std::vector<std::thread> acceptors_;
    std::mutex mu_;
    ControlHandler control_;
    std::mutex conns_mu_;
    std::map<int, std::shared_ptr<std::atomic<bool>>> live_; // fd -> alive flag (for stop())
    std::atomic<int> active_{0};
};

} // namespace kfrt
