// kungfu_b200 — C++ unit tests of the native runtime, independent of Python (the pybind layer is tested by pytest).
//
// Counterpart of the reference's gtest suite (tests/cpp/unit/test_{kungfu,kungfu_world,operations,mst,placement}.cpp,
// srcs/cmake/tests.cmake:58-63) and of its Go unit tests for the pieces that are C++ here (plan/topology_test.go,
// plan/{cluster,hostspec}_test.go, plan/hostfile/hostfile_test.go, plan/graph/graph_test.go, store/versionedstore_test.go).
// No framework dependency: CHECK records a failure with file:line, main() prints a summary and returns the failure count.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <numeric>
#include <random>
#include <set>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#include <unistd.h>

#include "base.hpp"
#include "graph.hpp"
#include "numa.hpp"
#include "peer.hpp"
#include "plan.hpp"
#include "reduce.hpp"
#include "store.hpp"
#include "transport.hpp"

#include "fake_model.hpp"
#include "runner_core.hpp"

using namespace kfrt;

static int g_fail = 0, g_checks = 0;
#define CHECK(cond)                                                                  \
    do {                                                                             \
        ++g_checks;                                                                  \
        if (!(cond)) {                                                               \
            ++g_fail;                                                                \
            std::fprintf(stderr, "FAIL %s:%d: %s\n", __FILE__, __LINE__, #cond);     \
        }                                                                            \
    } while (0)
#define CHECK_THROWS(expr)                                                           \
    do {                                                                             \
        ++g_checks;                                                                  \
        bool threw = false;                                                          \
        try {                                                                        \
            (void)(expr);                                                            \
        } catch (const std::exception&) {                                            \
            threw = true;                                                            \
        }                                                                            \
        if (!threw) {                                                                \
            ++g_fail;                                                                \
            std::fprintf(stderr, "FAIL %s:%d: expected an exception: %s\n", __FILE__, __LINE__, #expr); \
        }                                                                            \
    } while (0)

static PeerList peers_on(std::vector<std::pair<const char*, int>> hosts) {
    PeerList pl;
    for (auto& h : hosts)
        for (int i = 0; i < h.second; ++i) pl.push_back(PeerID{parse_ipv4(h.first), (uint16_t)(10000 + i)});
    return pl;
}

// ---- dtype sizes / reduce kernels (test_operations.cpp) ---------------------------------------------------------------
static void test_dtypes_and_transform2() {
    CHECK(dtype_size(kF32) == 4 && dtype_size(kBF16) == 2 && dtype_size(kF16) == 2 && dtype_size(kI64) == 8 && dtype_size(kU8) == 1 &&
          dtype_size(kF64) == 8 && dtype_size(kI16) == 2);
    std::mt19937 rng(7);
    for (size_t n : {1, 7, 64, 1000, 4099}) {
        std::vector<float> x(n), y(n), z(n);
        for (size_t i = 0; i < n; ++i) x[i] = (float)(rng() % 100) / 7.f, y[i] = (float)(rng() % 100) / 3.f;
        transform2(z.data(), x.data(), y.data(), n, kF32, kSum);
        bool ok = true;
        for (size_t i = 0; i < n; ++i) ok = ok && z[i] == x[i] + y[i];
        CHECK(ok);
        transform2(z.data(), x.data(), y.data(), n, kF32, kMax);
        ok = true;
        for (size_t i = 0; i < n; ++i) ok = ok && z[i] == std::max(x[i], y[i]);
        CHECK(ok);
        std::vector<int32_t> a(n), b(n), c(n);
        for (size_t i = 0; i < n; ++i) a[i] = (int32_t)(rng() % 1000) - 500, b[i] = (int32_t)(rng() % 1000) - 500;
        transform2(c.data(), a.data(), b.data(), n, kI32, kMin);
        ok = true;
        for (size_t i = 0; i < n; ++i) ok = ok && c[i] == std::min(a[i], b[i]);
        CHECK(ok);
        transform2(c.data(), a.data(), b.data(), n, kI32, kProd);
        ok = true;
        for (size_t i = 0; i < n; ++i) ok = ok && c[i] == a[i] * b[i];
        CHECK(ok);
        // 16-bit floats: sum in fp32, rounded once (the reference sums bf16 as IEEE half by mistake, tensorflow/ops.h:21-22)
        std::vector<uint16_t> hx(n), hy(n), hz(n), bx(n), by(n), bz(n);
        for (size_t i = 0; i < n; ++i) {
            hx[i] = float_to_half(x[i]), hy[i] = float_to_half(y[i]);
            bx[i] = float_to_bf16(x[i]), by[i] = float_to_bf16(y[i]);
        }
        transform2(hz.data(), hx.data(), hy.data(), n, kF16, kSum);
        transform2(bz.data(), bx.data(), by.data(), n, kBF16, kSum);
        ok = true;
        for (size_t i = 0; i < n; ++i) {
            ok = ok && hz[i] == float_to_half(half_to_float(hx[i]) + half_to_float(hy[i]));
            ok = ok && bz[i] == float_to_bf16(bf16_to_float(bx[i]) + bf16_to_float(by[i]));
        }
        CHECK(ok);
    }
    CHECK(half_to_float(float_to_half(1.5f)) == 1.5f && bf16_to_float(float_to_bf16(-2.0f)) == -2.0f);
}

// ---- plan: ids, host specs, peer lists, cluster resize (plan/*_test.go) ----------------------------------------------
static void test_plan() {
    CHECK(format_ipv4(parse_ipv4("10.1.2.3")) == "10.1.2.3");
    CHECK_THROWS(parse_ipv4("10.1.2"));
    const PeerID p = PeerID::parse("192.168.1.7:10003");
    CHECK(p.port == 10003 && p.str() == "192.168.1.7:10003");
    const HostList hl = HostList::parse("10.0.0.1:4,10.0.0.2:2:pub.example");
    CHECK(hl.size() == 2 && hl.cap() == 6 && hl[1].public_addr == "pub.example");
    const PeerList pl = hl.gen_peer_list(5, PortRange{});
    CHECK(pl.size() == 5 && pl[0].port == 10000 && pl[3].port == 10003 && pl[4].ipv4 == parse_ipv4("10.0.0.2") && pl[4].port == 10000);
    CHECK_THROWS(hl.gen_peer_list(7, PortRange{}));
    CHECK(pl.rank_of(pl[4]) == 4 && pl.local_rank_of(pl[4]) == 0 && pl.local_size_of(pl[0]) == 4 && pl.host_count() == 2);
    std::vector<int> masters, parents;
    pl.partition_by_host(masters, parents);
    CHECK(masters == std::vector<int>({0, 4}) && parents == std::vector<int>({0, 0, 0, 0, 4}));
    CHECK(PeerList::parse(pl.str()) == pl);
    const HostList hf = HostList::parse_hostfile("# comment\n10.0.0.1 slots=4\n\n10.0.0.2 slots=2\n");
    CHECK(hf.size() == 2 && hf.cap() == 6);
    const PortRange pr = PortRange::parse("20000-20009");
    CHECK(pr.cap() == 10 && pr.str() == "20000-20009");
    // cluster: validate, shrink keeps the prefix, grow lands on the least-loaded host with port = max + 1
    Cluster c;
    c.runners = hl.gen_runner_list(kDefaultRunnerPort);
    c.workers = hl.gen_peer_list(3, PortRange{});
    c.validate();
    const Cluster shr = c.resized(2);
    CHECK(shr.workers.size() == 2 && shr.workers[0] == c.workers[0] && shr.workers[1] == c.workers[1]);
    const Cluster gr = c.resized(5);
    CHECK(gr.workers.size() == 5);
    CHECK(gr.workers[3].ipv4 == parse_ipv4("10.0.0.2") && gr.workers[4].ipv4 == parse_ipv4("10.0.0.2")); // host 2 had 0 workers
    gr.validate();
    CHECK(Cluster::from_json(gr.to_json()) == gr);
    CHECK(c.digest() != gr.digest());
    Cluster bad = c;
    bad.workers.push_back(bad.workers[0]); // duplicate worker
    CHECK_THROWS(bad.validate());
    Stage st{3, gr};
    CHECK(Stage::from_json(st.to_json()).version == 3 && Stage::from_json(st.to_json()).cluster == gr);
    // even partition covers [0, n) exactly
    for (size_t n : {0, 1, 10, 1023}) {
        for (size_t k : {1, 3, 8}) {
            size_t at = 0;
            bool ok = true;
            for (size_t i = 0; i < k; ++i) {
                auto be = even_partition(n, k, i);
                ok = ok && be.first == at && be.second >= be.first;
                at = be.second;
            }
            CHECK(ok && at == n);
        }
    }
}

// ---- topology generators yield rooted trees (plan/topology_test.go:14-98) ---------------------------------------------
static void test_topologies() {
    const std::vector<PeerList> shapes = {peers_on({{"10.0.0.1", 1}}), peers_on({{"10.0.0.1", 4}}), peers_on({{"10.0.0.1", 3}, {"10.0.0.2", 3}}),
                                          peers_on({{"10.0.0.1", 4}, {"10.0.0.2", 2}, {"10.0.0.3", 1}}),
                                          peers_on({{"10.0.0.1", 2}, {"10.0.0.2", 2}, {"10.0.0.3", 2}, {"10.0.0.4", 2}})};
    for (auto& pl : shapes) {
        const int n = (int)pl.size();
        for (auto& name : strategy_names()) {
            const Strategy s = parse_strategy(name);
            auto sl = gen_global_strategies(pl, s);
            CHECK(!sl.empty());
            for (auto& gp : sl) {
                CHECK(gp.bcast.size() == n && gp.reduce.size() == n);
                // the broadcast graph is a tree rooted somewhere that reaches everybody; the reduce graph is its reverse + self loops
                int roots = 0, root = -1;
                for (int i = 0; i < n; ++i)
                    if (gp.bcast.prevs(i).empty()) ++roots, root = i;
                CHECK(roots == 1);
                if (roots == 1) CHECK(gp.bcast.is_rooted_tree(root));
                // the reduce graph funnels every node into the SAME root (reversed: a tree rooted there) and every node folds
                // its own contribution in (self loop)
                bool loops = true;
                for (int i = 0; i < n; ++i) loops = loops && gp.reduce.is_self_loop(i);
                CHECK(loops);
                if (roots == 1) CHECK(gp.reduce.reversed().is_rooted_tree(root));
            }
            CHECK(digest_strategies(sl) == digest_strategies(gen_global_strategies(pl, s)));
        }
        CHECK(auto_select(pl) == (pl.host_count() == 1 ? Strategy::Star : Strategy::BinaryTreeStar));
    }
    CHECK_THROWS(parse_strategy("NO_SUCH_STRATEGY"));
    // forest arrays (set_tree): valid forest, cycle, out of range
    Graph g;
    int roots = 0;
    CHECK(Graph::from_forest_array({0, 0, 1, 1, 0}, g, roots) && roots == 1 && g.is_rooted_tree(0));
    CHECK(Graph::from_forest_array({0, 1, 0, 1}, g, roots) && roots == 2);
    CHECK(!Graph::from_forest_array({1, 0}, g, roots));
    CHECK(!Graph::from_forest_array({0, 7}, g, roots));
}

// ---- MST (test_mst.cpp): Prim on random symmetric weights equals the brute-force optimum for small n ----------------
static double tree_weight(const std::vector<std::pair<int, int>>& edges, const std::vector<double>& w, int n) {
    double t = 0;
    for (auto& e : edges) t += w[(size_t)e.first * n + e.second];
    return t;
}
static void test_mst() {
    std::mt19937 rng(11);
    for (int n = 1; n <= 7; ++n) {
        std::vector<double> w((size_t)n * n, 0.0);
        for (int i = 0; i < n; ++i)
            for (int j = i + 1; j < n; ++j) w[(size_t)i * n + j] = w[(size_t)j * n + i] = 1.0 + (double)(rng() % 1000) / 10.0;
        auto edges = minimum_spanning_tree(w, n);
        CHECK((int)edges.size() == n - 1);
        // spanning: union-find
        std::vector<int> uf(n);
        std::iota(uf.begin(), uf.end(), 0);
        std::function<int(int)> find = [&](int x) { return uf[x] == x ? x : uf[x] = find(uf[x]); };
        for (auto& e : edges) uf[find(e.first)] = find(e.second);
        bool spanning = true;
        for (int i = 0; i < n; ++i) spanning = spanning && find(i) == find(0);
        CHECK(spanning);
        // optimal: compare with Kruskal
        std::vector<std::tuple<double, int, int>> all;
        for (int i = 0; i < n; ++i)
            for (int j = i + 1; j < n; ++j) all.emplace_back(w[(size_t)i * n + j], i, j);
        std::sort(all.begin(), all.end());
        std::iota(uf.begin(), uf.end(), 0);
        double best = 0;
        for (auto& t : all) {
            int a = find(std::get<1>(t)), b = find(std::get<2>(t));
            if (a != b) uf[a] = b, best += std::get<0>(t);
        }
        CHECK(std::fabs(tree_weight(edges, w, n) - best) < 1e-9);
        auto fa = mst_forest_array(w, n, 0);
        Graph g;
        int roots = 0;
        CHECK((int)fa.size() == n && Graph::from_forest_array(fa, g, roots) && roots == 1 && fa[0] == 0);
    }
}

// ---- NUMA placement (test_placement.cpp) ----------------------------------------------------------------------------
static void test_placement() {
    std::vector<std::vector<int>> numa = {{0, 1, 2, 3, 4, 5, 6, 7}, {8, 9, 10, 11, 12, 13, 14, 15}};
    std::set<int> seen;
    for (int r = 0; r < 4; ++r) {
        auto cpus = select_cpus(numa, r, 4);
        CHECK(cpus.size() == 4);
        // contiguous slice inside ONE node, disjoint from the other ranks
        bool one_node = true;
        for (int c : cpus) one_node = one_node && (c / 8 == cpus[0] / 8);
        CHECK(one_node);
        for (int c : cpus) CHECK(seen.insert(c).second);
    }
    CHECK(seen.size() == 16);
    CHECK(select_cpus(numa, 0, 1).size() == 16 || select_cpus(numa, 0, 1).size() == 8);
    CHECK(!numa_topology().empty());
    // topology (hwloc-like) order: a permutation of the OS order per node
    auto os = numa_topology(), topo = numa_topology_ordered();
    CHECK(os.size() == topo.size());
    for (size_t i = 0; i < os.size() && i < topo.size(); ++i) {
        auto a = os[i], b = topo[i];
        std::sort(a.begin(), a.end());
        std::sort(b.begin(), b.end());
        CHECK(a == b);
    }
}

// ---- stores (store/versionedstore_test.go) ---------------------------------------------------------------------------
static void test_stores() {
    Store s;
    const char a[] = "hello", b[] = "world";
    s.save("x", a, 5);
    CHECK(s.get("x") && s.get("x")->data.size() == 5 && std::memcmp(s.get("x")->data.data(), a, 5) == 0);
    s.save("x", b, 5);
    CHECK(std::memcmp(s.get("x")->data.data(), b, 5) == 0);
    CHECK_THROWS(s.save("x", a, 3)); // size conflict
    CHECK(s.get("missing") == nullptr && s.size() == 1);
    VersionedStore vs(3);
    for (int v = 0; v < 5; ++v) vs.save(std::to_string(v), "m", a, 5);
    CHECK(vs.versions().size() == 3 && vs.get("0", "m") == nullptr && vs.get("1", "m") == nullptr && vs.get("4", "m") != nullptr);
    CHECK(vs.get("4", "other") == nullptr);
}

// ---- a Peer in single-process mode can be built and torn down twice, collectives with np=1 are identities (test_kungfu.cpp) ----
static void test_single_process_peer() {
    for (int round = 0; round < 2; ++round) {
        for (const char* k : {"KUNGFU_SELF_SPEC", "KUNGFU_INIT_PEERS", "RANK", "WORLD_SIZE", "LOCAL_RANK"}) unsetenv(k);
        PeerConfig cfg = PeerConfig::from_env(); // nothing in the environment: single-process mode
        CHECK(cfg.single);
        Peer p(cfg);
        p.start();
        auto sess = p.session();
        CHECK(sess->rank() == 0 && sess->size() == 1 && !p.detached());
        std::vector<float> x = {1.f, 2.f, 3.f}, y(3, 0.f);
        Workspace w{x.data(), y.data(), 3, kF32, kSum, "t"};
        sess->all_reduce(w);
        CHECK(y == x);
        sess->barrier();
        p.close();
    }
}

// ---- the JSON reader behind the config server and the runners' control messages ----------------------------------------------------
static void test_json_reader() {
    Cluster c = Cluster::from_json("{\"Runners\":[{\"IPv4\":2130706433,\"Port\":38080}],\"Workers\":[{\"IPv4\":2130706433,\"Port\":10000}]}");
    CHECK(c.runners.size() == 1 && c.workers.size() == 1 && c.workers[0].port == 10000);
    CHECK(Cluster::from_json(c.to_json()) == c);
    Stage st;
    st.version = 7;
    st.cluster = c;
    CHECK(Stage::from_json(st.to_json()).version == 7 && Stage::from_json(st.to_json()).cluster == c);
    for (const char* bad : {"", "{", "[1,2", "{\"Runners\":3}", "{\"Runners\":[{\"IPv4\":\"x\"}]}", "nul", "\"unterminated", "{\"a\" 1}", "[1 2]"}) {
        bool threw = false;
        try {
            (void)JsonValue::parse(bad);
            (void)Cluster::from_json(bad);
        } catch (const std::exception&) {
            threw = true;
        }
        CHECK(threw);
    }
    // depth: 64 levels are plenty for our documents; a stranger's 100 000 '[' must be an error, not a stack overflow
    CHECK_THROWS(JsonValue::parse(std::string(100000, '[')));
    std::string deep_obj;
    for (int i = 0; i < 50000; ++i) deep_obj += "{\"a\":";
    CHECK_THROWS(JsonValue::parse(deep_obj));
    std::string ok_depth = std::string(40, '[') + "1" + std::string(40, ']');
    CHECK(JsonValue::parse(ok_depth).kind == JsonValue::Array);
}

// ---- transport: stale-epoch fencing of collective connections, shutdown / reopen of the mailboxes ----------------------------------
static void test_epoch_fence() {
    const uint16_t port = (uint16_t)(41000 + getpid() % 2000);
    const PeerID srv{parse_ipv4("127.0.0.1"), port}, cli{parse_ipv4("127.0.0.1"), (uint16_t)(port + 1)};
    CollectiveEndpoint coll;
    Server server(srv, &coll, nullptr, false);
    coll.reopen(5);
    server.set_token(5);
    server.start();
    Client a(cli);
    a.set_token(5);
    const char m1[] = "epoch5";
    a.send(srv, "m", m1, sizeof(m1), kConnCollective, kNoFlag);
    Bytes got = coll.recv(cli, "m");
    CHECK(got.size() == sizeof(m1) && std::memcmp(got.data(), m1, sizeof(m1)) == 0);
    // a receiver that waits while the endpoint is shut down is woken with an error; later receives fail at once
    std::string err;
    std::thread waiter([&] {
        try {
            coll.recv(cli, "never");
        } catch (const std::exception& e) {
            err = e.what();
        }
    });
    sleep_ms(50);
    coll.shutdown("a peer of this cluster was lost");
    waiter.join();
    CHECK(err.find("lost") != std::string::npos);
    CHECK_THROWS(coll.recv(cli, "m"));
    // while closed, traffic is consumed and dropped, not parked
    a.send(srv, "m", m1, sizeof(m1), kConnCollective, kNoFlag);
    sleep_ms(50);
    CHECK(coll.parked() == 0);
    // the membership moves to epoch 6: the connection that was accepted under epoch 5 may still deliver — it must not reach a mailbox
    coll.reopen(6);
    server.set_token(6);
    a.send(srv, "m", m1, sizeof(m1), kConnCollective, kNoFlag);
    sleep_ms(50);
    CHECK(coll.parked() == 0);
    // a client that still believes in epoch 5 cannot open a new collective connection
    Client stale(PeerID{parse_ipv4("127.0.0.1"), (uint16_t)(port + 2)});
    stale.set_token(5);
    CHECK_THROWS(stale.send(srv, "m", m1, sizeof(m1), kConnCollective, kNoFlag));
    // one that moved on can
    Client b(cli);
    b.set_token(6);
    const char m2[] = "epoch6";
    b.send(srv, "m", m2, sizeof(m2), kConnCollective, kNoFlag);
    got = coll.recv(cli, "m");
    CHECK(got.size() == sizeof(m2) && std::memcmp(got.data(), m2, sizeof(m2)) == 0);
    a.close_all();
    b.close_all();
    stale.close_all();
    server.stop();
}

// ---- transport: a listener that anybody can reach must shrug off garbage (lengths come from the wire) -----------------
static void test_server_survives_garbage() {
    const uint16_t port = (uint16_t)(45000 + getpid() % 2000);
    const PeerID srv{parse_ipv4("127.0.0.1"), port}, cli{parse_ipv4("127.0.0.1"), (uint16_t)(port + 1)};
    CollectiveEndpoint coll;
    std::atomic<int> controls{0};
    Client p2p_client(srv);
    P2PEndpoint p2p(&p2p_client, srv);
    Server server(srv, &coll, &p2p, false);
    server.set_control_handler([&](const PeerID&, const std::string&, const Bytes&) { ++controls; });
    coll.reopen(1);
    server.set_token(1);
    server.start();
    std::mt19937 rng(7);
    auto raw = [&] { return Socket::connect_tcp(srv.ipv4, srv.port, 1000); };
    // 1. random bytes instead of a connection header, then silence
    for (int i = 0; i < 8; ++i) {
        Socket s = raw();
        CHECK(s.valid());
        std::vector<uint8_t> junk(1 + rng() % 64);
        for (auto& b : junk) b = (uint8_t)rng();
        s.send_all(junk.data(), junk.size());
    }
    // 2. a valid collective handshake followed by headers that announce absurd sizes: the connection is dropped, nothing is allocated
    struct __attribute__((packed)) Hdr {
        uint16_t type, src_port;
        uint32_t src_ipv4;
    };
    struct __attribute__((packed)) Msg {
        uint32_t name_len, flags;
        uint64_t data_len;
    };
    for (uint16_t type : {(uint16_t)kConnCollective, (uint16_t)kConnControl, (uint16_t)kConnPing, (uint16_t)kConnPeerToPeer, (uint16_t)77}) {
        Socket s = raw();
        Hdr h{type, cli.port, cli.ipv4};
        uint32_t ack = 0;
        CHECK(s.send_all(&h, sizeof(h)) && s.recv_all(&ack, sizeof(ack)) && ack == 1);
        Msg big{8, 0, ~0ull}; // 16 EiB of payload
        s.send_all(&big, sizeof(big));
        s.send_all("deadbeef", 8);
        Socket s2 = raw();
        CHECK(s2.send_all(&h, sizeof(h)) && s2.recv_all(&ack, sizeof(ack)));
        Msg longname{1u << 30, 0, 4}; // 1 GiB name
        s2.send_all(&longname, sizeof(longname));
    }
    { // a p2p request whose "version label" claims a gigabyte (below the generic data cap): refused, not allocated
        Socket s = raw();
        Hdr h{(uint16_t)kConnPeerToPeer, cli.port, cli.ipv4};
        uint32_t ack = 0;
        CHECK(s.send_all(&h, sizeof(h)) && s.recv_all(&ack, sizeof(ack)));
        Msg m{4, 0, 1ull << 30};
        s.send_all(&m, sizeof(m));
        s.send_all("blob", 4);
        char c;
        CHECK(!s.recv_all(&c, 1)); // the server hung up
    }
    sleep_ms(100);
    CHECK(coll.parked() == 0);
    // 3. the server still serves well-formed peers
    Client a(cli);
    a.set_token(1);
    const char m[] = "still here";
    a.send(srv, "m", m, sizeof(m), kConnCollective, kNoFlag);
    Bytes got = coll.recv(cli, "m");
    CHECK(got.size() == sizeof(m) && std::memcmp(got.data(), m, sizeof(m)) == 0);
    a.send(srv, "note", m, sizeof(m), kConnControl, kNoFlag);
    for (int i = 0; i < 50 && controls.load() == 0; ++i) sleep_ms(10);
    CHECK(controls.load() == 1);
    CHECK(a.ping(srv, 1000) >= 0);
    a.close_all();
    p2p.shutdown();
    p2p_client.close_all();
    server.stop();
}

// ---- four peers in one process, every strategy: the collectives agree with their definition (test_kungfu_world.cpp; under TSan this is
// the session / transport / thread-pool code with real concurrency and none of Python's serialisation) ---------------------------------
static void test_world_all_strategies() {
    const int n = 4;
    int round = 0;
    for (const std::string& name : strategy_names()) {
        const uint16_t base = (uint16_t)(47000 + (getpid() % 500) * 4 + 16 * (round++ % 60));
        std::vector<std::unique_ptr<Peer>> peers;
        for (int r = 0; r < n; ++r) {
            PeerConfig cfg = PeerConfig::single_machine(r, n, base);
            cfg.strategy = parse_strategy(name);
            peers.emplace_back(new Peer(cfg));
        }
        std::vector<std::string> errs(n);
        std::vector<int> ok(n, 0);
        std::vector<std::thread> ts;
        for (int r = 0; r < n; ++r)
            ts.emplace_back([&, r] {
                try {
                    peers[r]->start();
                    auto sess = peers[r]->session();
                    bool good = sess->rank() == r && sess->size() == n;
                    // all-reduce over several chunk sizes (1 element ... more than one 1 MiB chunk), sum and max
                    for (size_t cnt : {(size_t)1, (size_t)1000, (size_t)300000}) {
                        std::vector<float> x(cnt, (float)(r + 1)), y(cnt, -1.f);
                        Workspace w{x.data(), y.data(), cnt, kF32, kSum, "sum" + std::to_string(cnt)};
                        sess->all_reduce(w);
                        good = good && y[0] == 10.f && y[cnt - 1] == 10.f && y[cnt / 2] == 10.f;
                        std::vector<int32_t> a(cnt, r), b(cnt, -1);
                        Workspace w2{a.data(), b.data(), cnt, kI32, kMax, "max" + std::to_string(cnt)};
                        sess->all_reduce(w2);
                        good = good && b[0] == n - 1 && b[cnt - 1] == n - 1;
                    }
                    std::vector<double> bc(777, r == 0 ? 3.25 : -1.0), out(777, 0.0);
                    Workspace wb{bc.data(), out.data(), bc.size(), kF64, kSum, "bcast"};
                    sess->broadcast(wb);
                    good = good && out[0] == 3.25 && out[776] == 3.25;
                    std::vector<int32_t> mine(5, r), all(5 * n, -1);
                    Workspace wg{mine.data(), all.data(), mine.size(), kI32, kSum, "allgather"};
                    sess->all_gather(wg);
                    for (int q = 0; q < n; ++q) good = good && all[5 * q] == q && all[5 * q + 4] == q;
                    const char same[] = "identical", diff_a[] = "aaaa", diff_b[] = "bbbb";
                    good = good && sess->consensus(same, sizeof(same), "same");
                    good = good && !sess->consensus(r == 2 ? diff_b : diff_a, 4, "differs");
                    sess->barrier();
                    ok[r] = good ? 1 : 0;
                } catch (const std::exception& e) {
                    errs[r] = e.what();
                }
            });
        for (auto& t : ts) t.join();
        for (int r = 0; r < n; ++r) {
            if (!errs[r].empty()) std::fprintf(stderr, "strategy %s rank %d: %s\n", name.c_str(), r, errs[r].c_str());
            CHECK(errs[r].empty() && ok[r] == 1);
        }
        ts.clear();
        for (int r = 0; r < n; ++r) ts.emplace_back([&, r] { peers[r]->close(); });
        for (auto& t : ts) t.join();
    }
}

// ---- three peers in one process; one goes away; a "runner" announces it; the other two recover and keep reducing (test_kungfu_world.cpp +
// the shrink-and-continue protocol, which the reference does not have) ----------------------------------------------------------------
static void test_world_loses_a_peer() {
    const uint16_t base = (uint16_t)(43000 + getpid() % 2000);
    const int n = 3;
    std::vector<std::unique_ptr<Peer>> peers;
    for (int r = 0; r < n; ++r) peers.emplace_back(new Peer(PeerConfig::single_machine(r, n, base)));
    std::vector<std::string> errs(n);
    auto on_all = [&](std::vector<int> ranks, std::function<void(int)> f) {
        std::vector<std::thread> ts;
        for (int r : ranks)
            ts.emplace_back([&, r] {
                try {
                    f(r);
                } catch (const std::exception& e) {
                    errs[r] = e.what();
                }
            });
        for (auto& t : ts) t.join();
    };
    on_all({0, 1, 2}, [&](int r) { peers[r]->start(); });
    std::vector<float> sums(n, 0.f);
    on_all({0, 1, 2}, [&](int r) {
        float x = (float)(r + 1), y = 0;
        Workspace w{&x, &y, 1, kF32, kSum, "before"};
        peers[r]->session()->all_reduce(w);
        sums[r] = y;
    });
    CHECK(sums[0] == 6.f && sums[1] == 6.f && sums[2] == 6.f && errs[0].empty() && errs[1].empty() && errs[2].empty());
    // peer 2 disappears while 0 and 1 are inside an all-reduce that needs it
    Cluster before = peers[0]->current_cluster();
    std::thread announcer([&] {
        sleep_ms(150);
        peers[2]->close();
        Stage st;
        st.version = peers[0]->cluster_version() + 1;
        st.cluster = before;
        st.cluster.workers = before.workers.minus(PeerList{before.workers[2]});
        const std::string payload = st.to_json();
        Client runner(PeerID{parse_ipv4("127.0.0.1"), (uint16_t)(base + 100)});
        for (int r : {0, 1}) runner.send(before.workers[r], "lost", payload.data(), payload.size(), kConnControl, kNoFlag);
        runner.close_all();
    });
    on_all({0, 1}, [&](int r) {
        float x = 1.f, y = 0;
        Workspace w{&x, &y, 1, kF32, kSum, "during"};
        peers[r]->session()->all_reduce(w); // must throw, not hang
    });
    announcer.join();
    CHECK(!errs[0].empty() && !errs[1].empty());
    CHECK(peers[0]->failure_pending() && peers[1]->failure_pending());
    errs.assign(n, "");
    std::vector<int> changed(n, 0), detached(n, 0);
    on_all({0, 1}, [&](int r) {
        auto res = peers[r]->recover(10.0);
        changed[r] = res.first;
        detached[r] = res.second;
    });
    CHECK(errs[0].empty() && errs[1].empty() && changed[0] && changed[1] && !detached[0] && !detached[1]);
    CHECK(peers[0]->cluster_version() == peers[1]->cluster_version() && !peers[0]->failure_pending());
    on_all({0, 1}, [&](int r) {
        float x = (float)(r + 1), y = 0;
        Workspace w{&x, &y, 1, kF32, kSum, "after"};
        auto sess = peers[r]->session();
        sess->all_reduce(w);
        sums[r] = y + 100.f * (float)sess->size();
    });
    CHECK(errs[0].empty() && errs[1].empty() && sums[0] == 203.f && sums[1] == 203.f);
    CHECK_THROWS(peers[0]->recover(0.05)); // nothing announced: times out instead of inventing a membership
    on_all({0, 1}, [&](int r) { peers[r]->close(); });
}

// ---- launcher core: flags, worker environment, GPU slots (runner/flags_test.go, job/cuda_visible_device_test.go, proc/proc_test.go) ----
static kfrun::Flags parse_cmd(std::vector<std::string> words) {
    std::vector<char*> argv;
    words.insert(words.begin(), "kungfu-run");
    for (auto& w : words) argv.push_back(w.data());
    return kfrun::parse_flags((int)argv.size(), argv.data());
}

static void test_launcher_core() {
    using namespace kfrun;
    { // defaults of the reference: np 1, all cores of this host, runner port 38080, BINARY_TREE_STAR, NVLink allowed
        Flags f = parse_cmd({"python", "train.py", "--lr", "0.1"});
        CHECK(f.np == 1 && f.port == kDefaultRunnerPort && f.strategy == "BINARY_TREE_STAR" && !f.watch && !f.keep && f.allow_nvlink);
        CHECK(f.hosts.rfind("127.0.0.1:", 0) == 0 && f.prog == "python" && f.args.size() == 3 && f.args[2] == "0.1");
    }
    { // "-flag value", "-flag=value", "--flag", boolean flags without a value, "--" ends the launcher's flags
        Flags f = parse_cmd({"-np", "4", "-H=10.0.0.1:4,10.0.0.2:4", "--strategy", "RING", "-w", "-k", "-port-range", "20000-20100", "-timeout", "90s",
                             "-auto-recover=1m", "-delay", "500ms", "-q", "-init-version", "-1", "-builtin-config-port", "9100", "-standby", "2",
                             "-shrink-on-failure", "-refill", "-allow-nvlink=false", "-logdir", "logs", "--", "-prog-that-starts-with-a-dash", "-x"});
        CHECK(f.np == 4 && f.hosts == "10.0.0.1:4,10.0.0.2:4" && f.strategy == "RING" && f.watch && f.keep && f.quiet);
        CHECK(f.port_range.begin == 20000 && f.port_range.end == 20100 && f.timeout_s == 90 && f.auto_recover_s == 60 && f.delay_s == 0.5);
        CHECK(f.init_version == -1 && f.builtin_config_port == 9100 && f.standby == 2 && f.shrink_on_failure && f.refill && !f.allow_nvlink);
        CHECK(f.logdir == "logs" && f.prog == "-prog-that-starts-with-a-dash" && f.args.size() == 1 && f.args[0] == "-x");
    }
    // the launcher's own flags stop at the program name: "-np" AFTER it belongs to the program
    CHECK(parse_cmd({"-np", "2", "prog", "-np", "8"}).np == 2);
    CHECK_THROWS(parse_cmd({"-np"}));                        // value missing
    CHECK_THROWS(parse_cmd({"-no-such-flag", "1", "prog"})); // unknown flag
    CHECK_THROWS(parse_cmd({"-np", "2"}));                   // no program
    CHECK_THROWS(parse_cmd({"-strategy", "HYPERCUBE", "prog"}));
    try {
        parse_cmd({"-h"});
        CHECK(false);
    } catch (const UsageError& e) {
        CHECK(std::string(e.what()).empty()); // a help request is not an error
    }
    CHECK(usage_text().find("-shrink-on-failure") != std::string::npos && usage_text().find("MULTI_BINARY_TREE_STAR") != std::string::npos);

    { // smallest free GPU slot first, returned slots are reused, an exhausted pool says -1
        GpuPool pool(3);
        CHECK(pool.get() == 0 && pool.get() == 1 && pool.get() == 2 && pool.get() == -1);
        pool.put(1);
        CHECK(pool.get() == 1);
        pool.put(7); // out of range: ignored
        pool.put(0);
        pool.put(2);
        CHECK(pool.get() == 0 && pool.get() == 2);
        CHECK(GpuPool(0).get() == -1 && GpuPool(-3).get() == -1);
    }
    { // a slot is an index INTO the launcher's CUDA_VISIBLE_DEVICES (ids or UUIDs), or the slot itself when the variable is unset
        const char* old = getenv("CUDA_VISIBLE_DEVICES");
        const std::string saved = old ? old : "";
        unsetenv("CUDA_VISIBLE_DEVICES");
        CHECK(Job::cuda_index(0) == "0" && Job::cuda_index(5) == "5");
        setenv("CUDA_VISIBLE_DEVICES", "4,6,GPU-8f3c", 1);
        CHECK(Job::cuda_index(0) == "4" && Job::cuda_index(1) == "6" && Job::cuda_index(2) == "GPU-8f3c" && Job::cuda_index(3) == "-1" && Job::cuda_index(-1) == "-1");
        // the environment a worker is started with: the runner -> worker contract (env/envs.go)
        Job job;
        job.strategy = "RING";
        job.prog = "python";
        job.args = {"train.py"};
        job.parent = PeerID{parse_ipv4("10.0.0.1"), 38080};
        job.hosts = HostList::parse("10.0.0.1:2,10.0.0.2:2");
        job.config_server = "http://10.0.0.1:9100/config";
        job.start_time = 1234;
        Cluster c;
        c.runners = PeerList::parse("10.0.0.1:38080,10.0.0.2:38080");
        c.workers = PeerList::parse("10.0.0.1:10000,10.0.0.1:10001,10.0.0.2:10000");
        auto procs = job.create_procs(c, parse_ipv4("10.0.0.1"), 3);
        CHECK(procs.size() == 2 && procs[0].name == "10.0.0.1.10000" && procs[1].name == "10.0.0.1.10001" && procs[1].log_prefix == "10.0.0.1.10001@3");
        auto& e = procs[1].envs;
        CHECK(e[envkeys::kSelfSpec] == "10.0.0.1:10001" && e[envkeys::kInitPeers] == "10.0.0.1:10000,10.0.0.1:10001,10.0.0.2:10000");
        CHECK(e[envkeys::kInitRunners] == "10.0.0.1:38080,10.0.0.2:38080" && e[envkeys::kParentID] == "10.0.0.1:38080");
        CHECK(e[envkeys::kInitClusterVersion] == "3" && e[envkeys::kAllReduceStrategy] == "RING" && e[envkeys::kJobStart] == "1234");
        CHECK(e[envkeys::kConfigServer] == "http://10.0.0.1:9100/config" && e[envkeys::kCudaVisibleDevices] == "6" && e[envkeys::kAllowNvlink] == "1");
        CHECK(e.count("CUDA_VISIBLE_DEVICES") == 0); // every GPU stays visible: peer-to-peer and multicast need it
        job.allow_nvlink = false;
        auto hidden = job.create_procs(c, parse_ipv4("10.0.0.1"), 3);
        CHECK(hidden[0].envs["CUDA_VISIBLE_DEVICES"] == "4" && hidden[0].envs.count(envkeys::kAllowNvlink) == 0);
        CHECK(job.create_procs(c, parse_ipv4("10.0.0.9"), 3).empty());
        if (old) setenv("CUDA_VISIBLE_DEVICES", saved.c_str(), 1);
        else unsetenv("CUDA_VISIBLE_DEVICES");
    }
    CHECK(infer_self_ipv4("127.0.0.7", "") == parse_ipv4("127.0.0.7") && infer_self_ipv4("", "") == parse_ipv4("127.0.0.1"));
    CHECK(infer_self_ipv4("", "lo") == parse_ipv4("127.0.0.1"));
    CHECK_THROWS(infer_self_ipv4("", "no-such-nic0"));
}

// ---- start-up retry rule (runner/local/hack.go), the "-flag value" grammar of the tools, fake model lists (tests/go/fakemodel) ------------
static void test_tools_core() {
    using namespace kfrun;
    const std::string bug = loader_bug_prefix();
    CHECK(restart_after_startup_failure(bug + ": dl-open.c: 4711\n", 127, 1, bug, 10));
    CHECK(restart_after_startup_failure(bug, 1, 10, bug, 10));
    CHECK(!restart_after_startup_failure(bug, 1, 11, bug, 10));                          // retries used up
    CHECK(!restart_after_startup_failure(bug, 0, 1, bug, 10));                           // it finished fine: nothing to do
    CHECK(!restart_after_startup_failure("Traceback (most recent call last):\n", 1, 1, bug, 10));
    CHECK(!restart_after_startup_failure("note: " + bug, 1, 1, bug, 10));                // only as the FIRST thing on stderr
    CHECK(!restart_after_startup_failure("", 1, 1, bug, 10) && !restart_after_startup_failure(bug, 1, 1, "", 10));
    CHECK(!restart_after_startup_failure(bug, 1, 1, bug, 0));                            // switched off
    {
        const char* av[] = {"tool", "-model", "vgg16", "--mode=par", "-fuse", "-epochs", "7", "-rate=0.25", "-versioned=true", "-n"};
        kftools::Flags fl{10, const_cast<char**>(av)};
        CHECK(fl.str("model", "x") == "vgg16" && fl.str("mode", "seq") == "par" && fl.str("absent", "dflt") == "dflt");
        CHECK(fl.num("epochs", 1) == 7 && fl.num("warmup", 2) == 2 && fl.str("rate", "") == "0.25");
        CHECK(fl.flag("fuse") && fl.flag("versioned") && !fl.flag("shuffle"));
        CHECK(fl.str("n", "none") == "none"); // a trailing flag without a value keeps its default
    }
    // the lists are derived from layer shapes: totals must be the architectures' parameter counts
    CHECK(kftools::fake_model_sizes("resnet50").size() == 161 && kftools::total_elements(kftools::fake_model_sizes("resnet50")) == 25557032u);
    CHECK(kftools::fake_model_sizes("vgg16").size() == 32 && kftools::total_elements(kftools::fake_model_sizes("vgg16")) == 138357544u);
    CHECK(kftools::fake_model_sizes("slp").size() == 2 && kftools::total_elements(kftools::fake_model_sizes("slp")) == 7850u);
    CHECK(kftools::fake_model_sizes("resnet50-keras").size() == 214 && kftools::fake_model_sizes("bert").size() == 599);
    CHECK(kftools::fake_model_sizes("bert-base").size() == 199 && kftools::fake_model_sizes("bert-large").size() == 391);
    CHECK(kftools::fake_model_sizes("alexnet").empty());
    for (auto& n : kftools::fake_model_names()) CHECK(!kftools::fake_model_sizes(n).empty());
}

// ---- traffic counters and the metrics text (monitor/counters_test.go), duration / env helpers (utils/utils_test.go) --------------------
static void test_counters_and_utils() {
    NetMonitor& m = NetMonitor::get();
    m.reset();
    const PeerID a{parse_ipv4("10.1.0.1"), 10000}, b{parse_ipv4("10.1.0.2"), 10001};
    m.egress(a, 1000);
    m.egress(a, 24);
    m.egress(b, 7);
    m.ingress(b, 512);
    auto tot = m.egress_totals();
    CHECK(tot.size() == 2);
    uint64_t sum = 0;
    for (auto& kv : tot) sum += kv.second;
    CHECK(sum == 1031);
    PeerList pl;
    pl.push_back(a);
    pl.push_back(b);
    (void)m.egress_rates(pl); // first query sets the reference point
    sleep_ms(30);
    m.egress(a, 3000);
    auto rates = m.egress_rates(pl);
    CHECK(rates.size() == 2 && rates[0] > 0 && rates[1] == 0);
    m.set_gauge("kungfu_gradient_noise_scale", 12.5);
    m.set_gauge("kungfu_gradient_noise_scale", 13.5); // overwritten
    m.add_counter("gpu_collective_bytes_total{op=\"all_reduce\"}", 100);
    m.add_counter("gpu_collective_bytes_total{op=\"all_reduce\"}", 28);
    const std::string text = m.metrics_text();
    CHECK(text.find("egress_total_bytes{peer=\"10.1.0.1:10000\"} 4024") != std::string::npos);
    CHECK(text.find("ingress_total_bytes{peer=\"10.1.0.2:10001\"} 512") != std::string::npos);
    CHECK(text.find("kungfu_gradient_noise_scale 13.5") != std::string::npos);
    CHECK(text.find("gpu_collective_bytes_total{op=\"all_reduce\"} 128") != std::string::npos);
    m.reset();
    CHECK(m.egress_totals().empty());

    CHECK(parse_duration_s("10s", 0) == 10 && parse_duration_s("500ms", 0) == 0.5 && parse_duration_s("2m", 0) == 120 && parse_duration_s("1.5", 0) == 1.5);
    CHECK(parse_duration_s("", 7) == 7 && parse_duration_s("soon", 7) == 7);
    setenv("KF_UNIT_BOOL", "true", 1);
    CHECK(getenv_bool("KF_UNIT_BOOL", false) && !getenv_bool("KF_UNIT_UNSET_BOOL", false) && getenv_bool("KF_UNIT_UNSET_BOOL", true));
    CHECK(getenv_str("KF_UNIT_UNSET", "dflt") == "dflt");
    CHECK(merge_errors({"", ""}, "x").empty());
    const std::string merged = merge_errors({"", "boom", "bang"}, "2 tasks");
    CHECK(merged.find("boom") != std::string::npos && merged.find("bang") != std::string::npos);
    const double t0 = now_s();
    sleep_ms(20);
    CHECK(now_s() - t0 >= 0.015 && unix_time_ns() > 1600000000ull * 1000000000ull);
    CHECK(strprintf("%d-%s", 7, "x") == "7-x");
}

// ---- framed messages end to end (rchannel/connection/message_test.go: header / message round trips up to 16 MiB) ------------------------
static void test_message_round_trip() {
    const uint16_t port = (uint16_t)(45000 + getpid() % 2000);
    const PeerID srv{parse_ipv4("127.0.0.1"), port}, cli{parse_ipv4("127.0.0.1"), (uint16_t)(port + 1)};
    for (bool use_unix : {false, true}) {
        CollectiveEndpoint coll;
        Server server(srv, &coll, nullptr, use_unix);
        std::vector<std::pair<std::string, size_t>> control;
        std::mutex mu;
        server.set_control_handler([&](const PeerID& src, const std::string& name, const Bytes& payload) {
            std::lock_guard<std::mutex> g(mu);
            if (src == cli) control.emplace_back(name, payload.size());
        });
        server.start();
        Client c(cli);
        CHECK(c.ping(srv) >= 0 && c.wait_peer(srv, 2.0));
        std::mt19937 rng(11);
        for (size_t len : {(size_t)0, (size_t)1, (size_t)4095, (size_t)(1 << 20) + 3, (size_t)16 << 20}) {
            std::vector<uint8_t> msg(len);
            for (size_t i = 0; i < len; i += 257) msg[i] = (uint8_t)rng();
            if (len) msg[len - 1] = 0xA5;
            const std::string name = "blob-" + std::to_string(len) + std::string(len % 7, 'n'); // names of different lengths
            c.send(srv, name, msg.data(), len, kConnCollective, kNoFlag);
            std::vector<uint8_t> got(len, 0xFF);
            coll.recv_into(cli, name, got.data(), len); // straight into the receiver's buffer
            CHECK(got == msg);
        }
        // a receive that expects another size than what was sent is an error, not a truncated copy
        const char five[5] = {1, 2, 3, 4, 5};
        c.send(srv, "five", five, 5, kConnCollective, kNoFlag);
        char four[4];
        CHECK_THROWS(coll.recv_into(cli, "five", four, 4));
        c.send(srv, "update", "{}", 2, kConnControl, kNoFlag);
        for (int i = 0; i < 200; ++i) {
            {
                std::lock_guard<std::mutex> g(mu);
                if (!control.empty()) break;
            }
            sleep_ms(5);
        }
        {
            std::lock_guard<std::mutex> g(mu);
            CHECK(control.size() == 1 && control[0].first == "update" && control[0].second == 2);
        }
        c.close_all();
        server.stop();
        CHECK(c.ping(srv, 200) < 0); // nobody listens any more
    }
}

int main(int argc, char** argv) {
    // read once, before any thread exists: a refused dial fails fast in these tests (50 x 20 ms instead of 500 x 200 ms)
    setenv("KUNGFU_CONFIG_CONN_RETRY_COUNT", "50", 1);
    setenv("KUNGFU_CONFIG_CONN_RETRY_PERIOD", "20ms", 1);
    struct {
        const char* name;
        void (*fn)();
    } tests[] = {{"dtypes_and_transform2", test_dtypes_and_transform2}, {"plan", test_plan}, {"topologies", test_topologies}, {"mst", test_mst},
                 {"placement", test_placement}, {"stores", test_stores}, {"single_process_peer", test_single_process_peer},
                 {"json_reader", test_json_reader}, {"epoch_fence", test_epoch_fence}, {"server_survives_garbage", test_server_survives_garbage}, {"world_all_strategies", test_world_all_strategies},
                 {"world_loses_a_peer", test_world_loses_a_peer}, {"launcher_core", test_launcher_core}, {"tools_core", test_tools_core},
                 {"counters_and_utils", test_counters_and_utils}, {"message_round_trip", test_message_round_trip}};
    for (auto& t : tests) {
        bool wanted = argc <= 1; // `kungfu-unit-tests name…` runs the named tests only
        for (int i = 1; i < argc; ++i) wanted = wanted || std::string(argv[i]) == t.name;
        if (!wanted) continue;
        const int before = g_fail;
        try {
            t.fn();
        } catch (const std::exception& e) {
            ++g_fail;
            std::fprintf(stderr, "FAIL %s: uncaught exception: %s\n", t.name, e.what());
        }
        std::printf("[%s] %s\n", g_fail == before ? " OK " : "FAIL", t.name);
    }
    std::printf("%d checks, %d failures\n", g_checks, g_fail);
    return g_fail ? 1 : 0;
}
