// kungfu-test-public-apis — the public C++ API of the runtime driven the way an application would, under `kungfu-run -np N`.
//
//   kungfu-run -np 4 kungfu-test-public-apis          (add KUNGFU_CONFIG_ENABLE_MONITORING=true to include the /metrics checks)
//
// Reference: tests/go/cmd/kungfu-test-public-apis/kungfu-test-public-apis.go (all-reduce x20 incl. a 1-byte tensor and MAX, all-reduce
// along a given tree over 1 Mi ints, all-gather x10, peer latencies, P2P save / request incl. the negative test at :164-170) and
// tests/go/cmd/kungfu-test-monitor/kungfu-test-monitor.go (traffic, then the worker's own metrics endpoint is read back).
// Every result is checked against its definition; a failed check names the line and exits 3.
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "base.hpp"
#include "graph.hpp"
#include "http.hpp"
#include "peer.hpp"

using namespace kfrt;

#define CHECK(cond)                                                                                       \
    do {                                                                                                  \
        if (!(cond)) {                                                                                    \
            std::fprintf(stderr, "CHECK FAILED %s:%d: %s (rank %d)\n", __FILE__, __LINE__, #cond, g_rank); \
            std::exit(3);                                                                                 \
        }                                                                                                 \
    } while (0)

static int g_rank = -1;

int main() {
    try {
        Peer peer(PeerConfig::from_env());
        peer.start();
        auto sess = peer.session();
        const int np = sess->size(), rank = g_rank = sess->rank();
        CHECK(rank >= 0 && rank < np && sess->local_rank() < sess->local_size() && sess->host_count() >= 1);
        CHECK(peer.uid() != 0 && !peer.detached());

        // all-reduce x20: sizes from one byte to past the chunk size, SUM and MAX, in place and out of place
        for (int i = 0; i < 20; ++i) {
            const size_t cnt = i == 0 ? 1 : (size_t)1 << i; // 1 ... 512 Ki elements
            std::vector<uint8_t> b(i == 0 ? 1 : 3, 1), bo(b.size(), 0);
            Workspace wb{b.data(), bo.data(), b.size(), kU8, kSum, "u8-" + std::to_string(i)};
            sess->all_reduce(wb);
            CHECK(bo[0] == (uint8_t)np);
            std::vector<int32_t> x(cnt, rank + 1), y(cnt, 0);
            Workspace ws{x.data(), y.data(), cnt, kI32, kSum, "sum-" + std::to_string(i)};
            sess->all_reduce(ws);
            CHECK(y[0] == np * (np + 1) / 2 && y[cnt - 1] == np * (np + 1) / 2);
            Workspace wm{x.data(), x.data(), cnt, kI32, kMax, "max-" + std::to_string(i)}; // in place
            sess->all_reduce(wm);
            CHECK(x[0] == np && x[cnt / 2] == np);
        }

        // all-reduce along an explicit tree (a chain: i's father is i-1) over 1 Mi ints
        {
            std::vector<int32_t> tree(np);
            for (int i = 0; i < np; ++i) tree[i] = i == 0 ? 0 : i - 1;
            std::vector<int32_t> x((size_t)1 << 20, 1), y(x.size(), 0);
            Workspace w{x.data(), y.data(), x.size(), kI32, kSum, "with-tree"};
            sess->monitored_all_reduce(w, tree);
            CHECK(y[0] == np && y.back() == np);
            sess->set_tree(tree); // consensus-checked switch of the strategy, then a plain all-reduce over it
            Workspace w2{x.data(), y.data(), 1000, kI32, kSum, "after-set-tree"};
            sess->all_reduce(w2);
            sess->calc_stats(); // one active strategy now: its throughput window is closed without complaint
            CHECK(y[0] == np && y[999] == np);
        }

        // all-gather x10, gather, broadcast, reduce, local / cross variants
        for (int i = 0; i < 10; ++i) {
            const size_t cnt = 1 + 37 * i;
            std::vector<int64_t> mine(cnt, rank), all(cnt * np, -1);
            Workspace w{mine.data(), all.data(), cnt, kI64, kSum, "ag-" + std::to_string(i)};
            sess->all_gather(w);
            for (int q = 0; q < np; ++q) CHECK(all[q * cnt] == q && all[q * cnt + cnt - 1] == q);
        }
        {
            std::vector<float> mine(5, (float)rank), all(5 * np, -1.f);
            Workspace w{mine.data(), all.data(), 5, kF32, kSum, "gather"};
            sess->gather(w);
            if (rank == 0)
                for (int q = 0; q < np; ++q) CHECK(all[5 * q] == (float)q);
            std::vector<double> b(100, rank == 0 ? 2.5 : 0.0), o(100, 0.0);
            Workspace wb{b.data(), o.data(), 100, kF64, kSum, "bcast"};
            sess->broadcast(wb);
            CHECK(o[0] == 2.5 && o[99] == 2.5);
            std::vector<float> one(10, 1.f), red(10, 0.f);
            Workspace wr{one.data(), red.data(), 10, kF32, kSum, "reduce"};
            sess->reduce(wr);
            if (rank == 0) CHECK(red[0] == (float)np);
            Workspace wl{one.data(), red.data(), 10, kF32, kSum, "local-reduce"};
            sess->local_reduce(wl);
            Workspace wlb{red.data(), red.data(), 10, kF32, kSum, "local-bcast"};
            sess->local_broadcast(wlb);
            CHECK(red[0] == (float)sess->local_size());
            Workspace wc{one.data(), red.data(), 10, kF32, kSum, "cross"};
            sess->cross_all_reduce(wc);
            if (sess->local_rank() == 0) CHECK(red[0] == (float)sess->host_count());
        }

        // consensus: true for identical bytes, false when one worker differs (unless alone)
        {
            const std::string same = "same";
            CHECK(sess->consensus(same.data(), same.size(), "c1"));
            const std::string mine = rank == np - 1 ? "x" : "y";
            CHECK(sess->consensus(mine.data(), mine.size(), "c2") == (np == 1));
        }

        // latencies and egress rates: one number per peer, 0 for self
        {
            auto lat = sess->peer_latencies();
            CHECK((int)lat.size() == np && lat[rank] == 0.0);
            for (int q = 0; q < np; ++q) CHECK(lat[q] >= 0.0 && lat[q] < 5.0);
            CHECK((int)sess->egress_rates().size() == np);
        }

        // peer-to-peer store: plain and versioned, window of three versions, unknown names / versions answer "not there"
        {
            std::vector<float> blob(1000, (float)(rank + 1)), got(1000, 0.f);
            peer.save("weights", blob.data(), blob.size() * 4);
            for (int v = 0; v < 5; ++v) {
                std::vector<float> bv(1000, (float)(100 * v + rank));
                peer.save_version(std::to_string(v), "weights", bv.data(), bv.size() * 4);
            }
            sess->barrier();
            const int target = (rank + 1) % np;
            CHECK(peer.request(target, "", "weights", got.data(), got.size() * 4));
            CHECK(got[0] == (float)(target + 1) && got[999] == (float)(target + 1));
            CHECK(peer.request(target, "4", "weights", got.data(), got.size() * 4));
            CHECK(got[0] == (float)(400 + target));
            CHECK(peer.request(target, "2", "weights", got.data(), got.size() * 4));
            CHECK(got[0] == (float)(200 + target));
            CHECK(!peer.request(target, "0", "weights", got.data(), got.size() * 4)); // evicted: only the last three versions are kept
            CHECK(!peer.request(target, "", "never-saved", got.data(), got.size() * 4));
            CHECK(!peer.request(target, "9", "weights", got.data(), got.size() * 4));
            sess->barrier();
        }

        // monitoring endpoint (when enabled): the traffic above shows up in this worker's own counters
        if (Config::get().enable_monitoring && np > 1) {
            std::string body;
            const std::string url = "http://" + format_ipv4(peer.self().ipv4) + ":" + std::to_string(peer.self().port + 10000) + "/metrics";
            const int st = http_request("GET", url, "", &body);
            CHECK(st == 200);
            CHECK(body.find("egress_total_bytes") != std::string::npos);
            if (rank == 0) std::printf("metrics endpoint %s: %zu bytes\n", url.c_str(), body.size());
        }

        // MST over measured latencies is a tree over all peers (what adapt_strategy feeds to set_tree)
        {
            std::vector<double> w((size_t)np * np, 0.0);
            for (int i = 0; i < np; ++i)
                for (int j = 0; j < np; ++j) w[i * np + j] = i == j ? 0.0 : 1.0 + ((i * 7 + j * 13) % 5);
            auto father = mst_forest_array(w, np);
            CHECK((int)father.size() == np && father[0] == 0);
        }
        sess->barrier();
        if (rank == 0) std::printf("test-public-apis ok np=%d\n", np);
        peer.close();
    } catch (const std::exception& e) {
        std::fprintf(stderr, "kungfu-test-public-apis: %s\n", e.what());
        return 1;
    }
    return 0;
}
