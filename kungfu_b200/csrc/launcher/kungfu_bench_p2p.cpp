// kungfu-bench-p2p — model-shaped micro-benchmark of the peer-to-peer blob store (the transport under PairAveraging on the host plane).
//
//   kungfu-run -np 4 kungfu-bench-p2p [-model resnet50|vgg16|bert|slp] [-mode seq|par] [-fuse] [-epochs 15] [-warmup 2] [-versioned]
//
// Reference: tests/go/cmd/kungfu-bench-p2p/kungfu-bench-p2p.go:18-131 (every worker saves the model's tensors into its store, picks the
// next rank as its target and requests all tensors from it, one after the other or all at once; rate = epochs * 2 np * model bytes / t).
// Written against the C++ Peer directly (not the C ABI), with a verification pass the reference does not have: the payload of worker r
// is filled with r + 1, so a pull that landed in the wrong buffer or came from the wrong peer is caught.
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <thread>
#include <vector>

#include "base.hpp"
#include "fake_model.hpp"
#include "peer.hpp"

using namespace kfrt;

static std::string show_rate(double bytes, double seconds) {
    const double r = bytes / std::max(seconds, 1e-9);
    const char* units[] = {"B/s", "KiB/s", "MiB/s", "GiB/s", "TiB/s"};
    int u = 0;
    double v = r;
    while (v >= 1024.0 && u < 4) v /= 1024.0, ++u;
    return strprintf("%.2f %s", v, units[u]);
}

int main(int argc, char** argv) {
    kftools::Flags fl{argc, argv};
    const std::string model = fl.str("model", "resnet50"), mode = fl.str("mode", "seq");
    const bool fuse = fl.flag("fuse"), versioned = fl.flag("versioned");
    const int epochs = (int)fl.num("epochs", 15), warmup = (int)fl.num("warmup", 2);
    auto sizes = kftools::fake_model_sizes(model);
    if (sizes.empty()) {
        std::fprintf(stderr, "invalid model name: %s (", model.c_str());
        for (auto& n : kftools::fake_model_names()) std::fprintf(stderr, " %s", n.c_str());
        std::fprintf(stderr, " )\n");
        return 2;
    }
    if (mode != "seq" && mode != "par") {
        std::fprintf(stderr, "invalid mode: %s (seq | par)\n", mode.c_str());
        return 2;
    }
    if (fuse) sizes = {kftools::total_elements(sizes)};
    int rc = 0;
    try {
        Peer peer(PeerConfig::from_env());
        peer.start();
        auto sess = peer.session();
        const int np = sess->size(), rank = sess->rank();
        const size_t total = kftools::total_elements(sizes);
        if (rank == 0)
            std::printf("model: %s, %zu tensors, %zu parameters, %.2f MiB, mode: %s%s\n", model.c_str(), sizes.size(), total, total * 4.0 / (1 << 20),
                        mode.c_str(), versioned ? ", versioned store" : "");
        // one send / one receive buffer per tensor, views of two flat allocations
        std::vector<float> send(total, (float)(rank + 1)), recv(total, 0.f);
        std::vector<size_t> offs(sizes.size());
        for (size_t i = 0, o = 0; i < sizes.size(); ++i) offs[i] = o, o += sizes[i];
        auto name_of = [](size_t i) { return "t" + std::to_string(i); };
        const std::string version = versioned ? "0" : "";
        for (size_t i = 0; i < sizes.size(); ++i) {
            if (versioned)
                peer.save_version(version, name_of(i), send.data() + offs[i], sizes[i] * 4);
            else
                peer.save(name_of(i), send.data() + offs[i], sizes[i] * 4);
        }
        sess->barrier();
        const int target = np > 1 ? (rank + 1) % np : rank;
        std::atomic<int> missed{0};
        auto pull = [&](size_t i) {
            if (!peer.request(target, version, name_of(i), recv.data() + offs[i], sizes[i] * 4)) missed.fetch_add(1);
        };
        auto run_epoch = [&] {
            if (mode == "seq") {
                for (size_t i = 0; i < sizes.size(); ++i) pull(i);
            } else {
                // all requests in flight at once: a small crew of threads drains an index counter (the store serves them concurrently)
                std::atomic<size_t> next{0};
                const int crew = (int)std::min<size_t>(sizes.size(), 16);
                std::vector<std::thread> ts;
                for (int t = 0; t < crew; ++t)
                    ts.emplace_back([&] {
                        for (size_t i; (i = next.fetch_add(1)) < sizes.size();) pull(i);
                    });
                for (auto& t : ts) t.join();
            }
        };
        for (int e = 0; e < warmup; ++e) {
            run_epoch();
            sess->barrier();
        }
        // what arrived is the target's payload
        bool good = missed.load() == 0;
        for (size_t i = 0; i < sizes.size() && good; ++i)
            good = recv[offs[i]] == (float)(target + 1) && recv[offs[i] + sizes[i] - 1] == (float)(target + 1);
        const double t0 = now_s();
        for (int e = 0; e < epochs; ++e) {
            run_epoch();
            sess->barrier();
        }
        const double dt = now_s() - t0;
        const double workload = (double)epochs * 2.0 * np * total * 4.0; // the reference's accounting: request + response, all workers
        { // a request for a name nobody saved answers "not there" without an error (kungfu-test-public-apis.go:164-170)
            float probe = 0;
            if (peer.request(target, version, "no-such-tensor", &probe, 4)) good = false;
        }
        int64_t bad = good ? 0 : 1, bad_all = 0;
        Workspace w{&bad, &bad_all, 1, kI64, kSum, "bench-p2p-verdict"};
        sess->all_reduce(w);
        if (rank == 0) {
            std::printf("took %.3fs, rate: %s\n", dt, show_rate(workload, dt).c_str());
            std::printf("Result: model: %s, %zu tensors, mode: %s, np: %d, rate: %s, %.3f ms per pulled model%s\n", model.c_str(), sizes.size(),
                        mode.c_str(), np, show_rate(workload, dt).c_str(), dt / epochs * 1e3, bad_all ? "  [VERIFICATION FAILED]" : "");
        }
        if (bad_all) rc = 3;
        sess->barrier();
        peer.close();
    } catch (const std::exception& e) {
        std::fprintf(stderr, "kungfu-bench-p2p: %s\n", e.what());
        return 1;
    }
    return rc;
}
