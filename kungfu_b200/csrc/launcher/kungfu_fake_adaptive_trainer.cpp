// kungfu-fake-adaptive-trainer — a trainer with no model that exercises the elastic protocol natively: every step all-reduces a few
// tensors, then asks for a resize; after a membership change the survivors and the newcomers agree on the step to continue from.
//
//   kungfu-run -w -np 1 -H 127.0.0.1:8 -builtin-config-port 9100 kungfu-fake-adaptive-trainer -max-step 30 -schedule 3:2,6:4,9:8,12:4,15:1
//   kungfu-run -w -np 2 -H 127.0.0.1:4 -config-server http://…/get kungfu-fake-adaptive-trainer -max-step 100      (resizes come from outside)
//
// Reference: tests/go/cmd/kungfu-fake-adaptive-trainer/kungfu-fake-adaptive-trainer.go:29-131 (train step, `ResizeClusterFromURL`
// after every step, `syncStep` = all-reduce MAX of the step counter after a change; rank 0 sleeps longer than the others to provoke
// skew). Additions: `-schedule step:size,…` drives the resizes from inside (rank 0 proposes), and every step checks that the reduced
// value equals the cluster size of that step, so a worker that trains with a stale membership is caught.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <string>
#include <vector>

#include "base.hpp"
#include "fake_model.hpp"
#include "peer.hpp"

using namespace kfrt;

static std::map<int, int> parse_schedule(const std::string& s) {
    std::map<int, int> out;
    size_t i = 0;
    while (i < s.size()) {
        size_t j = s.find(',', i);
        if (j == std::string::npos) j = s.size();
        const std::string item = s.substr(i, j - i);
        const size_t c = item.find(':');
        if (c != std::string::npos) out[std::atoi(item.substr(0, c).c_str())] = std::atoi(item.substr(c + 1).c_str());
        i = j + 1;
    }
    return out;
}

static int sync_step(Session& sess, int step) {
    int64_t x = step, y = 0;
    Workspace w{&x, &y, 1, kI64, kMax, "sync-step"};
    sess.all_reduce(w);
    return (int)y;
}

int main(int argc, char** argv) {
    kftools::Flags fl{argc, argv};
    const int max_step = (int)fl.num("max-step", 10);
    const bool train = fl.str("train", "true") != "false";
    const auto schedule = parse_schedule(fl.str("schedule", ""));
    const int slow_ms = (int)fl.num("slow-ms", 60), fast_ms = (int)fl.num("fast-ms", 10);
    auto sizes = kftools::fake_model_sizes(fl.str("model", "slp"));
    if (sizes.empty()) sizes = {1};
    const size_t total = kftools::total_elements(sizes);
    try {
        Peer peer(PeerConfig::from_env());
        peer.start();
        std::vector<float> send(total, 1.f), recv(total, 0.f);
        bool should_sync = true; // newcomers start at 0 and learn the step from the survivors
        int resizes = 0, checks = 0;
        for (int step = 0; step < max_step; ++step) {
            auto sess = peer.session();
            if (should_sync) {
                const int agreed = sync_step(*sess, step);
                std::printf("sync step: %d -> %d (rank %d of %d, cluster version %d)\n", step, agreed, sess->rank(), sess->size(), peer.cluster_version());
                step = agreed;
                should_sync = false;
                if (step >= max_step) break;
            }
            if (train) {
                const double t0 = now_s();
                size_t off = 0;
                for (size_t i = 0; i < sizes.size(); ++i) {
                    Workspace w{send.data() + off, recv.data() + off, sizes[i], kF32, kSum,
                                "g" + std::to_string(i) + "v" + std::to_string(peer.cluster_version())};
                    sess->all_reduce(w);
                    off += sizes[i];
                }
                if ((int)recv[0] != sess->size() || (int)recv[total - 1] != sess->size()) {
                    std::fprintf(stderr, "step %d: reduced %g, cluster size %d\n", step, recv[0], sess->size());
                    return 3;
                }
                ++checks;
                std::printf("step: %d, rank=%d, np=%d, took %.2f ms\n", step, sess->rank(), sess->size(), (now_s() - t0) * 1e3);
                sleep_ms(sess->rank() == 0 ? slow_ms : fast_ms); // skew on purpose: rank 0 is the straggler
            }
            const int old_size = sess->size(), old_rank = sess->rank();
            const double t0 = now_s();
            std::pair<bool, bool> r{false, false};
            auto it = schedule.find(step);
            if (it != schedule.end() && it->second != old_size)
                r = peer.resize_cluster(it->second); // rank 0 proposes, everybody adopts
            else if (!peer.config().config_server.empty() && schedule.empty())
                r = peer.resize_cluster_from_url();
            if (r.first) {
                ++resizes;
                if (r.second) {
                    std::printf("resize took %.1f ms, rank %d is no longer in the cluster of %d peers\n", (now_s() - t0) * 1e3, old_rank, old_size);
                    break;
                }
                auto ns = peer.session();
                std::printf("resize %d -> %d took %.1f ms, rank %d -> %d\n", old_size, ns->size(), (now_s() - t0) * 1e3, old_rank, ns->rank());
                should_sync = true;
            }
        }
        std::printf("finished: rank %d%s, %d resizes seen, %d steps checked\n", peer.detached() ? -1 : peer.session()->rank(),
                    peer.detached() ? " (detached)" : "", resizes, checks);
        std::fflush(stdout);
        peer.close();
    } catch (const std::exception& e) {
        std::fprintf(stderr, "kungfu-fake-adaptive-trainer: %s\n", e.what());
        return 1;
    }
    return 0;
}
