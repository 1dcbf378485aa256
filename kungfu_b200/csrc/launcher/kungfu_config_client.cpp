// kungfu-config-client — drive an elastic job from outside by rewriting the cluster document on the config server.
//
//   kungfu-config-client -url http://127.0.0.1:9100/config                       print the current size
//   kungfu-config-client -url … -size 4                                          one resize (GET, resize, PUT)
//   kungfu-config-client -url … -random -min 1 -max 8 -period 1s -duration 30s [-final 2] [-seed 7] [-shuffle]
//   kungfu-config-client -url … -clear                                           DELETE: workers see an empty cluster and detach
//
// Reference: tests/go/cmd/kungfu-config-client-example/kungfu-config-client-example.go:96-116 — the chaos half of the reference's
// docker-compose elastic test: every period it PUTs a random-size subset of the workers. Shrinking keeps a prefix and growing adds to the
// least loaded runner host (`Cluster::resized`, the same rule `resize_cluster(n)` applies inside a job); `-shuffle` re-draws the kept
// workers at random like the reference's client does, which also exercises rank changes of the survivors.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <string>
#include <thread>

#include "base.hpp"
#include "fake_model.hpp"
#include "http.hpp"
#include "plan.hpp"

using namespace kfrt;

static bool get_cluster(const std::string& url, Cluster* c) {
    std::string body;
    if (http_request("GET", url, "", &body) != 200) return false;
    try {
        *c = Cluster::from_json(body);
    } catch (const std::exception&) {
        return false;
    }
    return true;
}

static bool put_size(const std::string& url, int size, bool shuffle, std::mt19937_64& rng) {
    Cluster c;
    if (!get_cluster(url, &c)) return false;
    try {
        if (shuffle && size < (int)c.workers.size()) {
            // keep a random subset (in its old order, so survivors' relative ranks are stable) instead of a prefix
            std::vector<int> idx(c.workers.size());
            for (size_t i = 0; i < idx.size(); ++i) idx[i] = (int)i;
            std::shuffle(idx.begin(), idx.end(), rng);
            idx.resize(size);
            std::sort(idx.begin(), idx.end());
            c.workers = c.workers.select(idx);
        } else {
            c = c.resized(size);
        }
        c.validate();
    } catch (const std::exception&) {
        return false; // e.g. more workers than the hosts have slots
    }
    std::string out;
    return http_request("PUT", url, c.to_json(), &out) == 200;
}

int main(int argc, char** argv) {
    kftools::Flags fl{argc, argv};
    const std::string url = fl.str("url", "");
    if (url.empty()) {
        std::fprintf(stderr, "usage: kungfu-config-client -url http://host:port/config [-size n | -random -min a -max b -period 1s -duration 10s "
                             "[-final n] [-seed s] [-shuffle] | -clear]\n");
        return 2;
    }
    std::mt19937_64 rng((uint64_t)fl.num("seed", 0));
    const bool shuffle = fl.flag("shuffle");
    if (fl.flag("clear")) {
        std::string out;
        const int st = http_request("DELETE", url, "", &out);
        std::printf("{\"cleared\": %s}\n", st == 200 ? "true" : "false");
        return st == 200 ? 0 : 1;
    }
    const int size = (int)fl.num("size", 0);
    if (size > 0) {
        const bool ok = put_size(url, size, shuffle, rng);
        std::printf("{\"resize_to\": %d, \"ok\": %s}\n", size, ok ? "true" : "false");
        return ok ? 0 : 1;
    }
    if (fl.flag("random")) {
        const int lo = (int)fl.num("min", 1), hi = std::max(lo, (int)fl.num("max", 4)), final_size = (int)fl.num("final", 0);
        const double period = parse_duration_s(fl.str("period", "1s"), 1.0), duration = parse_duration_s(fl.str("duration", "10s"), 10.0);
        const double t_end = now_s() + duration;
        int proposals = 0;
        while (now_s() < t_end) {
            const int n = std::uniform_int_distribution<int>(lo, hi)(rng);
            if (put_size(url, n, shuffle, rng)) {
                ++proposals;
                std::printf("{\"resize_to\": %d}\n", n);
                std::fflush(stdout);
            }
            std::this_thread::sleep_for(std::chrono::duration<double>(period));
        }
        if (final_size > 0) put_size(url, final_size, false, rng);
        std::printf("{\"proposals\": %d, \"final\": %d}\n", proposals, final_size);
        return 0;
    }
    Cluster c;
    if (!get_cluster(url, &c)) {
        std::printf("{\"size\": -1}\n");
        return 1;
    }
    std::printf("{\"size\": %zu, \"runners\": %zu}\n", c.workers.size(), c.runners.size());
    return 0;
}
