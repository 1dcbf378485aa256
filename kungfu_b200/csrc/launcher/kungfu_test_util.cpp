// kungfu-test-util — fault injection and manual proposals from OUTSIDE a job (no Peer, just a transport client).
//
//   kungfu-test-util -kill 127.0.0.1:10001            the worker (or runner) at that address receives the control message `exit`
//   kungfu-test-util -ping 127.0.0.1:10001 [-n 5]     round-trip time of the transport's ping connection
//   kungfu-test-util -wait 127.0.0.1:10001 [-timeout 10s]   returns once the peer's listener answers
//   kungfu-test-util -propose cluster.json -version 3 the cluster document is pushed to ITS runners as Stage v3 (`update`)
//   kungfu-test-util -notify-start cluster.json       ≙ kungfu-notify-start: Stage v0 to the runners of an `-init-version -1` job
//
// Reference: tests/go/cmd/kungfu-test-util/kungfu-test-util.go:28-35 (kill / ping over the control connection type),
// tests/go/cmd/kungfu-notify-start (first Stage for runners started with `-init-version -1`). A worker only honours `exit` / `update`
// from hosts of its own job (transport.cpp), so the tool is run on one of the job's machines — loopback in the tests.
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <sstream>
#include <string>
#include <chrono>
#include <thread>

#include "base.hpp"
#include "fake_model.hpp"
#include "plan.hpp"
#include "transport.hpp"

using namespace kfrt;

static std::string slurp(const std::string& path) {
    std::ifstream f(path);
    if (!f) throw std::runtime_error("cannot read " + path);
    std::stringstream ss;
    ss << f.rdbuf();
    return ss.str();
}

int main(int argc, char** argv) {
    kftools::Flags fl{argc, argv};
    const std::string kill = fl.str("kill", ""), ping = fl.str("ping", ""), wait = fl.str("wait", ""), propose = fl.str("propose", ""),
                      notify = fl.str("notify-start", "");
    if (kill.empty() && ping.empty() && wait.empty() && propose.empty() && notify.empty()) {
        std::fprintf(stderr, "usage: kungfu-test-util [-kill ip:port] [-ping ip:port [-n k]] [-wait ip:port [-timeout 10s]] "
                             "[-propose cluster.json -version v] [-notify-start cluster.json]\n");
        return 2;
    }
    try {
        Client c(PeerID{parse_ipv4("127.0.0.1"), 0});
        if (!wait.empty()) {
            const double budget = parse_duration_s(fl.str("timeout", "10s"), 10), t0 = now_s();
            const PeerID target = PeerID::parse(wait);
            bool up = false;
            while (!up && now_s() - t0 < budget) {
                try {
                    up = c.ping(target, 500) >= 0;
                } catch (const std::exception&) {
                }
                if (!up) std::this_thread::sleep_for(std::chrono::milliseconds(50));
            }
            std::printf("{\"peer\": \"%s\", \"up\": %s, \"waited_s\": %.3f}\n", wait.c_str(), up ? "true" : "false", now_s() - t0);
            if (!up) return 1;
        }
        if (!ping.empty()) {
            const int n = (int)fl.num("n", 1);
            double best = 1e30, sum = 0;
            for (int i = 0; i < n; ++i) {
                const double rtt = c.ping(PeerID::parse(ping));
                if (rtt < 0) throw std::runtime_error("no answer from " + ping);
                best = std::min(best, rtt), sum += rtt;
            }
            std::printf("{\"peer\": \"%s\", \"pings\": %d, \"rtt_s\": %.6f, \"mean_rtt_s\": %.6f}\n", ping.c_str(), n, best, sum / std::max(n, 1));
        }
        auto push = [&](const std::string& path, int version) {
            Stage st;
            st.version = version;
            st.cluster = Cluster::from_json(slurp(path));
            st.cluster.validate();
            const std::string doc = st.to_json();
            int sent = 0;
            for (const PeerID& r : st.cluster.runners) {
                c.send(r, "update", doc.data(), doc.size(), kConnControl, kNoFlag);
                ++sent;
            }
            std::printf("{\"stage\": %d, \"workers\": %zu, \"runners_notified\": %d}\n", version, st.cluster.workers.size(), sent);
        };
        if (!propose.empty()) push(propose, (int)fl.num("version", 0));
        if (!notify.empty()) push(notify, 0);
        if (!kill.empty()) {
            c.send(PeerID::parse(kill), "exit", "", 0, kConnControl, kNoFlag);
            std::printf("{\"killed\": \"%s\"}\n", kill.c_str());
        }
        c.close_all();
    } catch (const std::exception& e) {
        std::fprintf(stderr, "kungfu-test-util: %s\n", e.what());
        return 1;
    }
    return 0;
}
