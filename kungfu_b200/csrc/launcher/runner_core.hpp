// Launcher core shared by `kungfu-run` and the C++ unit tests: command-line flags, worker process description (environment contract),
// GPU slot pool and the mapping of a slot through the launcher's own CUDA_VISIBLE_DEVICES.
//
// Reference: srcs/go/kungfu/runner/flags.go:69-104 (flag set), srcs/go/kungfu/job/{job.go:30-98, gpu_resource.go:11-51,
// cuda_visible_device.go} and their unit tests (runner/flags_test.go, job/cuda_visible_device_test.go, proc/proc_test.go).
#pragma once
#include <arpa/inet.h>
#include <ifaddrs.h>
#include <net/if.h>
#include <netinet/in.h>

#include <algorithm>
#include <cstdlib>
#include <ctime>
#include <map>
#include <mutex>
#include <sstream>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "base.hpp"
#include "graph.hpp"
#include "plan.hpp"
#include "transport.hpp"

namespace kfrun {
using namespace kfrt;

// ---------------------------------------------------------------------------------------------------
// flags
// ---------------------------------------------------------------------------------------------------
struct Flags {
    int np = 1;
    double auto_recover_s = 0;
    std::string hosts = "127.0.0.1:" + std::to_string(std::max(1u, std::thread::hardware_concurrency()));
    std::string hostfile;
    std::string peer_list; // -P: parsed for parity, unused (as in the reference)
    std::string user;
    PortRange port_range;
    std::string self;
    double timeout_s = 0;
    bool verbose = true;
    std::string nic;
    bool allow_nvlink = true;
    std::string strategy = "BINARY_TREE_STAR";
    int port = kDefaultRunnerPort;
    int debug_port = 0;
    int standby = 0; // -standby N (watch mode): warmed-up processes kept ready to become workers at the next grow
    bool watch = false;
    bool keep = false;
    bool shrink_on_failure = false; // -shrink-on-failure (watch mode): a crashed worker is removed from the cluster, the rest go on
    bool refill = false;            // -refill: after a shrink, propose a cluster with a replacement worker to the config server
    int init_version = 0;
    std::string config_server;
    long t0 = (long)time(nullptr);
    std::string logfile, logdir;
    bool quiet = false;
    double delay_s = 0;
    int builtin_config_port = 0;
    std::string prog;
    std::vector<std::string> args;
};

// A command line the launcher cannot act on (or `-h`): `what` is empty for a plain help request. Thrown instead of exiting so that the
// parser can be unit-tested; main() prints `usage_text()` and exits 2 (0 for help).
struct UsageError : std::runtime_error {
    using std::runtime_error::runtime_error;
};

inline std::string usage_text() {
    std::string t =
        "usage: kungfu-run [-np N] [-H ip:slots,...] [-hostfile f] [-self ip] [-nic name] [-strategy S]\n"
        "                  [-port-range a-b] [-port p] [-w] [-k] [-config-server URL] [-builtin-config-port p]\n"
        "                  [-init-version v] [-timeout d] [-logdir d] [-logfile f] [-q] [-v] [-allow-nvlink]\n"
        "                  [-auto-recover d] [-delay d] [-debug-port p] [-standby n] [-shrink-on-failure] [-refill]\n"
        "                  prog args...\n"
        "  -w -shrink-on-failure: a worker that exits with an error is removed, the others continue; with -auto-recover d also a\n"
        "  worker whose batch stays open for d while it is not waiting in a collective (it hangs) is killed and removed.\n"
        "  -refill: after such a loss the runner proposes a cluster with a replacement worker (a warm standby if there is one) to\n"
        "  the config server; workers that follow it (elastic.ShrinkOnFailure(follow_config_server=True)) grow back.\n"
        "strategies: ";
    for (auto& s : strategy_names()) t += s + " ";
    return t + "\n";
}

[[noreturn]] inline void usage(const std::string& err) { throw UsageError(err); }

inline bool parse_bool(const std::string& v) { return v == "1" || v == "true" || v == "True" || v == "yes" || v.empty(); }

inline Flags parse_flags(int argc, char** argv) {
    Flags f;
    int i = 1;
    auto is_bool_flag = [](const std::string& k) {
        return k == "w" || k == "k" || k == "q" || k == "v" || k == "allow-nvlink" || k == "h" || k == "help" || k == "shrink-on-failure" || k == "refill";
    };
    for (; i < argc; ++i) {
        std::string a = argv[i];
        if (a == "--") {
            ++i;
            break;
        }
        if (a.size() < 2 || a[0] != '-') break;
        std::string k = a.substr(a[1] == '-' ? 2 : 1), v;
        bool has_v = false;
        auto eq = k.find('=');
        if (eq != std::string::npos) {
            v = k.substr(eq + 1);
            k = k.substr(0, eq);
            has_v = true;
        }
        if (!has_v && !is_bool_flag(k)) {
            if (i + 1 >= argc) usage("flag -" + k + " needs a value");
            v = argv[++i];
        }
        if (k == "np") f.np = atoi(v.c_str());
        else if (k == "auto-recover") f.auto_recover_s = parse_duration_s(v, 0);
        else if (k == "H") f.hosts = v;
        else if (k == "hostfile") f.hostfile = v;
        else if (k == "P") f.peer_list = v;
        else if (k == "u") f.user = v;
        else if (k == "port-range") f.port_range = PortRange::parse(v);
        else if (k == "self") f.self = v;
        else if (k == "timeout") f.timeout_s = parse_duration_s(v, 0);
        else if (k == "v") f.verbose = parse_bool(v);
        else if (k == "nic") f.nic = v;
        else if (k == "allow-nvlink") f.allow_nvlink = parse_bool(v);
        else if (k == "strategy") f.strategy = v;
        else if (k == "port") f.port = atoi(v.c_str());
        else if (k == "debug-port") f.debug_port = atoi(v.c_str());
        else if (k == "standby") f.standby = atoi(v.c_str());
        else if (k == "w") f.watch = parse_bool(v);
        else if (k == "k") f.keep = parse_bool(v);
        else if (k == "shrink-on-failure") f.shrink_on_failure = parse_bool(v);
        else if (k == "refill") f.refill = parse_bool(v);
        else if (k == "init-version") f.init_version = atoi(v.c_str());
        else if (k == "config-server") f.config_server = v;
        else if (k == "t0") f.t0 = atol(v.c_str());
        else if (k == "logfile") f.logfile = v;
        else if (k == "logdir") f.logdir = v;
        else if (k == "q") f.quiet = parse_bool(v);
        else if (k == "delay") f.delay_s = parse_duration_s(v, 0);
        else if (k == "builtin-config-port") f.builtin_config_port = atoi(v.c_str());
        else if (k == "h" || k == "help") usage("");
        else usage("unknown flag -" + k);
    }
    if (i >= argc) usage("missing program name");
    f.prog = argv[i++];
    for (; i < argc; ++i) f.args.push_back(argv[i]);
    parse_strategy(f.strategy); // validate early
    return f;
}

inline uint32_t infer_self_ipv4(const std::string& self, const std::string& nic) {
    if (!self.empty()) return resolve_ipv4(self);
    if (!nic.empty()) {
        ifaddrs* ifa = nullptr;
        if (getifaddrs(&ifa) == 0) {
            for (ifaddrs* p = ifa; p; p = p->ifa_next)
                if (p->ifa_addr && p->ifa_addr->sa_family == AF_INET && nic == p->ifa_name) {
                    uint32_t ip = ntohl(reinterpret_cast<sockaddr_in*>(p->ifa_addr)->sin_addr.s_addr);
                    freeifaddrs(ifa);
                    return ip;
                }
            freeifaddrs(ifa);
        }
        fail("network interface '" + nic + "' has no IPv4 address");
    }
    return parse_ipv4("127.0.0.1");
}

// ---------------------------------------------------------------------------------------------------
// job / proc
// ---------------------------------------------------------------------------------------------------
struct Proc {
    std::string name; // ip.port
    std::string prog;
    std::vector<std::string> args;
    std::map<std::string, std::string> envs;
    std::string log_prefix;
};

// smallest-free GPU id pool (reference: job/gpu_resource.go:11-51)
class GpuPool {
public:
    explicit GpuPool(int n) : used_(n > 0 ? n : 0, false) {}
    int get() {
        std::lock_guard<std::mutex> g(mu_);
        for (size_t i = 0; i < used_.size(); ++i)
            if (!used_[i]) {
                used_[i] = true;
                return (int)i;
            }
        return -1;
    }
    void put(int id) {
        std::lock_guard<std::mutex> g(mu_);
        if (id >= 0 && id < (int)used_.size()) used_[id] = false;
    }

private:
    std::mutex mu_;
    std::vector<bool> used_;
};

struct Job {
    std::string hang_watch; // URL the workers report "waiting in a collective" to (watch mode with -shrink-on-failure -auto-recover)
    long start_time = 0;
    std::string strategy;
    PeerID parent;
    HostList hosts;
    PortRange port_range;
    std::string prog;
    std::vector<std::string> args;
    std::string logdir;
    std::string config_server;
    bool allow_nvlink = true;

    // maps a local GPU slot through the launcher's own CUDA_VISIBLE_DEVICES (reference: job/cuda_visible_device.go)
    static std::string cuda_index(int slot) {
        const char* v = getenv("CUDA_VISIBLE_DEVICES");
        if (!v) return std::to_string(slot);
        std::vector<std::string> ids;
        std::stringstream ss(v);
        std::string part;
        while (std::getline(ss, part, ',')) ids.push_back(part);
        if (slot >= 0 && slot < (int)ids.size()) return ids[slot];
        return "-1";
    }

    Proc new_proc(const PeerID& peer, int gpu_slot, int version, const Cluster& cluster) const {
        Proc p;
        p.name = format_ipv4(peer.ipv4) + "." + std::to_string(peer.port);
        p.prog = prog;
        p.args = args;
        p.log_prefix = p.name + "@" + std::to_string(version);
        auto& e = p.envs;
        e[envkeys::kSelfSpec] = peer.str();
        e[envkeys::kInitPeers] = cluster.workers.str();
        e[envkeys::kInitRunners] = cluster.runners.str();
        e[envkeys::kParentID] = parent.str();
        e[envkeys::kInitClusterVersion] = std::to_string(version);
        e[envkeys::kAllReduceStrategy] = strategy;
        e[envkeys::kJobStart] = std::to_string(start_time);
        e[envkeys::kProcStart] = std::to_string((long)time(nullptr));
        if (!config_server.empty()) e[envkeys::kConfigServer] = config_server;
        if (!hang_watch.empty()) e["KUNGFU_HANG_WATCH"] = hang_watch;
        const std::string idx = cuda_index(gpu_slot);
        e[envkeys::kCudaVisibleDevices] = idx;
        if (allow_nvlink) e[envkeys::kAllowNvlink] = "1";
        else e["CUDA_VISIBLE_DEVICES"] = idx; // reference default: hide the other GPUs (kills P2P)
        e["PYTHONUNBUFFERED"] = "1";
        if (!getenv("OMP_NUM_THREADS")) {
            // One OpenMP pool of ncores threads per worker oversubscribes the host as soon as several workers share it (measured:
            // a 1 MiB torch.zeros took 25 ms with 4 workers x 8 threads on 8 cores). Split the cores between the workers this host
            // can hold, like torchrun's default of 1, unless the user chose a value.
            const int cores = (int)std::thread::hardware_concurrency();
            const int slots = std::max(1, hosts.slots_of(peer.ipv4));
            e["OMP_NUM_THREADS"] = std::to_string(std::max(1, cores / slots));
        }
        return p;
    }

    std::vector<Proc> create_procs(const Cluster& cluster, uint32_t host, int version) const {
        std::vector<Proc> ps;
        int slot = 0;
        for (auto& w : cluster.workers)
            if (w.ipv4 == host) ps.push_back(new_proc(w, slot++, version, cluster));
        return ps;
    }
};

// Start-up retry (reference: srcs/go/utils/runner/local/hack.go:14-35, srcs/go/nccl/bug.go:12): a worker that exits non-zero with the
// dynamic loader's "Inconsistency detected by ld.so" as the FIRST line on its stderr hit a known loader race of some CUDA / NCCL builds
// before it joined anything, so it is simply started again with the same environment. The reference retries for ever; here `attempt`
// (1 = first restart) is bounded by `limit`.
inline const char* loader_bug_prefix() { return "Inconsistency detected by ld.so"; }
inline bool restart_after_startup_failure(const std::string& first_stderr_line, int exit_code, int attempt, const std::string& prefix, int limit) {
    return exit_code != 0 && !prefix.empty() && attempt >= 1 && attempt <= limit && first_stderr_line.compare(0, prefix.size(), prefix) == 0;
}

} // namespace kfrun
