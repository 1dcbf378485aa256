// Implementation of numa.hpp (the header carries the design notes and the reference file:line mapping).
#include "numa.hpp"

#include <sched.h>

#include <algorithm>
#include <array>
#include <fstream>
#include <sstream>
#include <string>

#include "base.hpp"

namespace kfrt {

static std::vector<int> parse_cpulist(const std::string& s) {
    std::vector<int> out;
    std::stringstream ss(s);
    std::string part;
    while (std::getline(ss, part, ',')) {
        if (part.empty()) continue;
        auto dash = part.find('-');
        int a = atoi(part.c_str()), b = a;
        if (dash != std::string::npos) b = atoi(part.c_str() + dash + 1);
        for (int i = a; i <= b; ++i) out.push_back(i);
    }
    return out;
}

std::vector<std::vector<int>> numa_topology() {
    std::vector<std::vector<int>> nodes;
    for (int n = 0; n < 64; ++n) {
        std::ifstream f("/sys/devices/system/node/node" + std::to_string(n) + "/cpulist");
        if (!f) break;
        std::string line;
        std::getline(f, line);
        auto cpus = parse_cpulist(line);
        if (!cpus.empty()) nodes.push_back(cpus);
    }
    if (nodes.empty()) {
        cpu_set_t set;
        CPU_ZERO(&set);
        std::vector<int> all;
        if (sched_getaffinity(0, sizeof(set), &set) == 0)
            for (int i = 0; i < CPU_SETSIZE; ++i)
                if (CPU_ISSET(i, &set)) all.push_back(i);
        nodes.push_back(all);
    }
    return nodes;
}

// Topology (hwloc-like) order of the PUs of every NUMA node: package -> core -> hardware thread, i.e. the depth-first order of
// the machine tree that the reference takes from hwloc (srcs/cpp/src/numa/hwloc.cpp:8-63 `get_pu_numa_order`), read from sysfs
// instead of a library. With Linux' OS numbering the two hyper-threads of a core are N/2 apart, so a contiguous OS-order
// slice hands a worker one thread of many cores that it then shares with another worker; in topology order a slice is a
// set of whole cores.
std::vector<std::vector<int>> numa_topology_ordered() {
    auto nodes = numa_topology();
    auto read_int = [](const std::string& path, int dflt) {
        std::ifstream f(path);
        int v = dflt;
        if (f) f >> v;
        return v;
    };
    for (auto& cpus : nodes) {
        std::vector<std::array<int, 4>> keyed; // package, core, position among the core's threads (os order), cpu
        for (int c : cpus) {
            const std::string base = "/sys/devices/system/cpu/cpu" + std::to_string(c) + "/topology/";
            keyed.push_back({read_int(base + "physical_package_id", 0), read_int(base + "core_id", c), c, c});
        }
        std::sort(keyed.begin(), keyed.end());
        for (size_t i = 0; i < cpus.size(); ++i) cpus[i] = keyed[i][3];
    }
    return nodes;
}

std::vector<int> select_cpus(const std::vector<std::vector<int>>& numa_cpus, int local_rank, int local_size) {
    if (numa_cpus.empty() || local_size <= 0 || local_rank < 0 || local_rank >= local_size) return {};
    const int nn = (int)numa_cpus.size();
    // workers are spread over NUMA nodes in contiguous groups; inside a node each gets a contiguous slice
    const int per_node = (local_size + nn - 1) / nn;
    const int node = (local_rank / per_node) % nn;
    const int idx = local_rank % per_node;
    int on_node = per_node;
    if ((node + 1) * per_node > local_size) on_node = local_size - node * per_node;
    if (on_node <= 0) on_node = 1;
    const std::vector<int>& cpus = numa_cpus[node];
    const int n = (int)cpus.size();
    const int chunk = n / on_node > 0 ? n / on_node : 1;
    const int b = (idx * chunk) % n;
    std::vector<int> out;
    for (int i = 0; i < chunk && b + i < n; ++i) out.push_back(cpus[b + i]);
    return out;
}

bool set_affinity(const std::vector<int>& cpus) {
    if (cpus.empty()) return false;
    cpu_set_t set;
    CPU_ZERO(&set);
    for (int c : cpus) CPU_SET(c, &set);
    return sched_setaffinity(0, sizeof(set), &set) == 0;
}

std::vector<int> apply_affinity_from_env(int local_rank, int local_size) {
    if (!getenv_bool("KUNGFU_USE_AFFINITY", false)) return {};
    // KUNGFU_AFFINITY_ORDER=topology (default: whole cores per worker, the reference's hwloc build) | os (plain OS numbering)
    const bool topo = getenv_str("KUNGFU_AFFINITY_ORDER", "topology") != "os";
    auto cpus = select_cpus(topo ? numa_topology_ordered() : numa_topology(), local_rank, local_size);
    if (!set_affinity(cpus)) KF_LOGW("failed to set CPU affinity");
    return cpus;
}

} // namespace kfrt
