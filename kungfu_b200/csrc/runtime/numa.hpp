// kungfu_b200 runtime — NUMA-aware CPU affinity for co-located workers.
//
// Reference: srcs/cpp/src/numa/{placement.cpp:5-16, affinity.cpp:26-62, hwloc.cpp:8-63}, gated by KUNGFU_USE_AFFINITY
// (srcs/cpp/src/python/init.cpp:23-28): each local rank gets a contiguous slice of the CPUs of "its" NUMA node.
// hwloc is replaced by /sys/devices/system/node (always present on Linux).
#pragma once
#include <vector>

namespace kfrt {

// CPUs grouped by NUMA node, read from sysfs (one group with every online CPU when sysfs has no NUMA info)
std::vector<std::vector<int>> numa_topology();
std::vector<std::vector<int>> numa_topology_ordered(); // per node: package -> core -> thread (hwloc-like PU order from sysfs)
// pure placement function (unit-testable): CPUs for `local_rank` out of `local_size` workers
std::vector<int> select_cpus(const std::vector<std::vector<int>>& numa_cpus, int local_rank, int local_size);
bool set_affinity(const std::vector<int>& cpus);
// applies select_cpus + set_affinity when KUNGFU_USE_AFFINITY is truthy; returns the chosen CPUs (empty if disabled)
std::vector<int> apply_affinity_from_env(int local_rank, int local_size);

} // namespace kfrt
