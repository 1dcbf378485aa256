// kungfu_b200 runtime — identity and placement model.
//
// Reference: srcs/go/plan/{id.go:9-54, addr.go:11-97, peerlist.go:11-191, hostspec.go:15-218, cluster.go:10-118,
// hostfile/hostfile.go:14-81, interval.go:4-33}. A peer is (IPv4, port); its rank is its index in the PeerList;
// local rank / local size group peers by IPv4. Workers get ports 10000+slot host by host, runners port 38080.
#pragma once
#include <cstdint>
#include <string>
#include <utility>
#include <vector>

namespace kfrt {

constexpr uint16_t kDefaultRunnerPort = 38080;
constexpr uint16_t kDefaultPortBegin = 10000;
constexpr uint16_t kDefaultPortEnd = 11000;

uint32_t parse_ipv4(const std::string& s);       // "a.b.c.d" -> host-order u32; throws on error
std::string format_ipv4(uint32_t ip);
uint32_t resolve_ipv4(const std::string& host);  // dotted quad or DNS name

struct PeerID {
    uint32_t ipv4 = 0;
    uint16_t port = 0;
    std::string str() const;                     // "a.b.c.d:port"
    static PeerID parse(const std::string& s);
    std::string sock_file() const;               // unix socket path for colocated peers
    bool operator==(const PeerID& o) const { return ipv4 == o.ipv4 && port == o.port; }
    bool operator!=(const PeerID& o) const { return !(*this == o); }
    bool operator<(const PeerID& o) const { return ipv4 != o.ipv4 ? ipv4 < o.ipv4 : port < o.port; }
    uint64_t key() const { return ((uint64_t)ipv4 << 16) | port; }
    bool colocated_with(const PeerID& o) const { return ipv4 == o.ipv4; }
};

class PeerList : public std::vector<PeerID> {
public:
    using std::vector<PeerID>::vector;
    int rank_of(const PeerID& p) const;          // -1 if absent
    bool contains(const PeerID& p) const { return rank_of(p) >= 0; }
    int local_rank_of(const PeerID& p) const;    // index among peers with the same IPv4
    int local_size_of(const PeerID& p) const;
    int host_count() const;
    // masters = first rank of every host (in order of appearance); parents[i] = master rank of i's host
    void partition_by_host(std::vector<int>& masters, std::vector<int>& parents) const;
    PeerList select(const std::vector<int>& ranks) const;
    PeerList on_host(uint32_t ipv4) const;
    PeerList others(const PeerID& self) const;
    PeerList minus(const PeerList& o) const;     // elements not in o
    PeerList intersect(const PeerList& o) const;
    bool disjoint(const PeerList& o) const;
    std::string str() const;                     // "ip:port,ip:port"
    static PeerList parse(const std::string& s);
    std::string digest() const;                  // canonical bytes for consensus
};

struct PortRange {
    uint16_t begin = kDefaultPortBegin, end = kDefaultPortEnd;
    int cap() const { return (int)end - (int)begin + 1; }
    static PortRange parse(const std::string& s); // "a-b"
    std::string str() const;
};

struct HostSpec {
    uint32_t ipv4 = 0;
    int slots = 1;
    std::string public_addr;
    static HostSpec parse(const std::string& s); // "ip:slots[:public_addr]" or "ip"
    std::string str() const;
};

class HostList : public std::vector<HostSpec> {
public:
    using std::vector<HostSpec>::vector;
    int cap() const;
    static HostList parse(const std::string& s); // "ip:slots,ip:slots"
    static HostList parse_hostfile(const std::string& text); // OpenMPI syntax: "<host> slots=<n>"
    PeerList gen_peer_list(int np, PortRange pr) const;       // throws if np > capacity
    PeerList gen_runner_list(uint16_t port) const;
    std::string str() const;
    int slots_of(uint32_t ipv4) const;
};

struct Cluster {
    PeerList runners;
    PeerList workers;
    void validate() const;                       // unique peers, one runner per host, every worker host has a runner
    Cluster resized(int new_size) const;         // shrink = keep prefix; grow = add to the least-loaded runner host
    std::string to_json() const;                 // {"Runners":[{"IPv4":u32,"Port":u16}],"Workers":[...]}
    static Cluster from_json(const std::string& s);
    std::string digest() const;
    bool operator==(const Cluster& o) const { return runners == o.runners && workers == o.workers; }
    bool operator!=(const Cluster& o) const { return !(*this == o); }
};

// a versioned membership proposal sent to runners (reference: runner.Stage, srcs/go/kungfu/runner/handler.go:19-30)
struct Stage {
    int version = 0;
    Cluster cluster;
    std::string to_json() const;
    static Stage from_json(const std::string& s);
};

// [begin, end) of part i when n items are split into k nearly equal parts (reference: plan/interval.go:12-27)
std::pair<size_t, size_t> even_partition(size_t n, size_t k, size_t i);

// ---- minimal JSON value reader (objects / arrays / numbers / strings) used for Cluster and HTTP payloads ----
struct JsonValue {
    enum Kind { Null, Bool, Number, String, Array, Object } kind = Null;
    bool b = false;
    double num = 0;
    std::string str;
    std::vector<JsonValue> arr;
    std::vector<std::pair<std::string, JsonValue>> obj;
    const JsonValue* find(const std::string& key) const;
    static JsonValue parse(const std::string& text); // throws kfrt::Error
};
std::string json_escape(const std::string& s);

} // namespace kfrt
