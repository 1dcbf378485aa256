// Implementation of plan.hpp (the header carries the design notes and the reference file:line mapping).
#include "plan.hpp"

#include <arpa/inet.h>
#include <netdb.h>
#include <sys/socket.h>

#include <algorithm>
#include <cstring>
#include <map>
#include <set>
#include <sstream>

#include "base.hpp"

namespace kfrt {

// ---- helpers --------------------------------------------------------------------------------------
static std::vector<std::string> split(const std::string& s, char sep, bool keep_empty = false) {
    std::vector<std::string> out;
    std::string cur;
    for (char c : s) {
        if (c == sep) {
            if (keep_empty || !cur.empty()) out.push_back(cur);
            cur.clear();
        } else {
            cur.push_back(c);
        }
    }
    if (keep_empty || !cur.empty()) out.push_back(cur);
    return out;
}
static std::string trim(const std::string& s) {
    size_t a = 0, b = s.size();
    while (a < b && isspace((unsigned char)s[a])) ++a;
    while (b > a && isspace((unsigned char)s[b - 1])) --b;
    return s.substr(a, b - a);
}
static int parse_int(const std::string& s, const char* what) {
    char* end = nullptr;
    long v = strtol(s.c_str(), &end, 10);
    if (end == s.c_str() || *end != '\0') fail(std::string("invalid ") + what + ": '" + s + "'");
    return (int)v;
}

uint32_t parse_ipv4(const std::string& s) {
    in_addr a;
    if (inet_pton(AF_INET, s.c_str(), &a) != 1) fail("invalid IPv4 address: '" + s + "'");
    return ntohl(a.s_addr);
}
std::string format_ipv4(uint32_t ip) {
    return strprintf("%u.%u.%u.%u", (ip >> 24) & 255, (ip >> 16) & 255, (ip >> 8) & 255, ip & 255);
}
uint32_t resolve_ipv4(const std::string& host) {
    in_addr a;
    if (inet_pton(AF_INET, host.c_str(), &a) == 1) return ntohl(a.s_addr);
    addrinfo hints{}, *res = nullptr;
    hints.ai_family = AF_INET;
    if (getaddrinfo(host.c_str(), nullptr, &hints, &res) != 0 || !res) fail("cannot resolve host '" + host + "'");
    uint32_t ip = ntohl(reinterpret_cast<sockaddr_in*>(res->ai_addr)->sin_addr.s_addr);
    freeaddrinfo(res);
    return ip;
}

// ---- PeerID ---------------------------------------------------------------------------------------
std::string PeerID::str() const { return format_ipv4(ipv4) + ":" + std::to_string(port); }
PeerID PeerID::parse(const std::string& s) {
    auto pos = s.rfind(':');
    if (pos == std::string::npos) fail("invalid peer id '" + s + "' (want ip:port)");
    PeerID p;
    p.ipv4 = resolve_ipv4(s.substr(0, pos));
    int port = parse_int(s.substr(pos + 1), "port");
    if (port < 0 || port > 65535) fail("port out of range in '" + s + "'");
    p.port = (uint16_t)port;
    return p;
}
std::string PeerID::sock_file() const {
    // the address is part of the name so that loopback-alias "hosts" (127.0.0.2:10000 next to 127.0.0.1:10000) do not collide
    if (ipv4 == 0x7f000001u) return "/tmp/kungfu-b200-" + std::to_string(port) + ".sock";
    return "/tmp/kungfu-b200-" + format_ipv4(ipv4) + "-" + std::to_string(port) + ".sock";
}

// ---- PeerList -------------------------------------------------------------------------------------
int PeerList::rank_of(const PeerID& p) const {
    for (size_t i = 0; i < size(); ++i)
        if ((*this)[i] == p) return (int)i;
    return -1;
}
int PeerList::local_rank_of(const PeerID& p) const {
    int r = 0;
    for (auto& q : *this) {
        if (q == p) return r;
        if (q.ipv4 == p.ipv4) ++r;
    }
    return -1;
}
int PeerList::local_size_of(const PeerID& p) const {
    int n = 0;
    for (auto& q : *this)
        if (q.ipv4 == p.ipv4) ++n;
    return n;
}
int PeerList::host_count() const {
    std::set<uint32_t> s;
    for (auto& q : *this) s.insert(q.ipv4);
    return (int)s.size();
}
void PeerList::partition_by_host(std::vector<int>& masters, std::vector<int>& parents) const {
    masters.clear();
    parents.assign(size(), 0);
    std::map<uint32_t, int> first;
    for (size_t i = 0; i < size(); ++i) {
        auto it = first.find((*this)[i].ipv4);
        if (it == first.end()) {
            first[(*this)[i].ipv4] = (int)i;
            masters.push_back((int)i);
            parents[i] = (int)i;
        } else {
            parents[i] = it->second;
        }
    }
}
PeerList PeerList::select(const std::vector<int>& ranks) const {
    PeerList out;
    for (int r : ranks) out.push_back(at(r));
    return out;
}
PeerList PeerList::on_host(uint32_t ip) const {
    PeerList out;
    for (auto& q : *this)
        if (q.ipv4 == ip) out.push_back(q);
    return out;
}
PeerList PeerList::others(const PeerID& self) const {
    PeerList out;
    for (auto& q : *this)
        if (q != self) out.push_back(q);
    return out;
}
PeerList PeerList::minus(const PeerList& o) const {
    PeerList out;
    for (auto& q : *this)
        if (!o.contains(q)) out.push_back(q);
    return out;
}
PeerList PeerList::intersect(const PeerList& o) const {
    PeerList out;
    for (auto& q : *this)
        if (o.contains(q)) out.push_back(q);
    return out;
}
bool PeerList::disjoint(const PeerList& o) const { return intersect(o).empty(); }
std::string PeerList::str() const {
    std::string s;
    for (size_t i = 0; i < size(); ++i) {
        if (i) s += ",";
        s += (*this)[i].str();
    }
    return s;
}
PeerList PeerList::parse(const std::string& s) {
    PeerList out;
    for (auto& part : split(s, ',')) out.push_back(PeerID::parse(trim(part)));
    return out;
}
std::string PeerList::digest() const {
    std::string d;
    for (auto& q : *this) {
        uint64_t k = q.key();
        d.append(reinterpret_cast<const char*>(&k), sizeof(k));
    }
    return d;
}

// ---- PortRange / HostSpec / HostList --------------------------------------------------------------
PortRange PortRange::parse(const std::string& s) {
    auto parts = split(s, '-');
    if (parts.size() != 2) fail("invalid port range '" + s + "' (want begin-end)");
    PortRange pr;
    int b = parse_int(parts[0], "port"), e = parse_int(parts[1], "port");
    if (b <= 0 || e < b || e > 65535) fail("invalid port range '" + s + "'");
    pr.begin = (uint16_t)b;
    pr.end = (uint16_t)e;
    return pr;
}
std::string PortRange::str() const { return std::to_string(begin) + "-" + std::to_string(end); }

HostSpec HostSpec::parse(const std::string& s) {
    auto parts = split(s, ':', true);
    HostSpec h;
    if (parts.empty() || parts[0].empty()) fail("invalid host spec '" + s + "'");
    h.ipv4 = resolve_ipv4(parts[0]);
    h.slots = parts.size() > 1 && !parts[1].empty() ? parse_int(parts[1], "slots") : 1;
    if (h.slots < 0) fail("negative slots in host spec '" + s + "'");
    h.public_addr = parts.size() > 2 ? parts[2] : parts[0];
    if (parts.size() > 3) fail("invalid host spec '" + s + "'");
    return h;
}
std::string HostSpec::str() const { return format_ipv4(ipv4) + ":" + std::to_string(slots) + ":" + public_addr; }

int HostList::cap() const {
    int c = 0;
    for (auto& h : *this) c += h.slots;
    return c;
}
HostList HostList::parse(const std::string& s) {
    HostList hl;
    std::set<uint32_t> seen;
    for (auto& part : split(s, ',')) {
        HostSpec h = HostSpec::parse(trim(part));
        if (!seen.insert(h.ipv4).second) fail("duplicated host " + format_ipv4(h.ipv4));
        hl.push_back(h);
    }
    return hl;
}
HostList HostList::parse_hostfile(const std::string& text) {
    HostList hl;
    std::istringstream in(text);
    std::string line;
    while (std::getline(in, line)) {
        auto hash = line.find('#');
        if (hash != std::string::npos) line = line.substr(0, hash);
        line = trim(line);
        if (line.empty()) continue;
        std::istringstream ls(line);
        std::string host, tok;
        ls >> host;
        HostSpec h;
        h.ipv4 = resolve_ipv4(host);
        h.public_addr = host;
        h.slots = 1;
        while (ls >> tok) {
            auto eq = tok.find('=');
            if (eq == std::string::npos) fail("invalid hostfile token '" + tok + "'");
            std::string k = tok.substr(0, eq), v = tok.substr(eq + 1);
            if (k == "slots") h.slots = parse_int(v, "slots");
            else if (k == "public_addr") h.public_addr = v;
            // other OpenMPI keys (max_slots, ...) are accepted and ignored
        }
        hl.push_back(h);
    }
    return hl;
}
PeerList HostList::gen_peer_list(int np, PortRange pr) const {
    if (cap() < np) fail("no enough capacity: np=" + std::to_string(np) + " > " + std::to_string(cap()) + " slots");
    for (auto& h : *this)
        if (pr.cap() < h.slots) fail("port range " + pr.str() + " smaller than slots of " + format_ipv4(h.ipv4));
    PeerList pl;
    if (np == 0) return pl;
    for (auto& h : *this)
        for (int j = 0; j < h.slots; ++j) {
            pl.push_back(PeerID{h.ipv4, (uint16_t)(pr.begin + j)});
            if ((int)pl.size() >= np) return pl;
        }
    return pl;
}
PeerList HostList::gen_runner_list(uint16_t port) const {
    PeerList pl;
    for (auto& h : *this) pl.push_back(PeerID{h.ipv4, port});
    return pl;
}
std::string HostList::str() const {
    std::string s;
    for (size_t i = 0; i < size(); ++i) {
        if (i) s += ",";
        s += (*this)[i].str();
    }
    return s;
}
int HostList::slots_of(uint32_t ip) const {
    for (auto& h : *this)
        if (h.ipv4 == ip) return h.slots;
    return 0;
}

// ---- Cluster --------------------------------------------------------------------------------------
void Cluster::validate() const {
    std::set<uint64_t> seen;
    std::set<uint32_t> runner_hosts;
    for (auto& r : runners) {
        if (!seen.insert(r.key()).second) fail("duplicated peer " + r.str());
        if (!runner_hosts.insert(r.ipv4).second) fail("more than one runner on host " + format_ipv4(r.ipv4));
    }
    for (auto& w : workers) {
        if (!seen.insert(w.key()).second) fail("duplicated peer " + w.str());
        if (!runner_hosts.count(w.ipv4)) fail("worker " + w.str() + " has no runner on its host");
    }
}
Cluster Cluster::resized(int new_size) const {
    if (new_size < 0) fail("negative cluster size");
    Cluster d = *this;
    if ((int)d.workers.size() > new_size) d.workers.resize(new_size);
    while ((int)d.workers.size() < new_size) {
        if (d.runners.empty()) fail("no runner in cluster");
        std::map<uint32_t, int> used;
        for (auto& r : d.runners) used[r.ipv4] = 0;
        for (auto& w : d.workers) used[w.ipv4]++;
        uint32_t ip = d.runners[0].ipv4;
        for (auto& r : d.runners)
            if (used[r.ipv4] < used[ip]) ip = r.ipv4;
        uint16_t port = 0;
        for (auto& w : d.workers)
            if (w.ipv4 == ip && port <= w.port) port = (uint16_t)(w.port + 1);
        if (port == 0) port = kDefaultPortBegin;
        d.workers.push_back(PeerID{ip, port});
    }
    return d;
}
static void peers_json(std::ostringstream& os, const PeerList& pl) {
    os << "[";
    for (size_t i = 0; i < pl.size(); ++i) {
        if (i) os << ",";
        os << "{\"IPv4\":" << pl[i].ipv4 << ",\"Port\":" << pl[i].port << "}";
    }
    os << "]";
}
std::string Cluster::to_json() const {
    std::ostringstream os;
    os << "{\"Runners\":";
    peers_json(os, runners);
    os << ",\"Workers\":";
    peers_json(os, workers);
    os << "}";
    return os.str();
}
static PeerList peers_from_json(const JsonValue* v) {
    PeerList pl;
    if (!v || v->kind == JsonValue::Null) return pl;
    if (v->kind != JsonValue::Array) fail("cluster json: peer list must be an array");
    for (auto& e : v->arr) {
        const JsonValue* ip = e.find("IPv4");
        const JsonValue* port = e.find("Port");
        if (!ip || !port) fail("cluster json: peer needs IPv4 and Port");
        PeerID p;
        p.ipv4 = ip->kind == JsonValue::String ? parse_ipv4(ip->str) : (uint32_t)ip->num;
        p.port = (uint16_t)port->num;
        pl.push_back(p);
    }
    return pl;
}
Cluster Cluster::from_json(const std::string& s) {
    JsonValue v = JsonValue::parse(s);
    if (v.kind != JsonValue::Object) fail("cluster json: object expected");
    Cluster c;
    c.runners = peers_from_json(v.find("Runners"));
    c.workers = peers_from_json(v.find("Workers"));
    return c;
}
std::string Cluster::digest() const { return runners.digest() + "|" + workers.digest(); }

std::string Stage::to_json() const { return "{\"Version\":" + std::to_string(version) + ",\"Cluster\":" + cluster.to_json() + "}"; }
Stage Stage::from_json(const std::string& s) {
    JsonValue v = JsonValue::parse(s);
    Stage st;
    const JsonValue* ver = v.find("Version");
    const JsonValue* cl = v.find("Cluster");
    if (!ver || !cl) fail("stage json: Version and Cluster required");
    st.version = (int)ver->num;
    st.cluster.runners = peers_from_json(cl->find("Runners"));
    st.cluster.workers = peers_from_json(cl->find("Workers"));
    return st;
}

std::pair<size_t, size_t> even_partition(size_t n, size_t k, size_t i) {
    if (k == 0) return {0, 0};
    const size_t q = n / k, r = n % k;
    const size_t b = i * q + std::min(i, r);
    const size_t e = b + q + (i < r ? 1 : 0);
    return {b, e};
}

// ---- JSON -----------------------------------------------------------------------------------------
namespace {
struct JsonParser {
    const std::string& t;
    size_t i = 0;
    explicit JsonParser(const std::string& text) : t(text) {}
    void ws() {
        while (i < t.size() && isspace((unsigned char)t[i])) ++i;
    }
    [[noreturn]] void err(const char* what) { fail(std::string("json: ") + what + " at offset " + std::to_string(i)); }
    // The documents this reader sees (cluster, stage, heartbeat keys) are three levels deep; the text comes from the network (config
    // server PUT, control messages), and a recursive-descent reader without a limit turns 100 000 '[' into a stack overflow (found by
    // tests/test_parsers_fuzz.py).
    static constexpr int kMaxDepth = 64;
    int depth = 0;
    struct Nest {
        JsonParser& p;
        explicit Nest(JsonParser& q) : p(q) {
            if (++p.depth > kMaxDepth) p.err("nesting too deep");
        }
        ~Nest() { --p.depth; }
    };
    JsonValue value() {
        Nest nest(*this);
        ws();
        if (i >= t.size()) err("unexpected end");
        char c = t[i];
        JsonValue v;
        if (c == '{') {
            v.kind = JsonValue::Object;
            ++i;
            ws();
            if (i < t.size() && t[i] == '}') {
                ++i;
                return v;
            }
            while (true) {
                ws();
                JsonValue k = value();
                if (k.kind != JsonValue::String) err("object key must be a string");
                ws();
                if (i >= t.size() || t[i] != ':') err("':' expected");
                ++i;
                v.obj.emplace_back(k.str, value());
                ws();
                if (i < t.size() && t[i] == ',') {
                    ++i;
                    continue;
                }
                if (i < t.size() && t[i] == '}') {
                    ++i;
                    return v;
                }
                err("',' or '}' expected");
            }
        }
        if (c == '[') {
            v.kind = JsonValue::Array;
            ++i;
            ws();
            if (i < t.size() && t[i] == ']') {
                ++i;
                return v;
            }
            while (true) {
                v.arr.push_back(value());
                ws();
                if (i < t.size() && t[i] == ',') {
                    ++i;
                    continue;
                }
                if (i < t.size() && t[i] == ']') {
                    ++i;
                    return v;
                }
                err("',' or ']' expected");
            }
        }
        if (c == '"') {
            v.kind = JsonValue::String;
            ++i;
            while (i < t.size() && t[i] != '"') {
                if (t[i] == '\\' && i + 1 < t.size()) {
                    char e = t[i + 1];
                    i += 2;
                    switch (e) {
                    case 'n': v.str.push_back('\n'); break;
                    case 't': v.str.push_back('\t'); break;
                    case 'r': v.str.push_back('\r'); break;
                    case 'u': i += 4; v.str.push_back('?'); break;
                    default: v.str.push_back(e);
                    }
                } else {
                    v.str.push_back(t[i++]);
                }
            }
            if (i >= t.size()) err("unterminated string");
            ++i;
            return v;
        }
        if (!t.compare(i, 4, "null")) {
            i += 4;
            return v;
        }
        if (!t.compare(i, 4, "true")) {
            i += 4;
            v.kind = JsonValue::Bool;
            v.b = true;
            return v;
        }
        if (!t.compare(i, 5, "false")) {
            i += 5;
            v.kind = JsonValue::Bool;
            return v;
        }
        char* end = nullptr;
        v.num = strtod(t.c_str() + i, &end);
        if (end == t.c_str() + i) err("unexpected character");
        i = (size_t)(end - t.c_str());
        v.kind = JsonValue::Number;
        return v;
    }
};
} // namespace

const JsonValue* JsonValue::find(const std::string& key) const {
    for (auto& kv : obj)
        if (kv.first == key) return &kv.second;
    return nullptr;
}
JsonValue JsonValue::parse(const std::string& text) {
    JsonParser p(text);
    JsonValue v = p.value();
    p.ws();
    if (p.i != text.size()) p.err("trailing characters");
    return v;
}
std::string json_escape(const std::string& s) {
    std::string o;
    for (char c : s) {
        switch (c) {
        case '"': o += "\\\""; break;
        case '\\': o += "\\\\"; break;
        case '\n': o += "\\n"; break;
        case '\t': o += "\\t"; break;
        case '\r': o += "\\r"; break;
        default: o.push_back(c);
        }
    }
    return o;
}

} // namespace kfrt
