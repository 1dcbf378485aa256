// Implementation of http.hpp (the header carries the design notes and the reference file:line mapping).
#include "http.hpp"

#include <arpa/inet.h>
#include <netinet/in.h>
#include <poll.h>
#include <sys/socket.h>
#include <unistd.h>

#include <cstring>
#include <fstream>
#include <sstream>

#include "base.hpp"
#include "plan.hpp"
#include "transport.hpp"

namespace kfrt {

static bool read_http(Socket& s, std::string& head, std::string& body, int timeout_ms) {
    std::string buf;
    char tmp[4096];
    size_t hdr_end = std::string::npos;
    size_t content_len = 0;
    bool have_len = false;
    while (true) {
        if (hdr_end != std::string::npos && buf.size() >= hdr_end + 4 + content_len && have_len) break;
        pollfd p{s.fd(), POLLIN, 0};
        if (::poll(&p, 1, timeout_ms) != 1) return hdr_end != std::string::npos && !have_len;
        ssize_t k = ::recv(s.fd(), tmp, sizeof(tmp), 0);
        if (k < 0) return false;
        if (k == 0) {
            if (hdr_end == std::string::npos) return false;
            break; // connection closed: body is whatever arrived
        }
        buf.append(tmp, (size_t)k);
        // these endpoints (config server, monitors, metrics) exchange kilobytes: anything bigger is not ours, and nobody gets to make the
        // process buffer without bound
        constexpr size_t kMaxHead = 64u << 10, kMaxBody = 16u << 20;
        if (hdr_end == std::string::npos && buf.size() > kMaxHead && buf.find("\r\n\r\n") == std::string::npos) return false;
        if (hdr_end != std::string::npos && buf.size() > hdr_end + 4 + kMaxBody) return false;
        if (hdr_end == std::string::npos) {
            hdr_end = buf.find("\r\n\r\n");
            if (hdr_end != std::string::npos) {
                std::string h = buf.substr(0, hdr_end);
                std::string lower;
                for (char c : h) lower.push_back((char)tolower(c));
                auto pos = lower.find("content-length:");
                if (pos != std::string::npos) {
                    const long long cl = atoll(h.c_str() + pos + 15);
                    if (cl < 0 || (unsigned long long)cl > (16ull << 20)) return false;
                    content_len = (size_t)cl;
                    have_len = true;
                } else if (lower.compare(0, 4, "http") != 0) {
                    have_len = true; // a request without a body
                }
            }
        }
    }
    head = buf.substr(0, hdr_end);
    body = buf.size() > hdr_end + 4 ? buf.substr(hdr_end + 4) : "";
    if (have_len && body.size() > content_len) body.resize(content_len);
    return true;
}

HttpServer::~HttpServer() { stop(); }

int HttpServer::start(const std::string& host, int port, Handler h) {
    handler_ = std::move(h);
    lfd_ = ::socket(AF_INET, SOCK_STREAM | SOCK_CLOEXEC, 0);
    if (lfd_ < 0) fail("http: socket() failed");
    int one = 1;
    setsockopt(lfd_, SOL_SOCKET, SO_REUSEADDR, &one, sizeof(one));
    sockaddr_in a{};
    a.sin_family = AF_INET;
    a.sin_port = htons((uint16_t)port);
    a.sin_addr.s_addr = host.empty() || host == "0.0.0.0" ? htonl(INADDR_ANY) : htonl(resolve_ipv4(host));
    if (::bind(lfd_, reinterpret_cast<sockaddr*>(&a), sizeof(a)) != 0 || ::listen(lfd_, 128) != 0) {
        std::string e = strerror(errno);
        ::close(lfd_);
        lfd_ = -1;
        fail("http: cannot listen on " + host + ":" + std::to_string(port) + ": " + e);
    }
    socklen_t l = sizeof(a);
    getsockname(lfd_, reinterpret_cast<sockaddr*>(&a), &l);
    port_ = ntohs(a.sin_port);
    running_ = true;
    thread_ = std::thread([this] { loop(); });
    return port_;
}

void HttpServer::stop() {
    if (!running_.exchange(false)) return;
    if (lfd_ >= 0) ::shutdown(lfd_, SHUT_RDWR);
    if (thread_.joinable()) thread_.join();
    if (lfd_ >= 0) ::close(lfd_);
    lfd_ = -1;
}

void HttpServer::loop() {
    while (running_) {
        pollfd p{lfd_, POLLIN, 0};
        if (::poll(&p, 1, 200) <= 0) continue;
        int fd = ::accept4(lfd_, nullptr, nullptr, SOCK_CLOEXEC);
        if (fd < 0) continue;
        Handler h = handler_;
        Pool::get().submit([fd, h] {
            Socket s(fd);
            std::string head, body;
            if (!read_http(s, head, body, 5000)) return;
            HttpRequest req;
            std::istringstream hs(head);
            std::string line, target, ver;
            std::getline(hs, line);
            std::istringstream rl(line);
            rl >> req.method >> target >> ver;
            auto q = target.find('?');
            req.path = q == std::string::npos ? target : target.substr(0, q);
            req.query = q == std::string::npos ? "" : target.substr(q + 1);
            while (std::getline(hs, line)) {
                auto c = line.find(':');
                if (c == std::string::npos) continue;
                std::string k = line.substr(0, c), v = line.substr(c + 1);
                while (!v.empty() && (v.back() == '\r' || v.back() == ' ')) v.pop_back();
                while (!v.empty() && v.front() == ' ') v.erase(v.begin());
                for (auto& ch : k) ch = (char)tolower(ch);
                req.headers[k] = v;
            }
            req.body = body;
            HttpResponse resp;
            try {
                resp = h(req);
            } catch (const std::exception& e) {
                resp.status = 500;
                resp.body = e.what();
            }
            const char* reason = resp.status == 200 ? "OK" : resp.status == 404 ? "Not Found" : resp.status == 403 ? "Forbidden"
                                 : resp.status == 400 ? "Bad Request" : resp.status == 405 ? "Method Not Allowed" : "Error";
            std::ostringstream os;
            os << "HTTP/1.1 " << resp.status << " " << reason << "\r\nContent-Type: " << resp.content_type
               << "\r\nContent-Length: " << resp.body.size() << "\r\nConnection: close\r\n\r\n"
               << resp.body;
            const std::string out = os.str();
            s.send_all(out.data(), out.size());
        });
    }
}

Url Url::parse(const std::string& s) {
    Url u;
    auto p = s.find("://");
    if (p == std::string::npos) fail("invalid url '" + s + "'");
    u.scheme = s.substr(0, p);
    std::string rest = s.substr(p + 3);
    if (u.scheme == "file") {
        u.path = rest;
        return u;
    }
    auto slash = rest.find('/');
    std::string hostport = slash == std::string::npos ? rest : rest.substr(0, slash);
    u.path = slash == std::string::npos ? "/" : rest.substr(slash);
    auto colon = hostport.rfind(':');
    if (colon == std::string::npos) {
        u.host = hostport;
        u.port = u.scheme == "https" ? 443 : 80;
    } else {
        u.host = hostport.substr(0, colon);
        u.port = atoi(hostport.c_str() + colon + 1);
    }
    return u;
}

int http_request(const std::string& method, const std::string& url, const std::string& body, std::string* out,
                 const std::string& content_type, int timeout_ms) {
    Url u = Url::parse(url);
    if (u.scheme == "file") {
        if (method != "GET") return 405;
        std::ifstream f(u.path, std::ios::binary);
        if (!f) return 404;
        std::stringstream ss;
        ss << f.rdbuf();
        if (out) *out = ss.str();
        return 200;
    }
    if (u.scheme != "http") fail("unsupported url scheme '" + u.scheme + "'");
    uint32_t ip;
    try {
        ip = resolve_ipv4(u.host);
    } catch (...) {
        return -1;
    }
    Socket s = Socket::connect_tcp(ip, (uint16_t)u.port, timeout_ms);
    if (!s.valid()) return -1;
    std::ostringstream os;
    os << method << " " << u.path << " HTTP/1.1\r\nHost: " << u.host << ":" << u.port << "\r\nConnection: close\r\n";
    if (!body.empty() || method == "PUT" || method == "POST") os << "Content-Type: " << content_type << "\r\nContent-Length: " << body.size() << "\r\n";
    os << "\r\n" << body;
    const std::string req = os.str();
    if (!s.send_all(req.data(), req.size())) return -1;
    std::string head, rbody;
    if (!read_http(s, head, rbody, timeout_ms)) return -1;
    int status = -1;
    if (head.size() > 12) status = atoi(head.c_str() + 9);
    if (out) *out = rbody;
    return status;
}

} // namespace kfrt
