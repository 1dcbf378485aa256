// kungfu_b200 runtime — a minimal blocking HTTP/1.1 server and client.
//
// Used by the elastic config server (reference: srcs/go/kungfu/elastic/configserver/configserver.go:15-112), the
// failure-detector monitor server (srcs/go/kungfu/runner/monitorserver/monitor.go:17-198), the /metrics endpoint
// (srcs/go/monitor/server.go:15-30) and by peers fetching the cluster config (srcs/go/kungfu/peer/peer.go:236-263).
#pragma once
#include <atomic>
#include <functional>
#include <map>
#include <string>
#include <thread>
#include <vector>

namespace kfrt {

struct HttpRequest {
    std::string method, path, query, body;
    std::map<std::string, std::string> headers;
};
struct HttpResponse {
    int status = 200;
    std::string body;
    std::string content_type = "text/plain";
};

class HttpServer {
public:
    using Handler = std::function<HttpResponse(const HttpRequest&)>;
    HttpServer() = default;
    ~HttpServer();
    // port 0 picks a free port; returns the bound port. Throws on failure.
    int start(const std::string& host, int port, Handler h);
    void stop();
    int port() const { return port_; }
    bool running() const { return running_; }

private:
    void loop();
    Handler handler_;
    std::atomic<bool> running_{false};
    int lfd_ = -1;
    int port_ = 0;
    std::thread thread_;
};

struct Url {
    std::string scheme, host, path;
    int port = 80;
    static Url parse(const std::string& s); // http://host:port/path or file:///path
};

// Returns the HTTP status (or -1 on connection failure); body is filled on success. `file://` URLs are read from disk
// for GET (status 200/404).
int http_request(const std::string& method, const std::string& url, const std::string& body, std::string* out,
                 const std::string& content_type = "application/json", int timeout_ms = 3000);

} // namespace kfrt
