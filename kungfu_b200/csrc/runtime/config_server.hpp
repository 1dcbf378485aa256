// kungfu_b200 runtime — elastic config server: holds the DESIRED cluster (JSON) and a version.
//
// Reference: srcs/go/kungfu/elastic/configserver/configserver.go:15-112 (+ standalone `kungfu-config-server`, builtin
// mode `kungfu-run -builtin-config-port`). REST: GET <path> -> cluster JSON (404 when none); PUT -> validate & replace
// (403 once the config has been cleared to empty); POST -> reset to "none"; DELETE -> clear; GET /stop -> shut down.
#pragma once
#include <condition_variable>
#include <memory>
#include <mutex>
#include <string>

#include "http.hpp"
#include "plan.hpp"

namespace kfrt {

class ConfigServer {
public:
    explicit ConfigServer(const std::string& path = "/config") : path_(path) {}
    int start(const std::string& host, int port, const Cluster* init); // returns the bound port
    void stop();
    void wait_stopped(); // blocks until /stop is requested (or stop() is called)
    int version();
    bool get(Cluster& out);
    int port() const { return http_.port(); }
    std::string url() const { return "http://127.0.0.1:" + std::to_string(http_.port()) + path_; }

private:
    HttpResponse handle(const HttpRequest& r);
    std::string path_;
    HttpServer http_;
    std::mutex mu_;
    std::condition_variable cv_;
    bool has_ = false, stopped_ = false;
    Cluster cluster_;
    int version_ = 0;
};

} // namespace kfrt
