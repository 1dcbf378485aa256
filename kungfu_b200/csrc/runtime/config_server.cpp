// Implementation of config_server.hpp (the header carries the design notes and the reference file:line mapping).
#include "config_server.hpp"

#include "base.hpp"

namespace kfrt {

int ConfigServer::start(const std::string& host, int port, const Cluster* init) {
    if (init) {
        cluster_ = *init;
        has_ = true;
        version_ = 1;
    }
    return http_.start(host, port, [this](const HttpRequest& r) { return handle(r); });
}

void ConfigServer::stop() {
    {
        std::lock_guard<std::mutex> g(mu_);
        stopped_ = true;
    }
    cv_.notify_all();
    http_.stop();
}

void ConfigServer::wait_stopped() {
    std::unique_lock<std::mutex> lk(mu_);
    cv_.wait(lk, [this] { return stopped_; });
}

int ConfigServer::version() {
    std::lock_guard<std::mutex> g(mu_);
    return version_;
}
bool ConfigServer::get(Cluster& out) {
    std::lock_guard<std::mutex> g(mu_);
    if (!has_) return false;
    out = cluster_;
    return true;
}

HttpResponse ConfigServer::handle(const HttpRequest& r) {
    HttpResponse resp;
    if (r.path == "/stop") {
        {
            std::lock_guard<std::mutex> g(mu_);
            stopped_ = true;
        }
        cv_.notify_all();
        resp.body = "stopping\n";
        return resp;
    }
    // The endpoints of the reference's older example server, which its elastic test script and compose generator still call
    // (tests/go/cmd/kungfu-config-server-example/kungfu-config-server-example.go:36-44 `/get /put /clear /reset`;
    // scripts/tests/run-elastic-test.sh:21, benchmarks/adaptation/gen-compose.py:84,102): same state, the verb is in the path.
    std::string method = r.method;
    bool legacy = true;
    if (r.path == "/get") method = "GET";
    else if (r.path == "/put") method = "PUT";
    else if (r.path == "/reset") method = "POST";
    else if (r.path == "/clear") method = "CLEAR";
    else legacy = false;
    if (!legacy && r.path != path_) {
        resp.status = 404;
        resp.body = "not found\n";
        return resp;
    }
    std::lock_guard<std::mutex> g(mu_);
    if (method == "CLEAR") {
        // an EMPTY cluster, not "no cluster": workers that follow the server all detach, and later updates are refused (403) until a reset
        cluster_.workers = PeerList();
        has_ = true;
        ++version_;
        KF_LOGW("config server: config cleared (every worker will leave)");
        return resp;
    }
    if (method == "GET") {
        resp.content_type = "application/json";
        if (!has_) {
            resp.status = 404;
            resp.body = "No Config Found.\n";
        } else {
            resp.body = cluster_.to_json() + "\n";
        }
    } else if (method == "PUT") {
        Cluster c;
        try {
            c = Cluster::from_json(r.body);
            c.validate();
        } catch (const std::exception& e) {
            KF_LOGE("config server: rejected PUT: %s", e.what());
            resp.status = 400;
            resp.body = std::string("invalid cluster config: ") + e.what() + "\n";
            return resp;
        }
        if (!has_) {
            KF_LOGI("config server: init first config to %d peers: %s", (int)c.workers.size(), c.workers.str().c_str());
            version_ = 1;
            cluster_ = c;
            has_ = true;
        } else if (!cluster_.workers.empty()) {
            ++version_;
            cluster_ = c;
            KF_LOGI("config server: updated to %d peers: %s", (int)c.workers.size(), c.workers.str().c_str());
        } else {
            KF_LOGI("config server: config was cleared, update rejected");
            resp.status = 403;
            resp.body = "config was cleared\n";
        }
    } else if (method == "POST") {
        has_ = false;
        cluster_ = Cluster();
        KF_LOGI("config server: reset config");
    } else if (method == "DELETE") {
        has_ = false;
        cluster_ = Cluster();
        KF_LOGW("config server: config deleted!");
    } else {
        resp.status = 405;
        resp.body = "unsupported method\n";
    }
    return resp;
}

} // namespace kfrt
