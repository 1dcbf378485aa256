// kungfu-config-server — standalone elastic config server.
// Reference: srcs/go/cmd/kungfu-config-server/kungfu-config-server.go:19-65 (flags -host -port(9100) -init -ttl -endpoint).
#include <signal.h>

#include <cstring>
#include <fstream>
#include <sstream>

#include "base.hpp"
#include "config_server.hpp"

using namespace kfrt;

static ConfigServer* g_srv = nullptr;
static void on_sig(int) {
    if (g_srv) g_srv->stop();
}

int main(int argc, char** argv) {
    std::string host = "0.0.0.0", init, endpoint = "/config";
    int port = 9100;
    double ttl = 0;
    for (int i = 1; i < argc; ++i) {
        std::string a = argv[i];
        auto val = [&](const char* name) -> std::string {
            if (i + 1 >= argc) {
                fprintf(stderr, "flag %s needs a value\n", name);
                exit(2);
            }
            return argv[++i];
        };
        if (a == "-host" || a == "--host") host = val("-host");
        else if (a == "-port" || a == "--port") port = atoi(val("-port").c_str());
        else if (a == "-init" || a == "--init") init = val("-init");
        else if (a == "-ttl" || a == "--ttl") ttl = parse_duration_s(val("-ttl"), 0);
        else if (a == "-endpoint" || a == "--endpoint") endpoint = val("-endpoint");
        else {
            fprintf(stderr, "usage: kungfu-config-server [-host h] [-port 9100] [-init cluster.json] [-ttl d] [-endpoint /config]\n");
            return a == "-h" || a == "--help" ? 0 : 2;
        }
    }
    try {
        ConfigServer srv(endpoint);
        g_srv = &srv;
        signal(SIGINT, on_sig);
        signal(SIGTERM, on_sig);
        Cluster c;
        bool has = false;
        if (!init.empty()) {
            std::ifstream in(init);
            if (!in) fail("cannot read " + init);
            std::stringstream ss;
            ss << in.rdbuf();
            c = Cluster::from_json(ss.str());
            c.validate();
            has = true;
        }
        const int p = srv.start(host, port, has ? &c : nullptr);
        KF_LOGI("config server listening on http://%s:%d%s", host.c_str(), p, endpoint.c_str());
        if (ttl > 0)
            std::thread([&srv, ttl] {
                sleep_ms((int)(ttl * 1000));
                srv.stop();
            }).detach();
        srv.wait_stopped();
        srv.stop();
        return 0;
    } catch (const std::exception& e) {
        fprintf(stderr, "kungfu-config-server: %s\n", e.what());
        return 1;
    }
}
