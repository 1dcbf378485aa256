// kungfu_b200 — Python bindings of the CPU runtime (`kungfu_b200._kfrt`).
//
// Counterpart of the reference's C ABI (38 `GoKungfu*` cgo exports, srcs/go/libkungfu-comm/{main,collective,adapt,
// monitoring}.go) + `kungfu::Peer` C++ facade (srcs/cpp/include/kungfu/peer.hpp:13-163) + python C API
// (srcs/cpp/include/kungfu/python/c_api.h:8-54). Async contract: `*_async` returns an integer handle; the operation
// runs on the runtime's elastic pool; `wait(handle)` blocks on a condition variable (no busy-waiting, unlike
// `HandleManager`, srcs/cpp/include/kungfu/utils/handler_manager.hpp:40-83). Blocking calls release the GIL.
#include <pybind11/functional.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include <condition_variable>
#include <map>
#include <mutex>

#include "base.hpp"
#include "config_server.hpp"
#include "graph.hpp"
#include "http.hpp"
#include "numa.hpp"
#include "peer.hpp"
#include "plan.hpp"
#include "reduce.hpp"
#include "session.hpp"

namespace py = pybind11;
using namespace kfrt;

namespace {

struct Handles {
    struct State {
        bool done = false;
        std::string err;
    };
    std::mutex mu;
    std::condition_variable cv;
    std::map<int64_t, State> st;
    int64_t next = 1;
    int64_t create() {
        std::lock_guard<std::mutex> g(mu);
        st[next] = State();
        return next++;
    }
    void finish(int64_t h, const std::string& err) {
        std::lock_guard<std::mutex> g(mu);
        auto it = st.find(h);
        if (it != st.end()) {
            it->second.done = true;
            it->second.err = err;
        }
        cv.notify_all();
    }
    void wait(int64_t h) {
        std::unique_lock<std::mutex> lk(mu);
        auto it = st.find(h);
        if (it == st.end()) return; // already waited
        cv.wait(lk, [&] { return st[h].done; });
        const std::string err = st[h].err;
        st.erase(h);
        lk.unlock();
        if (!err.empty()) throw Error(err);
    }
};
Handles& handles() {
    static Handles* h = new Handles();
    return *h;
}

template <typename F> int64_t run_async(F&& fn) {
    const int64_t h = handles().create();
    Pool::get().submit([h, fn] {
        std::string err;
        try {
            fn();
        } catch (const std::exception& e) {
            err = e.what();
        }
        handles().finish(h, err);
    });
    return h;
}

Workspace mk(uintptr_t send, uintptr_t recv, size_t count, int dtype, int op, const std::string& name) {
    Workspace w;
    w.send = reinterpret_cast<const void*>(send);
    w.recv = reinterpret_cast<void*>(recv);
    w.count = count;
    w.dtype = dtype;
    w.op = op;
    w.name = name;
    return w;
}

py::dict cluster_to_dict(const Cluster& c) {
    py::dict d;
    py::list r, w;
    for (auto& p : c.runners) r.append(p.str());
    for (auto& p : c.workers) w.append(p.str());
    d["runners"] = r;
    d["workers"] = w;
    return d;
}

py::list graph_edges(const Graph& g) {
    py::list e;
    for (int i = 0; i < g.size(); ++i)
        for (int j : g.nexts(i)) e.append(py::make_tuple(i, j));
    return e;
}

py::dict graph_to_dict(const Graph& g) {
    py::dict d;
    d["n"] = g.size();
    d["edges"] = graph_edges(g);
    py::list loops;
    for (int i = 0; i < g.size(); ++i)
        if (g.is_self_loop(i)) loops.append(i);
    d["self_loops"] = loops;
    d["digest"] = py::bytes(g.digest());
    d["debug"] = g.debug_string();
    return d;
}

} // namespace

PYBIND11_MODULE(_kfrt, m) {
    m.doc() = "kungfu_b200 CPU runtime: transport, graph collectives, peer/elastic membership, config server";
    py::register_exception<Error>(m, "RuntimeError_", PyExc_RuntimeError);

    // ---- dtype / op / strategies
    m.def("dtype_size", [](int dt) { return dtype_size(dt); });
    m.def("strategy_names", &strategy_names);
    m.def("transform2", [](uintptr_t dst, uintptr_t x, uintptr_t y, size_t count, int dtype, int op) {
        transform2(reinterpret_cast<void*>(dst), reinterpret_cast<const void*>(x), reinterpret_cast<const void*>(y), count, dtype, op);
    });
    m.def("half_to_float", &half_to_float);
    m.def("float_to_half", &float_to_half);
    m.def("bf16_to_float", &bf16_to_float);
    m.def("float_to_bf16", &float_to_bf16);
    m.def("reload_config", [] { Config::reload(); });
    m.def("set_log_level", [](const std::string& l) { Logger::get().set_level(l); });

    // ---- plan
    m.def("parse_peer_list", [](const std::string& s) {
        py::list out;
        for (auto& p : PeerList::parse(s)) out.append(p.str());
        return out;
    });
    m.def("gen_peer_list", [](const std::string& hosts, int np, const std::string& port_range) {
        py::list out;
        for (auto& p : HostList::parse(hosts).gen_peer_list(np, PortRange::parse(port_range))) out.append(p.str());
        return out;
    }, py::arg("hosts"), py::arg("np"), py::arg("port_range") = "10000-11000");
    m.def("gen_runner_list", [](const std::string& hosts, int port) {
        py::list out;
        for (auto& p : HostList::parse(hosts).gen_runner_list((uint16_t)port)) out.append(p.str());
        return out;
    }, py::arg("hosts"), py::arg("port") = (int)kDefaultRunnerPort);
    m.def("host_list_cap", [](const std::string& hosts) { return HostList::parse(hosts).cap(); });
    m.def("parse_hostfile", [](const std::string& text) { return HostList::parse_hostfile(text).str(); });
    m.def("peer_list_info", [](const std::string& peers, const std::string& self) {
        PeerList pl = PeerList::parse(peers);
        PeerID me = PeerID::parse(self);
        std::vector<int> masters, parents;
        pl.partition_by_host(masters, parents);
        py::dict d;
        d["rank"] = pl.rank_of(me);
        d["local_rank"] = pl.local_rank_of(me);
        d["local_size"] = pl.local_size_of(me);
        d["host_count"] = pl.host_count();
        d["masters"] = masters;
        d["parents"] = parents;
        return d;
    });
    m.def("cluster_json", [](const std::string& runners, const std::string& workers) {
        Cluster c{PeerList::parse(runners), PeerList::parse(workers)};
        return c.to_json();
    });
    m.def("cluster_from_json", [](const std::string& s) { return cluster_to_dict(Cluster::from_json(s)); });
    m.def("cluster_validate", [](const std::string& s) { Cluster::from_json(s).validate(); });
    m.def("cluster_resize", [](const std::string& s, int n) { return Cluster::from_json(s).resized(n).to_json(); });
    m.def("stage_json", [](int version, const std::string& cluster_json) {
        Stage st;
        st.version = version;
        st.cluster = Cluster::from_json(cluster_json);
        return st.to_json();
    });
    m.def("stage_from_json", [](const std::string& s) {
        Stage st = Stage::from_json(s);
        return py::make_tuple(st.version, st.cluster.to_json());
    });
    m.def("even_partition", [](size_t n, size_t k, size_t i) { return even_partition(n, k, i); });

    // ---- graphs
    m.def("gen_strategy_graphs", [](const std::string& peers, const std::string& strategy, const std::string& scope) {
        PeerList pl = PeerList::parse(peers);
        std::vector<GraphPair> sl;
        if (scope == "local") sl = gen_local_strategies(pl);
        else if (scope == "cross") sl = gen_cross_strategies(pl, parse_strategy(strategy));
        else sl = gen_global_strategies(pl, parse_strategy(strategy));
        py::list out;
        for (auto& p : sl) out.append(py::make_tuple(graph_to_dict(p.reduce), graph_to_dict(p.bcast)));
        return out;
    }, py::arg("peers"), py::arg("strategy") = "AUTO", py::arg("scope") = "global");
    m.def("is_rooted_tree", [](int n, const std::vector<std::pair<int, int>>& edges, int root) {
        Graph g(n);
        for (auto& e : edges) g.add_edge(e.first, e.second);
        return g.is_rooted_tree(root);
    });
    m.def("forest_to_graph", [](const std::vector<int>& father) {
        Graph g;
        int roots = 0;
        const bool ok = Graph::from_forest_array(father, g, roots);
        return py::make_tuple(ok, roots, ok ? graph_to_dict(g) : py::dict());
    });
    m.def("minimum_spanning_tree", [](const std::vector<double>& w, int n) { return minimum_spanning_tree(w, n); });
    m.def("mst_forest_array", [](const std::vector<double>& w, int n, int root) { return mst_forest_array(w, n, root); }, py::arg("weights"),
          py::arg("n"), py::arg("root") = 0);

    // ---- numa
    m.def("numa_topology", &numa_topology);
    m.def("numa_topology_ordered", &numa_topology_ordered);
    m.def("select_cpus", &select_cpus);
    m.def("apply_affinity_from_env", &apply_affinity_from_env);

    // ---- http / config server
    m.def("metric_set", [](const std::string& name, double v) { NetMonitor::get().set_gauge(name, v); });
    m.def("metric_add", [](const std::string& name, double d) { NetMonitor::get().add_counter(name, d); });
    m.def("metrics_text", []() { return NetMonitor::get().metrics_text(); });
    m.def("http_request", [](const std::string& method, const std::string& url, const std::string& body) {
        std::string out;
        int st;
        {
            py::gil_scoped_release rel;
            st = http_request(method, url, body, &out);
        }
        return py::make_tuple(st, py::bytes(out));
    }, py::arg("method"), py::arg("url"), py::arg("body") = "");

    py::class_<ConfigServer>(m, "ConfigServer")
        .def(py::init<const std::string&>(), py::arg("path") = "/config")
        .def("start", [](ConfigServer& s, const std::string& host, int port, const std::string& init_json) {
            if (init_json.empty()) return s.start(host, port, nullptr);
            Cluster c = Cluster::from_json(init_json);
            return s.start(host, port, &c);
        }, py::arg("host") = "127.0.0.1", py::arg("port") = 0, py::arg("init_json") = "")
        .def("stop", &ConfigServer::stop, py::call_guard<py::gil_scoped_release>())
        .def("wait_stopped", &ConfigServer::wait_stopped, py::call_guard<py::gil_scoped_release>())
        .def_property_readonly("port", &ConfigServer::port)
        .def_property_readonly("url", &ConfigServer::url)
        .def("version", &ConfigServer::version)
        .def("get", [](ConfigServer& s) -> py::object {
            Cluster c;
            if (!s.get(c)) return py::none();
            return py::str(c.to_json());
        });

    // ---- monitor
    m.def("metrics_text", [] { return NetMonitor::get().metrics_text(); });

    // ---- peer
    py::class_<Peer>(m, "Peer")
        .def_static("from_env", [] { return new Peer(PeerConfig::from_env()); }, py::return_value_policy::take_ownership)
        .def_static("single_machine", [](int rank, int size, int base_port) { return new Peer(PeerConfig::single_machine(rank, size, (uint16_t)base_port)); },
                    py::arg("rank"), py::arg("size"), py::arg("base_port") = 0, py::return_value_policy::take_ownership)
        .def_static("explicit", [](const std::string& self, const std::string& peers, const std::string& runners, const std::string& strategy,
                                   int version, const std::string& config_server) {
            PeerConfig c;
            c.launcher = "explicit";
            c.self = PeerID::parse(self);
            c.init_peers = PeerList::parse(peers);
            c.init_runners = PeerList::parse(runners);
            c.strategy = parse_strategy(strategy);
            c.init_cluster_version = version;
            c.config_server = config_server;
            c.single = c.init_peers.size() <= 1 && runners.empty();
            return new Peer(c);
        }, py::arg("self"), py::arg("peers"), py::arg("runners") = "", py::arg("strategy") = "AUTO", py::arg("version") = 0,
                    py::arg("config_server") = "", py::return_value_policy::take_ownership)
        .def("start", &Peer::start, py::call_guard<py::gil_scoped_release>())
        .def("close", &Peer::close, py::call_guard<py::gil_scoped_release>())
        .def("rank", [](Peer& p) { return p.session()->rank(); })
        .def("size", [](Peer& p) { return p.session()->size(); })
        .def("local_rank", [](Peer& p) { return p.session()->local_rank(); })
        .def("local_size", [](Peer& p) { return p.session()->local_size(); })
        .def("host_count", [](Peer& p) { return p.session()->host_count(); })
        .def("peers", [](Peer& p) {
            py::list out;
            for (auto& q : p.session()->peers()) out.append(q.str());
            return out;
        })
        .def("uid", &Peer::uid)
        .def("detached", &Peer::detached)
        .def("single", &Peer::single)
        .def("cluster_version", &Peer::cluster_version)
        .def("launcher", [](Peer& p) { return p.config().launcher; })
        .def("self_spec", [](Peer& p) { return p.self().str(); })
        .def("config_server", [](Peer& p) { return p.config().config_server; })
        .def("current_cluster", [](Peer& p) { return p.current_cluster().to_json(); })
        .def("barrier", [](Peer& p) { p.session()->barrier(); }, py::call_guard<py::gil_scoped_release>())
        .def("consensus", [](Peer& p, const std::string& data, const std::string& name) {
            py::gil_scoped_release rel;
            return p.session()->consensus(data.data(), data.size(), name);
        }, py::arg("data"), py::arg("name") = "")
        .def("all_reduce", [](Peer& p, uintptr_t send, uintptr_t recv, size_t count, int dtype, int op, const std::string& name) {
            py::gil_scoped_release rel;
            p.session()->all_reduce(mk(send, recv, count, dtype, op, name));
        })
        .def("all_reduce_async", [](Peer& p, uintptr_t send, uintptr_t recv, size_t count, int dtype, int op, const std::string& name) {
            auto sess = p.session();
            Workspace w = mk(send, recv, count, dtype, op, name);
            return run_async([sess, w] { sess->all_reduce(w); });
        })
        .def("monitored_all_reduce", [](Peer& p, uintptr_t send, uintptr_t recv, size_t count, int dtype, int op, const std::string& name,
                                        const std::vector<int32_t>& tree) {
            py::gil_scoped_release rel;
            p.session()->monitored_all_reduce(mk(send, recv, count, dtype, op, name), tree);
        })
        .def("subset_all_reduce", [](Peer& p, uintptr_t send, uintptr_t recv, size_t count, int dtype, int op, const std::string& name,
                                     const std::vector<int32_t>& forest) {
            py::gil_scoped_release rel;
            p.session()->subset_all_reduce(mk(send, recv, count, dtype, op, name), forest);
        })
        .def("cross_all_reduce", [](Peer& p, uintptr_t send, uintptr_t recv, size_t count, int dtype, int op, const std::string& name) {
            py::gil_scoped_release rel;
            p.session()->cross_all_reduce(mk(send, recv, count, dtype, op, name));
        })
        .def("reduce", [](Peer& p, uintptr_t send, uintptr_t recv, size_t count, int dtype, int op, const std::string& name) {
            py::gil_scoped_release rel;
            p.session()->reduce(mk(send, recv, count, dtype, op, name));
        })
        .def("broadcast", [](Peer& p, uintptr_t send, uintptr_t recv, size_t count, int dtype, const std::string& name) {
            py::gil_scoped_release rel;
            p.session()->broadcast(mk(send, recv, count, dtype, kSum, name));
        })
        .def("broadcast_async", [](Peer& p, uintptr_t send, uintptr_t recv, size_t count, int dtype, const std::string& name) {
            auto sess = p.session();
            Workspace w = mk(send, recv, count, dtype, kSum, name);
            return run_async([sess, w] { sess->broadcast(w); });
        })
        .def("local_reduce", [](Peer& p, uintptr_t send, uintptr_t recv, size_t count, int dtype, int op, const std::string& name) {
            py::gil_scoped_release rel;
            p.session()->local_reduce(mk(send, recv, count, dtype, op, name));
        })
        .def("local_broadcast", [](Peer& p, uintptr_t send, uintptr_t recv, size_t count, int dtype, const std::string& name) {
            py::gil_scoped_release rel;
            p.session()->local_broadcast(mk(send, recv, count, dtype, kSum, name));
        })
        .def("gather", [](Peer& p, uintptr_t send, uintptr_t recv, size_t count, int dtype, const std::string& name) {
            py::gil_scoped_release rel;
            p.session()->gather(mk(send, recv, count, dtype, kSum, name));
        })
        .def("all_gather", [](Peer& p, uintptr_t send, uintptr_t recv, size_t count, int dtype, const std::string& name) {
            py::gil_scoped_release rel;
            p.session()->all_gather(mk(send, recv, count, dtype, kSum, name));
        })
        .def("all_gather_bytes", [](Peer& p, const std::string& mine, const std::string& name) {
            std::vector<std::string> parts;
            {
                py::gil_scoped_release rel;
                auto sess = p.session();
                const int n = sess->size();
                const std::string tag = name + "#" + std::to_string(sess->next_seq());
                std::vector<int64_t> lens(n, 0);
                int64_t len = (int64_t)mine.size();
                sess->all_gather(mk((uintptr_t)&len, (uintptr_t)lens.data(), 1, kI64, kSum, "kungfu::agb::len::" + tag));
                int64_t mx = 0;
                for (auto l : lens) mx = l > mx ? l : mx;
                std::string padded = mine;
                padded.resize((size_t)mx, '\0');
                std::string all((size_t)mx * n, '\0');
                if (mx > 0) sess->all_gather(mk((uintptr_t)padded.data(), (uintptr_t)&all[0], (size_t)mx, kU8, kSum, "kungfu::agb::data::" + tag));
                for (int i = 0; i < n; ++i) parts.push_back(all.substr((size_t)i * mx, (size_t)lens[i]));
            }
            py::list out;
            for (auto& s : parts) out.append(py::bytes(s));
            return out;
        }, py::arg("data"), py::arg("name") = "")
        .def("all_reduce_int_max", [](Peer& p, int64_t x) {
            py::gil_scoped_release rel;
            int64_t y = x;
            auto sess = p.session();
            sess->all_reduce(mk((uintptr_t)&x, (uintptr_t)&y, 1, kI64, kMax, "kungfu::int_max::" + std::to_string(sess->next_seq())));
            return y;
        })
        .def("set_tree", [](Peer& p, const std::vector<int32_t>& father) { p.session()->set_tree(father); }, py::call_guard<py::gil_scoped_release>())
        .def("num_strategies", [](Peer& p) { return p.session()->num_strategies(); })
        .def("peer_latencies", [](Peer& p) { return p.session()->peer_latencies(); }, py::call_guard<py::gil_scoped_release>())
        .def("egress_rates", [](Peer& p) { return p.session()->egress_rates(); })
        .def("calc_stats", [](Peer& p) { p.session()->calc_stats(); })
        .def("log_stats", [](Peer& p) { p.session()->log_stats(); })
        .def("print_strategy_stats", [](Peer& p) { p.session()->print_strategy_stats(); })
        .def("check_interference", [](Peer& p) { return p.session()->check_interference(); }, py::call_guard<py::gil_scoped_release>())
        .def("last_throughput", [](Peer& p) { return p.session()->last_throughput(); })
        .def("resize_cluster", &Peer::resize_cluster, py::call_guard<py::gil_scoped_release>())
        .def("resize_cluster_from_url", &Peer::resize_cluster_from_url, py::call_guard<py::gil_scoped_release>())
        .def("recover", &Peer::recover, py::arg("timeout_s") = 30.0, py::call_guard<py::gil_scoped_release>())
        .def("failure_pending", &Peer::failure_pending)
        .def("set_lost_callback", [](Peer& p, py::object fn) {
            if (fn.is_none()) {
                p.set_lost_callback(nullptr);
                return;
            }
            // the callable is invoked from a transport thread: take the GIL there, and never destroy it without the GIL
            auto holder = std::shared_ptr<py::object>(new py::object(fn), [](py::object* o) {
                if (!Py_IsInitialized()) return; // interpreter is gone: leak the reference rather than touch it
                py::gil_scoped_acquire g;
                delete o;
            });
            p.set_lost_callback([holder] {
                if (!Py_IsInitialized()) return;
                py::gil_scoped_acquire g;
                (*holder)();
            });
        })
        .def("propose_new_size", &Peer::propose_new_size, py::call_guard<py::gil_scoped_release>())
        .def("propose", [](Peer& p, const std::string& cluster_json) { return p.propose(Cluster::from_json(cluster_json)); },
             py::call_guard<py::gil_scoped_release>())
        .def("update", &Peer::update, py::call_guard<py::gil_scoped_release>())
        .def("save", [](Peer& p, const std::string& name, uintptr_t data, size_t len) { p.save(name, reinterpret_cast<const void*>(data), len); })
        .def("save_version", [](Peer& p, const std::string& version, const std::string& name, uintptr_t data, size_t len) {
            p.save_version(version, name, reinterpret_cast<const void*>(data), len);
        })
        .def("request", [](Peer& p, int rank, const std::string& version, const std::string& name, uintptr_t buf, size_t len) {
            py::gil_scoped_release rel;
            return p.request(rank, version, name, reinterpret_cast<void*>(buf), len);
        })
        .def("send_control", [](Peer& p, const std::string& target, const std::string& name, const std::string& payload) {
            py::gil_scoped_release rel;
            p.client().send(PeerID::parse(target), name, payload.data(), payload.size(), kConnControl, kNoFlag);
        })
        .def("ping", [](Peer& p, const std::string& target) { return p.client().ping(PeerID::parse(target)); }, py::call_guard<py::gil_scoped_release>());

    // fire-and-forget control message from a process that is not a peer (≙ kungfu-test-util -kill / -propose,
    // tests/go/cmd/kungfu-test-util/kungfu-test-util.go:28-35, and kungfu-notify-start)
    m.def("send_control", [](const std::string& target, const std::string& name, const std::string& payload) {
        py::gil_scoped_release rel;
        Client c(PeerID{parse_ipv4("127.0.0.1"), 0});
        c.send(PeerID::parse(target), name, payload.data(), payload.size(), kConnControl, kNoFlag);
        c.close_all();
    }, py::arg("target"), py::arg("name"), py::arg("payload") = "");
    m.def("ping", [](const std::string& target) {
        py::gil_scoped_release rel;
        Client c(PeerID{parse_ipv4("127.0.0.1"), 0});
        return c.ping(PeerID::parse(target));
    });
    m.def("trace_report", [] { return trace_report(); });

    m.def("wait", [](int64_t h) { handles().wait(h); }, py::call_guard<py::gil_scoped_release>());
    m.def("wait_all", [](const std::vector<int64_t>& hs) {
        for (auto h : hs) handles().wait(h);
    }, py::call_guard<py::gil_scoped_release>());
    m.def("pool_threads", [] { return Pool::get().threads(); });
}
