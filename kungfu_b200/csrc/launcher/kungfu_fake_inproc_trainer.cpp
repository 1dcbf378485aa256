// kungfu-fake-inproc-trainer — communication-only "trainer" whose N workers are threads of ONE process; no launcher needed.
//
//   kungfu-fake-inproc-trainer [-np 4] [-model resnet50|vgg16|bert|slp] [-fuse] [-backend sim|peers|both] [-steps 10] [-warmup 2] [-batch 32]
//
// Two back-ends behind the same harness ("Img/sec, communication only", the metric of the reference's fake trainers):
//   sim    all-reduce simulated through shared memory: worker r sums shard r of every worker's buffer (reduce-scatter), a thread barrier,
//          then everybody copies the full result (all-gather). No sockets, no framing: the upper bound a host-plane all-reduce can reach
//          on this machine, what the real runtime is compared against.
//   peers  N real `kfrt::Peer`s in one process talking over the loopback transport (Unix sockets) through `Session::all_reduce`.
//
// Reference: tests/cpp/integration/fake_in_proc_trainer.cpp:9-48 (in-process simulated all-reduce as a no-network upper bound) and
// fake_trainer.hpp:169-227 (run_experiment: warm-up + timed steps over the gradient sizes, per tensor or fused, prints Img/sec).
#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <unistd.h>
#include <vector>

#include "base.hpp"
#include "fake_model.hpp"
#include "peer.hpp"
#include "reduce.hpp"

using namespace kfrt;

namespace {

class ThreadBarrier {
public:
    explicit ThreadBarrier(int n) : n_(n) {}
    void wait() {
        std::unique_lock<std::mutex> lk(mu_);
        const uint64_t gen = gen_;
        if (++arrived_ == n_) {
            arrived_ = 0;
            ++gen_;
            cv_.notify_all();
        } else {
            cv_.wait(lk, [&] { return gen_ != gen; });
        }
    }

private:
    std::mutex mu_;
    std::condition_variable cv_;
    int n_, arrived_ = 0;
    uint64_t gen_ = 0;
};

struct Collective {
    virtual ~Collective() = default;
    virtual void all_reduce(int rank, const float* send, float* recv, size_t count, const std::string& name) = 0;
    virtual void barrier(int rank) = 0;
};

// shared-memory all-reduce: the send buffers of all workers are visible to all workers
class SimCollective : public Collective {
public:
    explicit SimCollective(int n) : n_(n), bar_(n), send_(n, nullptr), scratch_() {}
    void all_reduce(int rank, const float* send, float* recv, size_t count, const std::string&) override {
        send_[rank] = send;
        if (rank == 0 && scratch_.size() < count) scratch_.resize(count);
        bar_.wait();
        const auto part = even_partition(count, (size_t)n_, (size_t)rank);
        const size_t b = part.first, len = part.second - part.first;
        if (len) {
            std::memcpy(scratch_.data() + b, send_[0] + b, len * 4);
            for (int q = 1; q < n_; ++q) transform2(scratch_.data() + b, send_[q] + b, scratch_.data() + b, len, kF32, kSum);
        }
        bar_.wait();
        std::memcpy(recv, scratch_.data(), count * 4);
        bar_.wait(); // nobody reuses the scratch before everybody has copied
    }
    void barrier(int) override { bar_.wait(); }

private:
    int n_;
    ThreadBarrier bar_;
    std::vector<const float*> send_;
    std::vector<float> scratch_;
};

class PeerCollective : public Collective {
public:
    PeerCollective(int n, uint16_t base) {
        for (int r = 0; r < n; ++r) peers_.emplace_back(new Peer(PeerConfig::single_machine(r, n, base)));
        std::vector<std::thread> ts;
        std::vector<std::string> errs(n);
        for (int r = 0; r < n; ++r)
            ts.emplace_back([&, r] {
                try {
                    peers_[r]->start();
                } catch (const std::exception& e) {
                    errs[r] = e.what();
                }
            });
        for (auto& t : ts) t.join();
        for (auto& e : errs)
            if (!e.empty()) fail("peer start: " + e);
    }
    ~PeerCollective() override {
        std::vector<std::thread> ts;
        for (auto& p : peers_) ts.emplace_back([&p] { p->close(); });
        for (auto& t : ts) t.join();
    }
    void all_reduce(int rank, const float* send, float* recv, size_t count, const std::string& name) override {
        Workspace w{send, recv, count, kF32, kSum, name};
        peers_[rank]->session()->all_reduce(w);
    }
    void barrier(int rank) override { peers_[rank]->session()->barrier(); }

private:
    std::vector<std::unique_ptr<Peer>> peers_;
};

struct Result {
    double img_s = 0, ms_per_step = 0, gib_s = 0;
    bool ok = true;
};

Result run_experiment(Collective& coll, int np, const std::vector<size_t>& sizes, int warmup, int steps, int batch) {
    const size_t total = kftools::total_elements(sizes);
    std::vector<double> elapsed(np, 0.0);
    std::vector<int> good(np, 1);
    std::vector<std::string> errs(np);
    std::vector<std::thread> ts;
    for (int r = 0; r < np; ++r)
        ts.emplace_back([&, r] {
            try {
                std::vector<float> send(total, (float)(r + 1)), recv(total, 0.f);
                const float want = (float)(np * (np + 1) / 2);
                double t0 = 0;
                for (int it = 0; it < warmup + steps; ++it) {
                    if (it == warmup) {
                        coll.barrier(r);
                        t0 = now_s();
                    }
                    size_t off = 0;
                    for (size_t i = 0; i < sizes.size(); ++i) {
                        coll.all_reduce(r, send.data() + off, recv.data() + off, sizes[i], "g" + std::to_string(i) + "s" + std::to_string(it));
                        off += sizes[i];
                    }
                    if (recv[0] != want || recv[total - 1] != want || recv[total / 2] != want) good[r] = 0;
                }
                coll.barrier(r);
                elapsed[r] = now_s() - t0;
            } catch (const std::exception& e) {
                errs[r] = e.what();
                good[r] = 0;
            }
        });
    for (auto& t : ts) t.join();
    Result res;
    double worst = 0;
    for (int r = 0; r < np; ++r) {
        worst = std::max(worst, elapsed[r]);
        if (!good[r]) res.ok = false;
        if (!errs[r].empty()) std::fprintf(stderr, "worker %d: %s\n", r, errs[r].c_str());
    }
    res.ms_per_step = worst / steps * 1e3;
    res.img_s = (double)batch * np * steps / worst;                                      // whole "cluster"
    res.gib_s = 4.0 * (np - 1) * (double)total * 4.0 * steps / worst / (1 << 30) / std::max(np, 1); // per worker equivalent data rate
    return res;
}

} // namespace

int main(int argc, char** argv) {
    kftools::Flags fl{argc, argv};
    const int np = (int)fl.num("np", 4), steps = (int)fl.num("steps", 10), warmup = (int)fl.num("warmup", 2), batch = (int)fl.num("batch", 32);
    const std::string model = fl.str("model", "resnet50"), backend = fl.str("backend", "both");
    auto sizes = kftools::fake_model_sizes(model);
    if (sizes.empty() || np < 1 || np > 64) {
        std::fprintf(stderr, "usage: kungfu-fake-inproc-trainer [-np 1..64] [-model resnet50|vgg16|bert|slp] [-fuse] [-backend sim|peers|both] [-steps n] [-warmup n]\n");
        return 2;
    }
    if (fl.flag("fuse")) sizes = {kftools::total_elements(sizes)};
    const size_t total = kftools::total_elements(sizes);
    std::printf("model: %s, %zu tensors, %.2f MiB, np=%d (threads of one process), %d + %d steps, batch %d per worker\n", model.c_str(), sizes.size(),
                total * 4.0 / (1 << 20), np, warmup, steps, batch);
    int rc = 0;
    try {
        if (backend == "sim" || backend == "both") {
            SimCollective sim(np);
            const Result r = run_experiment(sim, np, sizes, warmup, steps, batch);
            std::printf("Img/sec %.1f  [sim: shared-memory all-reduce, upper bound]  %.3f ms/step, %.2f GiB/s equivalent per worker%s\n", r.img_s,
                        r.ms_per_step, r.gib_s, r.ok ? "" : "  [WRONG RESULT]");
            if (!r.ok) rc = 3;
        }
        if (backend == "peers" || backend == "both") {
            const uint16_t base = (uint16_t)fl.num("base-port", 41000 + (getpid() % 2000) * 8 % 16000);
            PeerCollective peers(np, base);
            const Result r = run_experiment(peers, np, sizes, warmup, steps, batch);
            std::printf("Img/sec %.1f  [peers: kfrt::Session over the loopback transport]  %.3f ms/step, %.2f GiB/s equivalent per worker%s\n", r.img_s,
                        r.ms_per_step, r.gib_s, r.ok ? "" : "  [WRONG RESULT]");
            if (!r.ok) rc = 3;
        }
    } catch (const std::exception& e) {
        std::fprintf(stderr, "kungfu-fake-inproc-trainer: %s\n", e.what());
        return 1;
    }
    return rc;
}
