// kungfu-bench-allreduce — model-shaped micro-benchmark of the host-plane all-reduce (the collective under every optimizer when the tensors
// live in host memory, and the cross-host step of a job that spans machines).
//
//   kungfu-run -np 4 [-strategy RING] kungfu-bench-allreduce [-model resnet50|vgg16|bert|slp] [-mode seq|par] [-fuse] [-epochs 15]
//                                                            [-warmup 2] [-dtype f32|f16|bf16|i32] [-rand-failure P]
//
// Reference: tests/go/cmd/kungfu-bench-allreduce/kungfu-bench-allreduce.go:18-120 — one all-reduce per tensor of the fake model, issued one
// after the other ("seq") or all at once ("par"), `-fuse` = one tensor of the total size, rate = epochs * 4 (np - 1) * model bytes / t
// (the bytes a ring moves in and out of every worker). Differences: written against the C++ Peer; every epoch's result is VERIFIED
// (worker r contributes r + 1 in every element, so the sum must be np (np + 1) / 2 — a mis-routed chunk or a tensor reduced under the wrong
// name shows up), the element type is selectable, and the reference's `-rand-nccl-failure` (a worker that dies at start-up with the
// probability of a known driver bug, to exercise the launcher's retry) is `-rand-failure P`: with probability P this worker prints the
// loader's message and exits 1 before it joins — `kungfu-run` starts such a worker again (start-up retry, kungfu_run.cpp `should_restart`).
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include <thread>
#include <vector>

#include <unistd.h>

#include "base.hpp"
#include "fake_model.hpp"
#include "peer.hpp"
#include "reduce.hpp"

using namespace kfrt;

static std::string show_rate(double bytes, double seconds) {
    const char* units[] = {"B/s", "KiB/s", "MiB/s", "GiB/s", "TiB/s"};
    int u = 0;
    double v = bytes / std::max(seconds, 1e-9);
    while (v >= 1024.0 && u < 4) v /= 1024.0, ++u;
    return strprintf("%.2f %s", v, units[u]);
}

// the value `x` in the element type `dtype`, as raw bytes (small integers are exact in every type the tool offers)
static void fill(std::vector<unsigned char>& buf, size_t count, int dtype, int x) {
    const size_t es = dtype_size(dtype);
    buf.resize(count * es);
    unsigned char one[8] = {0};
    if (dtype == kF32) {
        const float f = (float)x;
        std::memcpy(one, &f, 4);
    } else if (dtype == kI32) {
        const int32_t i = x;
        std::memcpy(one, &i, 4);
    } else {
        const float f = (float)x;
        const uint16_t h = dtype == kF16 ? float_to_half(f) : float_to_bf16(f);
        std::memcpy(one, &h, 2);
    }
    for (size_t i = 0; i < count; ++i) std::memcpy(buf.data() + i * es, one, es);
}

int main(int argc, char** argv) {
    kftools::Flags fl{argc, argv};
    const std::string model = fl.str("model", "resnet50"), mode = fl.str("mode", "seq"), dts = fl.str("dtype", "f32");
    const bool fuse = fl.flag("fuse");
    const int epochs = (int)fl.num("epochs", 15), warmup = (int)fl.num("warmup", 2);
    const double p_fail = std::atof(fl.str("rand-failure", "0").c_str());
    auto sizes = kftools::fake_model_sizes(model);
    if (sizes.empty()) {
        std::fprintf(stderr, "invalid model name: %s (", model.c_str());
        for (auto& n : kftools::fake_model_names()) std::fprintf(stderr, " %s", n.c_str());
        std::fprintf(stderr, " )\n");
        return 2;
    }
    if (mode != "seq" && mode != "par") {
        std::fprintf(stderr, "invalid mode: %s (seq | par)\n", mode.c_str());
        return 2;
    }
    int dtype = -1;
    if (dts == "f32") dtype = kF32;
    if (dts == "f16") dtype = kF16;
    if (dts == "bf16") dtype = kBF16;
    if (dts == "i32") dtype = kI32;
    if (dtype < 0) {
        std::fprintf(stderr, "invalid dtype: %s (f32 | f16 | bf16 | i32)\n", dts.c_str());
        return 2;
    }
    if (p_fail > 0) {
        std::mt19937_64 rng(std::random_device{}() ^ ((uint64_t)getpid() << 20));
        if (std::uniform_real_distribution<double>(0, 1)(rng) < p_fail) {
            // the first line on stderr is the loader's message, which is what the launcher's start-up retry looks for
            std::fprintf(stderr, "Inconsistency detected by ld.so: injected start-up failure (kungfu-bench-allreduce -rand-failure %.2f)\n", p_fail);
            return 1;
        }
    }
    if (fuse) sizes = {kftools::total_elements(sizes)};
    int rc = 0;
    try {
        Peer peer(PeerConfig::from_env());
        peer.start();
        auto sess = peer.session();
        const int np = sess->size(), rank = sess->rank();
        const size_t total = kftools::total_elements(sizes), es = dtype_size(dtype);
        if (rank == 0)
            std::printf("model: %s, %zu tensors, %zu parameters, %.2f MiB (%s), mode: %s, np: %d\n", model.c_str(), sizes.size(), total,
                        total * (double)es / (1 << 20), dts.c_str(), mode.c_str(), np);
        // send / receive buffers of all tensors are views of two flat allocations
        std::vector<unsigned char> send, recv(total * es, 0), want_first;
        fill(send, total, dtype, rank + 1);
        // np (np + 1) / 2 is exact in bf16 up to np = 22 (253 needs 8 bits); beyond that the check compares against the type's own sum
        fill(want_first, 1, dtype, np * (np + 1) / 2);
        std::vector<size_t> offs(sizes.size());
        for (size_t i = 0, o = 0; i < sizes.size(); ++i) offs[i] = o, o += sizes[i];
        auto reduce_one = [&](size_t i) {
            Workspace w{send.data() + offs[i] * es, recv.data() + offs[i] * es, sizes[i], dtype, kSum, "bench-ar-" + std::to_string(i)};
            sess->all_reduce(w);
        };
        auto run_epoch = [&] {
            if (mode == "seq") {
                for (size_t i = 0; i < sizes.size(); ++i) reduce_one(i);
            } else {
                // every tensor in flight at once: the session matches them by name, a crew of threads drains an index counter
                std::atomic<size_t> next{0};
                const int crew = (int)std::min<size_t>(sizes.size(), 32);
                std::vector<std::thread> ts;
                for (int t = 0; t < crew; ++t)
                    ts.emplace_back([&] {
                        for (size_t i; (i = next.fetch_add(1)) < sizes.size();) reduce_one(i);
                    });
                for (auto& t : ts) t.join();
            }
        };
        const bool exact = dtype == kF32 || dtype == kI32 || np <= 22;
        auto verify = [&]() -> bool {
            if (!exact) return true;
            for (size_t i = 0; i < sizes.size(); ++i) {
                const unsigned char* a = recv.data() + offs[i] * es;
                const unsigned char* b = recv.data() + (offs[i] + sizes[i] - 1) * es;
                if (std::memcmp(a, want_first.data(), es) || std::memcmp(b, want_first.data(), es)) return false;
            }
            return true;
        };
        sess->barrier();
        bool good = true;
        for (int e = 0; e < warmup; ++e) {
            std::fill(recv.begin(), recv.end(), (unsigned char)0);
            run_epoch();
            good = good && verify();
        }
        sess->barrier();
        const double t0 = now_s();
        for (int e = 0; e < epochs; ++e) run_epoch();
        const double dt = now_s() - t0;
        good = good && verify();
        const double workload = (double)epochs * 4.0 * (np - 1) * total * es; // the reference's accounting
        int64_t bad = good ? 0 : 1, bad_all = 0;
        Workspace w{&bad, &bad_all, 1, kI64, kSum, "bench-ar-verdict"};
        sess->all_reduce(w);
        if (rank == 0) {
            std::printf("took %.3fs, rate: %s\n", dt, show_rate(workload, dt).c_str());
            std::printf("Result: model: %s, %zu tensors, mode: %s%s, np: %d, rate: %s, %.3f ms per all-reduced model%s\n", model.c_str(),
                        sizes.size(), mode.c_str(), fuse ? " (fused)" : "", np, show_rate(workload, dt).c_str(), dt / std::max(epochs, 1) * 1e3,
                        bad_all ? "  [VERIFICATION FAILED]" : "");
        }
        if (bad_all) rc = 3;
        sess->barrier();
        peer.close();
    } catch (const std::exception& e) {
        std::fprintf(stderr, "kungfu-bench-allreduce: %s\n", e.what());
        return 1;
    }
    return rc;
}
