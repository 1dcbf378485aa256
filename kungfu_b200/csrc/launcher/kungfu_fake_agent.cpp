// kungfu-fake-agent — C++ integration program over the C ABI, run under `kungfu-run -np N -strategy S`.
//
// Reference: tests/cpp/integration/fake_agent.cpp:15-205 (AllReduce correctness y[i] = i*np, Reduce / AllReduce
// micro-bench loops, Gather, AllGatherTransform + distributed MST) and test_p2p_apis.cpp:7-38 (versioned Save -> Barrier
// -> Request). Also the model-shaped all-reduce micro-benchmark of `kungfu-bench-allreduce`
// (tests/go/cmd/kungfu-bench-allreduce/kungfu-bench-allreduce.go:73-110): prints
// `RESULT: <rate> GiB/s` with the reference's "equivalent data rate" 4(np-1)*S/t.
#include <algorithm>
#include <atomic>
#include <unistd.h>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "capi.h"
#include "fake_model.hpp"
#include "graph.hpp"

#define CHECK(cond)                                                                   \
    do {                                                                              \
        if (!(cond)) {                                                                \
            fprintf(stderr, "CHECK FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond);   \
            exit(3);                                                                  \
        }                                                                             \
    } while (0)

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

static std::vector<size_t> model_sizes(const std::string& name) {
    auto v = kftools::fake_model_sizes(name);
    return v.empty() ? kftools::fake_model_sizes("resnet50") : v;
}

int main(int argc, char** argv) {
    std::string mode = argc > 1 ? argv[1] : "test";
    std::string model = argc > 2 ? argv[2] : "resnet50";
    CHECK(kungfu_init() == 0);
    const int np = kungfu_size(), rank = kungfu_rank();
    if (mode == "test") {
        // all-reduce correctness: y[i] = i * np
        for (size_t n : {1ul, 7ul, 1024ul, (size_t)3 << 18}) {
            std::vector<int32_t> x(n), y(n, -1);
            for (size_t i = 0; i < n; ++i) x[i] = (int32_t)(i % 1000);
            CHECK(kungfu_all_reduce(x.data(), y.data(), n, KungFu_I32, KungFu_SUM, ("ar" + std::to_string(n)).c_str(), nullptr, nullptr) == 0);
            for (size_t i = 0; i < n; ++i) CHECK(y[i] == (int32_t)(i % 1000) * np);
        }
        // gather + all_gather
        std::vector<float> mine(3, (float)rank), all(3 * np, -1.f);
        CHECK(kungfu_all_gather(mine.data(), all.data(), 3, KungFu_F32, "ag") == 0);
        for (int r = 0; r < np; ++r) CHECK(all[3 * r] == (float)r);
        std::vector<float> g(rank == 0 ? 3 * np : 0);
        CHECK(kungfu_gather(mine.data(), g.data(), 3, KungFu_F32, "g") == 0);
        if (rank == 0)
            for (int r = 0; r < np; ++r) CHECK(g[3 * r + 2] == (float)r);
        // AllGatherTransform: distributed MST over latencies (gather -> Prim -> same result everywhere)
        std::vector<float> lat(np);
        CHECK(kungfu_get_peer_latencies(lat.data(), np) == 0);
        std::vector<float> allw((size_t)np * np);
        CHECK(kungfu_all_gather(lat.data(), allw.data(), np, KungFu_F32, "mstw") == 0);
        auto edges = kfrt::minimum_spanning_tree(std::vector<double>(allw.begin(), allw.end()), np);
        CHECK((int)edges.size() == np - 1);
        // p2p: versioned save -> barrier -> request; unknown blob must fail softly
        std::vector<int32_t> blob(256, rank + 1), got(256, 0);
        CHECK(kungfu_save_version("v1", "blob", blob.data(), blob.size() * 4) == 0);
        CHECK(kungfu_barrier() == 0);
        int ok = 0;
        const int tgt = (rank + 1) % np;
        CHECK(kungfu_request(tgt, "v1", "blob", got.data(), got.size() * 4, &ok) == 0 && ok == 1 && got[0] == tgt + 1);
        CHECK(kungfu_request(tgt, "v9", "blob", got.data(), got.size() * 4, &ok) == 0 && ok == 0);
        int agreed = 0;
        CHECK(kungfu_consensus("abc", 3, "c", &agreed) == 0 && agreed == 1);
        CHECK(kungfu_barrier() == 0);
        { // subset all-reduce: pairs {0,1} {2,3} ...; a trailing singleton keeps its value
            std::vector<int32_t> forest(np);
            for (int i = 0; i < np; ++i) forest[i] = (2 * (i / 2) + 1 < np) ? 2 * (i / 2) : i;
            std::vector<float> a(1000, float(rank + 1)), b(1000, 0.f);
            CHECK(kungfu_subset_all_reduce(a.data(), b.data(), a.size(), KungFu_F32, KungFu_SUM, forest.data(), np, "subset") == 0);
            float want = 0;
            for (int i = 0; i < np; ++i)
                if (forest[i] == forest[rank]) want += float(i + 1);
            CHECK(b[0] == want && b[999] == want);
        }
        { // noop with a completion callback: runs on the runtime's executor
            static std::atomic<int> fired{0};
            CHECK(kungfu_noop([](void*) { fired.fetch_add(1); }, nullptr) == 0);
            for (int i = 0; i < 2000 && fired.load() == 0; ++i) usleep(1000);
            CHECK(fired.load() == 1);
            CHECK(kungfu_noop(nullptr, nullptr) == 0);
        }
        CHECK(kungfu_barrier() == 0);
        if (rank == 0) printf("fake-agent ok np=%d\n", np);
    } else if (mode == "shrink") {
        // shrink-and-continue through the C ABI (under `kungfu-run -w -shrink-on-failure`): `argv[2]` = "<rank>@<step>", that worker
        // dies without a word inside that step; the others see their all-reduce fail, call kungfu_recover and finish 20 steps. The value
        // every step reduces is 1 per worker, so the result IS the cluster size of that step.
        int die_rank = -1, die_step = -1;
        if (argc > 2) sscanf(argv[2], "%d@%d", &die_rank, &die_step);
        int recovered = 0, size_seen_min = np, size_seen_max = np;
        for (int step = 0; step < 20; ++step) {
            if (recovered == 0 && kungfu_rank() == die_rank && step == die_step) _exit(17);
            float one = 1.f, sum = 0.f;
            const std::string name = "s" + std::to_string(step) + "v" + std::to_string(kungfu_cluster_version());
            if (kungfu_all_reduce(&one, &sum, 1, KungFu_F32, KungFu_SUM, name.c_str(), nullptr, nullptr) != 0) {
                int changed = 0, detached = 0;
                const double t0 = now();
                CHECK(kungfu_recover(20.0, &changed, &detached) == 0);
                CHECK(changed == 1 && detached == 0);
                printf("fake-agent rank %d: recover took %.1f ms\n", kungfu_rank(), (now() - t0) * 1e3);
                ++recovered;
                --step; // repeat the interrupted step with the smaller cluster
                continue;
            }
            CHECK((int)sum == kungfu_size());
            size_seen_min = std::min(size_seen_min, (int)sum);
            size_seen_max = std::max(size_seen_max, (int)sum);
            usleep(20000);
        }
        int pending = -1;
        CHECK(kungfu_failure_pending(&pending) == 0 && pending == 0);
        CHECK(kungfu_barrier() == 0);
        printf("fake-agent shrink ok rank=%d recovered=%d sizes=%d..%d final=%d\n", kungfu_rank(), recovered, size_seen_min, size_seen_max, kungfu_size());
    } else {
        // model-shaped all-reduce micro-benchmark (fused or per tensor)
        const bool fuse = mode == "bench-fused";
        const bool async = mode == "bench-async"; // every tensor in flight at once through the callback form of the C ABI
        auto sizes = model_sizes(model);
        size_t total = 0;
        for (auto s : sizes) total += s;
        std::vector<float> buf(total, 1.0f), out(total);
        const int warm = 2, iters = 8;
        std::vector<double> rates;
        double img_s = 0;
        for (int it = 0; it < warm + iters; ++it) {
            kungfu_barrier();
            const double t0 = now();
            if (fuse) {
                CHECK(kungfu_all_reduce(buf.data(), out.data(), total, KungFu_F32, KungFu_SUM, ("fused" + std::to_string(it)).c_str(), nullptr, nullptr) == 0);
            } else if (async) {
                static std::atomic<int> pending{0};
                pending = (int)sizes.size();
                size_t off = 0;
                for (size_t i = 0; i < sizes.size(); ++i) {
                    CHECK(kungfu_all_reduce(buf.data() + off, out.data() + off, sizes[i], KungFu_F32, KungFu_SUM,
                                            ("a" + std::to_string(i) + "_" + std::to_string(it)).c_str(), [](void*) { pending.fetch_sub(1); }, nullptr) == 0);
                    off += sizes[i];
                }
                while (pending.load() > 0) usleep(50);
            } else {
                size_t off = 0;
                for (size_t i = 0; i < sizes.size(); ++i) {
                    CHECK(kungfu_all_reduce(buf.data() + off, out.data() + off, sizes[i], KungFu_F32, KungFu_SUM,
                                            ("t" + std::to_string(i) + "_" + std::to_string(it)).c_str(), nullptr, nullptr) == 0);
                    off += sizes[i];
                }
            }
            const double dt = now() - t0;
            if (it >= warm) {
                rates.push_back(4.0 * (np - 1) * total * 4 / dt / (1 << 30));
                img_s += 32.0 / dt / iters; // the reference's fake trainer: one step of batch 32 per model all-reduce
            }
        }
        double mean = 0, var = 0;
        for (double r : rates) mean += r / rates.size();
        for (double r : rates) var += (r - mean) * (r - mean) / rates.size();
        if (rank == 0)
            printf("RESULT: %.3f +-%.3f {\"model\": \"%s\", \"np\": %d, \"fuse\": %s, \"unit\": \"GiB/s equivalent data rate\", \"tensors\": %zu, \"bytes\": %zu}\n",
                   mean, 1.96 * std::sqrt(var), model.c_str(), np, fuse ? "true" : "false", sizes.size(), total * 4);
        if (rank == 0) printf("Img/sec %.1f per worker (communication only, batch 32 per step; reference: fake_trainer.hpp:169-176)\n", img_s);
    }
    kungfu_finalize();
    return 0;
}
