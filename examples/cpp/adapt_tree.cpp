// A C++ application on the public header `kungfu.hpp`: measure peer latencies, agree on a minimum spanning tree of them, make it the
// all-reduce topology, and keep training (here: keep all-reducing) on the new tree.
//
//   g++ -std=c++17 -I kungfu_b200/csrc/runtime examples/cpp/adapt_tree.cpp -L kungfu_b200/lib -lkungfu-comm -Wl,-rpath,$PWD/kungfu_b200/lib -pthread -o adapt_tree
//   kungfu-run -np 4 ./adapt_tree
//
// Reference: the distributed-MST part of tests/cpp/integration/fake_agent.cpp:150-205 (GetPeerLatencies -> AllGatherTransform with
// MinimumSpanningTree -> SetTree) and experimental/adapt_strategy/adapt_strategy.py, which does the same from Python.
#include <atomic>
#include <chrono>
#include <cstdio>
#include <thread>
#include <vector>

#include "kungfu.hpp"

#define CHECK(x)                                                                                       \
    do {                                                                                               \
        if (!(x)) {                                                                                    \
            std::fprintf(stderr, "%s:%d: %s: %s\n", __FILE__, __LINE__, #x, kungfu::last_error().c_str()); \
            return 3;                                                                                  \
        }                                                                                              \
    } while (0)

int main() {
    kungfu::Peer peer;
    CHECK(peer.Ok());
    const int np = peer.Size(), rank = peer.Rank();

    std::vector<float> grad(1 << 18, 1.0f), sum(grad.size());
    CHECK(peer.AllReduce(grad.data(), sum.data(), (int)grad.size(), KungFu_SUM, "before") == 0);
    CHECK(sum.front() == (float)np && sum.back() == (float)np);

    // every worker measures its row of the latency matrix; rank 0 turns the matrix into a tree; everybody gets the same tree
    std::vector<float> row(np);
    CHECK(peer.GetPeerLatencies(row.data(), np) == 0 && row[rank] == 0.0f);
    std::vector<int32_t> tree(np, -1);
    CHECK(peer.AllGatherTransform(row.data(), np, tree.data(), np, "latency-mst", [np](const float* all, int, int32_t* out, int) {
        const auto father = kungfu::minimum_spanning_tree(np, all);
        for (int i = 0; i < np; ++i) out[i] = father[i];
    }) == 0);
    int roots = 0;
    for (int i = 0; i < np; ++i) {
        CHECK(tree[i] >= 0 && tree[i] < np);
        roots += tree[i] == i;
    }
    CHECK(roots == 1 && tree[0] == 0);
    bool same = false;
    CHECK(peer.Consensus(tree.data(), np, KungFu_I32, &same, "tree-agreed") == 0 && same);
    CHECK(peer.SetTree(tree.data()) == 0);

    // the new topology carries synchronous and asynchronous all-reduces alike
    CHECK(peer.AllReduce(grad.data(), sum.data(), (int)grad.size(), KungFu_SUM, "after") == 0);
    CHECK(sum.front() == (float)np && sum.back() == (float)np);
    std::atomic<int> done{0};
    std::vector<int64_t> step(4, rank), top(4, -1);
    for (int k = 0; k < 4; ++k)
        CHECK(peer.AllReduce(&step[k], &top[k], 1, KungFu_MAX, ("async" + std::to_string(k)).c_str(), [&done] { done.fetch_add(1); }) == 0);
    for (int i = 0; i < 5000 && done.load() < 4; ++i) std::this_thread::sleep_for(std::chrono::milliseconds(1));
    CHECK(done.load() == 4 && top[0] == np - 1 && top[3] == np - 1);

    // blob store: a missing blob is reported through `ok`, not as an error
    CHECK(peer.Save("w", grad.data(), 16, KungFu_F32) == 0 && peer.Barrier() == 0);
    bool ok = false;
    std::vector<float> got(16, 0.f);
    CHECK(peer.Request((rank + 1) % np, "w", got.data(), 16, KungFu_F32, &ok) == 0 && ok && got[15] == 1.0f);
    CHECK(peer.Request((rank + 1) % np, "absent", got.data(), 16, KungFu_F32, &ok) == 0 && !ok);
    CHECK(peer.Barrier() == 0);
    if (rank == 0) {
        std::printf("adapt_tree ok np=%d tree=[", np);
        for (int i = 0; i < np; ++i) std::printf("%s%d", i ? " " : "", tree[i]);
        std::printf("]\n");
    }
    return 0;
}
