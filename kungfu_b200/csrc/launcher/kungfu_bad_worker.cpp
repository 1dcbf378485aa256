// kungfu-bad-worker — a job in which one worker fails: the launcher must cancel the others and exit non-zero (fail-fast), or, with
// `-w -shrink-on-failure`, carry on without it.
//
//   kungfu-run -np 4 kungfu-bad-worker [-rank 0] [-after 5s] [-code 1] [-hang]
//
// Reference: tests/go/cmd/kungfu-bad-worker/kungfu-bad-worker.go (rank 0 exits 1 after 5 s while the others sleep). Here the rank,
// the delay and the exit code are flags, and `-hang` makes the bad worker stop answering instead of exiting (what `-auto-recover`
// with a hang watch is for); the good workers keep all-reducing so that the failure hits them inside a collective.
#include <cstdio>
#include <cstdlib>
#include <string>
#include <unistd.h>

#include "base.hpp"
#include "fake_model.hpp"
#include "peer.hpp"

using namespace kfrt;

int main(int argc, char** argv) {
    kftools::Flags fl{argc, argv};
    const int bad_rank = (int)fl.num("rank", 0), code = (int)fl.num("code", 1);
    const double after = parse_duration_s(fl.str("after", "5s"), 5.0), total = parse_duration_s(fl.str("total", "20s"), 20.0);
    const bool hang = fl.flag("hang");
    try {
        Peer peer(PeerConfig::from_env());
        peer.start();
        auto sess = peer.session();
        const int rank = sess->rank();
        const double t0 = now_s();
        for (int step = 0; now_s() - t0 < total; ++step) {
            if (rank == bad_rank && now_s() - t0 >= after) {
                std::fprintf(stderr, "bad worker (rank %d): %s after %.1f s\n", rank, hang ? "hanging" : "exiting", now_s() - t0);
                if (hang)
                    for (;;) pause();
                _exit(code);
            }
            float one = 1.f, sum = 0.f;
            Workspace w{&one, &sum, 1, kF32, kSum, "bad-worker-" + std::to_string(step)};
            sess->all_reduce(w);
            sleep_ms(50);
        }
        peer.close();
    } catch (const std::exception& e) {
        std::fprintf(stderr, "kungfu-bad-worker (a good one): %s\n", e.what());
        return 1;
    }
    return 0;
}
