/* A worker written in C against the runtime's C ABI — no C++, no Python.
 *
 *   gcc -std=c99 -D_DEFAULT_SOURCE -I kungfu_b200/csrc/runtime examples/c/allreduce.c -L kungfu_b200/lib -lkungfu-comm -Wl,-rpath,$PWD/kungfu_b200/lib -o allreduce
 *   kungfu-run -np 4 ./allreduce
 *
 * Reference: the cgo exports of libkungfu-comm (srcs/go/libkungfu-comm/{main,collective}.go) as used by the C++ fake trainers
 * (tests/cpp/integration/fake_kungfu_trainer.cpp). Synchronous calls pass done = NULL; the asynchronous form returns at once and
 * invokes the callback from a runtime thread. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include "capi.h"

static volatile int completed = 0;
static void on_done(void* arg) {
    (void)arg;
    __sync_fetch_and_add(&completed, 1);
}

#define CHECK(x)                                                                  \
    do {                                                                          \
        if (!(x)) {                                                               \
            fprintf(stderr, "%s:%d: %s failed: %s\n", __FILE__, __LINE__, #x, kungfu_last_error()); \
            exit(3);                                                              \
        }                                                                         \
    } while (0)

int main(void) {
    CHECK(kungfu_init() == 0);
    const int np = kungfu_size(), rank = kungfu_rank();
    enum { N = 1 << 16 };
    float* x = malloc(N * sizeof(float));
    float* y = malloc(N * sizeof(float));
    for (int i = 0; i < N; ++i) x[i] = (float)(rank + 1);

    CHECK(kungfu_all_reduce(x, y, N, KungFu_F32, KungFu_SUM, "c-sum", NULL, NULL) == 0);
    CHECK(y[0] == np * (np + 1) / 2.0f && y[N - 1] == y[0]);

    /* four tensors in flight at once through the callback form */
    const char* names[4] = {"c-a", "c-b", "c-c", "c-d"};
    for (int k = 0; k < 4; ++k) CHECK(kungfu_all_reduce(x + k * 1000, y + k * 1000, 1000, KungFu_F32, KungFu_MAX, names[k], on_done, NULL) == 0);
    for (int i = 0; i < 5000 && completed < 4; ++i) usleep(1000);
    CHECK(completed == 4 && y[0] == (float)np && y[3999] == (float)np);

    int agreed = 0;
    CHECK(kungfu_consensus("same", 4, "c-consensus", &agreed) == 0 && agreed == 1);
    int64_t step = 10 * rank;
    CHECK(kungfu_all_reduce_int_max(&step) == 0 && step == 10 * (np - 1));
    CHECK(kungfu_barrier() == 0);
    if (rank == 0) printf("c worker ok np=%d (cluster version %d, %d host(s))\n", np, kungfu_cluster_version(), kungfu_host_count());
    free(x);
    free(y);
    return kungfu_finalize();
}
