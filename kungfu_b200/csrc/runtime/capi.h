/* kungfu_b200 runtime — C ABI for C / C++ programs (tests, tools, foreign bindings).
 *
 * Reference: the cgo exports of libkungfu-comm (`GoKungfuInit … GoKungfuAllReduce … GoKungfuResizeCluster`,
 * srcs/go/libkungfu-comm/{main,collective,adapt,monitoring}.go) and the python C API
 * (srcs/cpp/include/kungfu/python/c_api.h:8-54). Async contract as in the reference: `done == NULL` runs synchronously,
 * otherwise the call returns at once and `done(arg)` is invoked from a runtime thread on completion.
 * Every function returns 0 on success and non-zero on error (`kungfu_last_error()` has the text). */
#pragma once
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void (*kungfu_callback_t)(void* arg);

enum { KungFu_F32 = 0, KungFu_BF16 = 1, KungFu_F16 = 2, KungFu_I32 = 3, KungFu_I64 = 4, KungFu_U8 = 5, KungFu_F64 = 6, KungFu_I8 = 7,
       KungFu_U32 = 8, KungFu_U64 = 9, KungFu_I16 = 10, KungFu_U16 = 11 };
enum { KungFu_SUM = 0, KungFu_MIN = 1, KungFu_MAX = 2, KungFu_PROD = 3 };

int kungfu_init(void);                                   /* from the environment (kungfu-run / torchrun / single) */
int kungfu_init_single_machine(int rank, int size, int base_port);
int kungfu_finalize(void);
const char* kungfu_last_error(void);

int kungfu_rank(void);
int kungfu_size(void);
int kungfu_local_rank(void);
int kungfu_local_size(void);
int kungfu_host_count(void);
uint64_t kungfu_uid(void);
int kungfu_detached(void);
int kungfu_cluster_version(void);

int kungfu_barrier(void);
int kungfu_consensus(const void* buf, size_t len, const char* name, int* agreed);
int kungfu_all_reduce(const void* send, void* recv, size_t count, int dtype, int op, const char* name, kungfu_callback_t done, void* arg);
int kungfu_monitored_all_reduce(const void* send, void* recv, size_t count, int dtype, int op, const int32_t* tree, int tree_len, const char* name);
int kungfu_subset_all_reduce(const void* send, void* recv, size_t count, int dtype, int op, const int32_t* forest, int forest_len, const char* name);
int kungfu_noop(kungfu_callback_t done, void* arg);
int kungfu_cross_all_reduce(const void* send, void* recv, size_t count, int dtype, int op, const char* name);
int kungfu_reduce(const void* send, void* recv, size_t count, int dtype, int op, const char* name);
int kungfu_broadcast(const void* send, void* recv, size_t count, int dtype, const char* name, kungfu_callback_t done, void* arg);
int kungfu_local_reduce(const void* send, void* recv, size_t count, int dtype, int op, const char* name);
int kungfu_local_broadcast(const void* send, void* recv, size_t count, int dtype, const char* name);
int kungfu_gather(const void* send, void* recv, size_t count, int dtype, const char* name);
int kungfu_all_gather(const void* send, void* recv, size_t count, int dtype, const char* name);
int kungfu_all_reduce_int_max(int64_t* x);

int kungfu_save(const char* name, const void* buf, size_t len);
int kungfu_save_version(const char* version, const char* name, const void* buf, size_t len);
int kungfu_request(int rank, const char* version, const char* name, void* buf, size_t len, int* ok);

int kungfu_resize(int new_size, int* changed, int* detached);
int kungfu_resize_from_url(int* changed, int* detached);
int kungfu_propose_new_size(int new_size);
/* shrink-and-continue (kungfu-run -w -shrink-on-failure): after a collective failed because a worker died, adopt the cluster the runner
 * announced (waits at most timeout_s for the announcement); collectives fail fast while *pending is set and recover has not run */
int kungfu_recover(double timeout_s, int* changed, int* detached);
int kungfu_failure_pending(int* pending);
int kungfu_set_tree(const int32_t* father, int n);
int kungfu_get_peer_latencies(float* out, int n);
int kungfu_get_egress_rates(float* out, int n);
int kungfu_check_interference(int* result);
int kungfu_calc_stats(void);
int kungfu_log_stats(void);
int kungfu_print_strategy_stats(void);

#ifdef __cplusplus
}
#endif
