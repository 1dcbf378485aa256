// kungfu_b200 — GradEngine: the native GPU op engine behind the distributed optimizers.
//
// What the reference keeps native for its NCCL path — the ordered scheduler that releases per-tensor NCCL calls in a
// cluster-agreed order on one dedicated thread (srcs/cpp/src/nccl/scheduler.cpp:8-130), its private CUDA stream with a
// blocking sync per op (srcs/cpp/src/cuda/stream.cpp:8-61, nccl/gpu_collective.cpp:146-155) and the HandleManager
// (include/kungfu/utils/handler_manager.hpp:40-83) — is re-imagined here for stream-ordered execution:
//
//   * the bucket plan (fixed at optimizer construction, consensus-checked once) lives in this object;
//   * `grad_ready(set, entry, grad_ptr, stream)` is the ONE call a post-accumulate-grad hook makes: it records the
//     gradient hand-over (a stolen autograd tensor is queued for the bucket's single multi-tensor copy), counts the bucket
//     and, strictly in bucket order, launches gate -> fused body -> wait on the high-priority comm stream, ordered after
//     the autograd stream by an event — no thread, no host-side waiting, CUDA-graph capturable;
//   * `finish(stream)` joins the comm stream back; completion handles are events.
//
// Python keeps only what needs Python objects: re-pointing `param.grad`, the NCCL baseline arm and elastic re-sharding.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../kernels/kernels.h"
#include "../mesh/peer_mesh.h"

namespace kf {

// kModeCopy: hand-over only — gradients stolen from autograd are moved into the arena bucket by bucket, nothing is reduced (the model-
// averaging optimizers consume the arena in their own kernel at the end of the step)
enum EngineMode : int { kModeSsgd = 0, kModeGrads = 1, kModeCopy = 2 };

struct EngineBucket {
    size_t start_bytes = 0; // offset of the bucket inside the set's flat region
    size_t nbytes = 0;      // padded size
    size_t numel = 0;
    int n_entries = 0;
    float* master = nullptr; // shard (ssgd over peers) or full (single rank) fp32 master, may be null
    float* s1 = nullptr;
    float* s2 = nullptr;
    // per-step
    int ready = 0;
    std::vector<const void*> src;
    std::vector<void*> dst;
    std::vector<size_t> bytes;
};

struct EngineEntry {
    size_t offset_bytes = 0; // inside the set's flat region
    size_t nbytes = 0;
    int bucket = 0;
    bool got = false;
};

struct EngineSet {
    size_t g_off = 0, p_off = 0; // arena offsets of the gradient / parameter regions
    int dtype = 0;
    float* hyper = nullptr;
    int opt_kind = 0;
    std::vector<EngineBucket> buckets;
    std::vector<EngineEntry> entries;
    int launched = 0;
};

class GradEngine {
public:
    GradEngine(PeerMesh* mesh, int mode, bool average, bool gated, int max_blocks, int threads, int algo);
    ~GradEngine();
    GradEngine(const GradEngine&) = delete;
    GradEngine& operator=(const GradEngine&) = delete;

    int add_set(size_t g_off, size_t p_off, int dtype, uintptr_t hyper, int opt_kind);
    int add_bucket(int set, size_t start_bytes, size_t nbytes, size_t numel);
    int add_entry(int set, int bucket, size_t offset_bytes, size_t nbytes);
    void set_bucket_state(int set, int bucket, uintptr_t master, uintptr_t s1, uintptr_t s2);
    void set_monitor(uintptr_t sumsq) { sumsq_ = reinterpret_cast<float*>(sumsq); }
    void set_comm_stream(uintptr_t s) { comm_ = reinterpret_cast<cudaStream_t>(s); }

    // first call of a step: orders the comm stream after `stream` and ticks the device-side hyper-parameter blocks
    void open_step(cudaStream_t stream);
    // returns true if the gradient was a foreign tensor that has been queued for the bucket copy (caller re-points p.grad)
    bool grad_ready(int set, int entry, const void* grad_ptr, cudaStream_t stream);
    // launch whatever the hooks have not launched (entries without a gradient contribute zeros)
    void flush(cudaStream_t stream, bool zero_missing);
    // current stream waits for the comm stream; per-step bookkeeping is reset. Returns the device status word.
    uint32_t finish(cudaStream_t stream);

    uint64_t launches() const { return launches_; }
    int nranks() const { return mesh_->nranks(); }
    void set_deferred(bool d) { deferred_ = d; }

private:
    void launch_bucket(EngineSet& s, int bi, cudaStream_t stream, bool after_stream);
    int pick_blocks(size_t nbytes) const;
    int pick_algo(int dtype) const;
    cudaEvent_t next_event();

    PeerMesh* mesh_;
    int mode_;
    bool average_, gated_;
    int max_blocks_, threads_, algo_;
    cudaStream_t comm_ = nullptr;
    bool own_stream_ = false;
    std::vector<EngineSet> sets_;
    std::vector<cudaEvent_t> events_;
    size_t ev_next_ = 0;
    bool step_open_ = false;
    bool deferred_ = false;
    float* sumsq_ = nullptr;
    uint64_t launches_ = 0;
};

} // namespace kf
