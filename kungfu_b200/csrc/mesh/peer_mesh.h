// kungfu_b200 — PeerMesh: the data plane of the runtime.
//
// Each rank owns one symmetric device allocation (arena + signal pads) created with the CUDA VMM API
// and exported as a POSIX file descriptor; descriptors travel between the per-GPU processes over
// SCM_RIGHTS on abstract Unix datagram sockets. Every rank maps every peer's allocation, and — when
// the device reports multicast support — binds the arena into an NVLS multicast object so kernels
// can use `multimem.*`. A cudaIpc path is kept as a fallback for drivers without fd export.
//
// Replaces: NCCL communicator lifecycle (ncclGetUniqueId / ncclCommInitRank / re-init on resize,
// srcs/cpp/src/nccl/gpu_collective.cpp:101-111,169-191; srcs/cpp/src/python/init_nccl.cpp:5-12) and
// CudaStream staging copies (srcs/cpp/src/cuda/stream.cpp:30-36). `connect()` after a membership
// change is the counterpart of `ResetNcclHelper` (srcs/cpp/src/tensorflow/ops/gpu/scheduler.cpp:55-69):
// local memory (and every tensor viewing it) survives, only the peer mappings are rebuilt.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <functional>
#include <string>
#include <vector>

#include "../kernels/mesh_args.h"

namespace kf {

// out-of-band all-gather of small byte strings supplied by the control plane
using OobAllGather = std::function<std::vector<std::string>(const std::string&)>;

struct MeshTopology {
    int device = -1;
    int device_count = 0;
    std::vector<int> can_access;   // per visible device: cudaDeviceCanAccessPeer
    std::vector<int> perf_rank;    // cudaDevP2PAttrPerformanceRank
    std::vector<int> native_atomic;
    int multicast_supported = 0;
    int sm_count = 0;
    size_t total_mem = 0;
    std::string name;
};

class PeerMesh {
public:
    // mode: "auto" | "vmm" | "ipc" | "local"
    PeerMesh(int device, size_t arena_bytes, const std::string& mode);
    ~PeerMesh();
    PeerMesh(const PeerMesh&) = delete;
    PeerMesh& operator=(const PeerMesh&) = delete;

    // (Re)build peer mappings for the given membership. Collective over all ranks of `oob`.
    void connect(int rank, int nranks, const OobAllGather& oob, bool want_multicast);
    void disconnect();

    MeshArgs args(int channel) const;
    MeshTopology topology() const;

    int rank() const { return rank_; }
    int nranks() const { return nranks_; }
    int device() const { return device_; }
    int epoch() const { return epoch_; }
    size_t arena_bytes() const { return arena_bytes_; }
    size_t alloc_bytes() const { return alloc_bytes_; }
    uintptr_t local_ptr() const { return (uintptr_t)local_va_; }
    uintptr_t peer_ptr(int r) const { return (uintptr_t)peer_va_.at(r); }
    uintptr_t mc_ptr() const { return (uintptr_t)mc_va_; }
    bool has_multicast() const { return mc_va_ != nullptr; }
    const std::string& mode() const { return mode_; }
    const std::string& multicast_note() const { return mc_note_; }

    std::vector<uint32_t> status() const;
    void clear_status();
    // host -> device: make every device-side wait of this mesh give up at its next look at the status word (a peer is known to
    // be dead: do not sit out the time-out). Kernels in flight finish with whatever they have; the caller discards the step.
    void abort();
    void set_timeout_ms(double ms) { timeout_ns_ = (unsigned long long)(ms * 1e6); }
    void reset_signals(cudaStream_t st);

private:
    void alloc_local();
    void free_local();
    void connect_vmm(const OobAllGather& oob, bool want_multicast);
    void connect_ipc(const OobAllGather& oob);

    int device_ = 0;
    std::string mode_;
    size_t arena_bytes_ = 0; // usable arena
    size_t alloc_bytes_ = 0; // arena + signal pads, rounded to granularity
    size_t signal_off_ = 0;
    int rank_ = 0, nranks_ = 1, epoch_ = 0;
    bool connected_ = false;

    void* local_va_ = nullptr;
    unsigned long long local_handle_ = 0; // CUmemGenericAllocationHandle (vmm mode)
    bool local_is_vmm_ = false;
    std::vector<void*> peer_va_;
    std::vector<unsigned long long> peer_handle_;
    std::vector<bool> peer_is_ipc_;
    void* mc_va_ = nullptr;
    unsigned long long mc_handle_ = 0;
    bool mc_bound_ = false;
    std::string mc_note_;

    uint32_t* seq_ = nullptr;        // device, kNumChannels * kMaxBlocks
    uint32_t* status_host_ = nullptr; // pinned + mapped
    uint32_t* status_dev_ = nullptr;
    unsigned long long timeout_ns_ = 20ull * 1000 * 1000 * 1000;
};

} // namespace kf
