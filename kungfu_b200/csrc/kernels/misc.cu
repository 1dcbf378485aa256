// kungfu_b200 — K3/K12/K14: broadcast, all-gather and barrier over peer-mapped memory.
//
// Reference counterparts: `ncclBroadcast` (srcs/cpp/src/nccl/gpu_collective.cpp:124-144), the CPU
// `Broadcast` used by BroadcastGlobalVariablesOp (srcs/cpp/src/tensorflow/ops/cpu/collective.cpp:
// 185-210), direct all-to-all `AllGather` (srcs/go/kungfu/session/allgather.go:17-45) and `Barrier`
// (srcs/go/kungfu/session/session.go:96-113).
#include "kernels.h"
#include "multimem.cuh"
#include "signal.cuh"

namespace kf {

struct BcastArgs {
    size_t off, nvec;
    int root;
};

// NVLS: the root multicast-stores the region once (switch fan-out); P2P: every non-root pulls from
// the root (N-1 concurrent readers of the root's HBM, each at full NVLink rate).
template <bool NVLS>
__global__ void __launch_bounds__(512) broadcast_kernel(const __grid_constant__ MeshArgs m,
                                                        const __grid_constant__ BcastArgs a) {
    uint32_t seq = load_seq(m);
    block_barrier<false>(m, seq);
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    if constexpr (NVLS) {
        if (m.rank == a.root) {
            const char* src = reinterpret_cast<const char*>(m.data[m.rank]) + a.off;
            char* mc = reinterpret_cast<char*>(m.mc) + a.off;
            for (size_t v = tid; v < a.nvec; v += stride) multimem_st(mc + v * 16, ld_stream(src + v * 16));
        }
    } else {
        if (m.rank != a.root) {
            const char* src = reinterpret_cast<const char*>(m.data[a.root]) + a.off;
            char* dst = reinterpret_cast<char*>(m.data[m.rank]) + a.off;
            size_t v = tid;
            // two vectors in flight per thread
            for (; v + stride < a.nvec; v += 2 * stride) {
                const Vec16 x0 = ld_stream(src + v * 16);
                const Vec16 x1 = ld_stream(src + (v + stride) * 16);
                st_stream(dst + v * 16, x0);
                st_stream(dst + (v + stride) * 16, x1);
            }
            for (; v < a.nvec; v += stride) st_stream(dst + v * 16, ld_stream(src + v * 16));
        }
    }
    block_barrier<true>(m, seq);
    store_seq(m, seq);
}

void launch_broadcast(int algo, const MeshArgs& m, size_t off, size_t nbytes, int root, int blocks, int threads,
                      cudaStream_t st) {
    if (nbytes % 16 || off % 16) throw std::runtime_error("kungfu_b200: broadcast region must be 16-byte aligned/padded");
    if (blocks < 1 || blocks > kMaxBlocks || threads < m.nranks || threads > 512)
        throw std::runtime_error("kungfu_b200: bad launch shape");
    if (root < 0 || root >= m.nranks) throw std::runtime_error("kungfu_b200: bad broadcast root");
    BcastArgs a{off, nbytes / 16, root};
    if (algo == kAlgoNvls) {
        if (!m.mc) throw std::runtime_error("kungfu_b200: NVLS requested but no multicast mapping");
        broadcast_kernel<true><<<blocks, threads, 0, st>>>(m, a);
    } else {
        broadcast_kernel<false><<<blocks, threads, 0, st>>>(m, a);
    }
    count_launch();
    KF_CUDA_CHECK(cudaGetLastError());
}

struct AgArgs {
    size_t in_off, out_off, nvec; // nvec per rank
};

// Every rank delivers its chunk into slot `rank` of every peer's output region.
template <bool NVLS>
__global__ void __launch_bounds__(512) allgather_kernel(const __grid_constant__ MeshArgs m,
                                                        const __grid_constant__ AgArgs a) {
    uint32_t seq = load_seq(m);
    block_barrier<false>(m, seq);
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const char* src = reinterpret_cast<const char*>(m.data[m.rank]) + a.in_off;
    const size_t slot = a.out_off + (size_t)m.rank * a.nvec * 16;
    for (size_t v = tid; v < a.nvec; v += stride) {
        const Vec16 x = ld_stream(src + v * 16);
        if constexpr (NVLS) {
            multimem_st(reinterpret_cast<char*>(m.mc) + slot + v * 16, x);
        } else {
#pragma unroll
            for (int i = 0; i < kMaxRanks; ++i) {
                if (i < m.nranks) {
                    int r = m.rank + i;
                    if (r >= m.nranks) r -= m.nranks;
                    st_stream(reinterpret_cast<char*>(m.data[r]) + slot + v * 16, x);
                }
            }
        }
    }
    block_barrier<true>(m, seq);
    store_seq(m, seq);
}

void launch_allgather(int algo, const MeshArgs& m, size_t in_off, size_t out_off, size_t nbytes_per_rank, int blocks,
                      int threads, cudaStream_t st) {
    if (nbytes_per_rank % 16 || in_off % 16 || out_off % 16)
        throw std::runtime_error("kungfu_b200: all-gather regions must be 16-byte aligned/padded");
    if (blocks < 1 || blocks > kMaxBlocks || threads < m.nranks || threads > 512)
        throw std::runtime_error("kungfu_b200: bad launch shape");
    AgArgs a{in_off, out_off, nbytes_per_rank / 16};
    if (algo == kAlgoNvls) {
        if (!m.mc) throw std::runtime_error("kungfu_b200: NVLS requested but no multicast mapping");
        allgather_kernel<true><<<blocks, threads, 0, st>>>(m, a);
    } else {
        allgather_kernel<false><<<blocks, threads, 0, st>>>(m, a);
    }
    count_launch();
    KF_CUDA_CHECK(cudaGetLastError());
}

__global__ void __launch_bounds__(32) barrier_kernel(const __grid_constant__ MeshArgs m) {
    uint32_t seq = load_seq(m);
    block_barrier<true>(m, seq);
    store_seq(m, seq);
}

// The tail of a split-phase collective: ONE thread waits for the body kernels' blocks (all ranks) to report completion.
__global__ void __launch_bounds__(32) wait_done_kernel(const __grid_constant__ MeshArgs m, int word, uint32_t add) {
    if (threadIdx.x == 0) wait_done(m, word, add);
}

void launch_wait_done(const MeshArgs& m, int word, uint32_t add, cudaStream_t st) {
    if (word < 0 || word >= kDoneWords || word >= kSeqWords - kMaxBlocks) throw std::runtime_error("kungfu_b200: bad completion word");
    wait_done_kernel<<<1, 32, 0, st>>>(m, word, add);
    count_launch();
    KF_CUDA_CHECK(cudaGetLastError());
}

void launch_barrier(const MeshArgs& m, cudaStream_t st) {
    barrier_kernel<<<1, 32, 0, st>>>(m);
    count_launch();
    KF_CUDA_CHECK(cudaGetLastError());
}

} // namespace kf
