// kungfu_b200 — TMA bulk-copy pull of a peer's published blob (request_variable / PairAveraging prefetch).
//
// The LDG/STG pull keeps bytes in flight with registers (4 x 16 B per thread); a peer load over NVLink has 2-3 us of
// latency, so link bandwidth needs megabytes outstanding. Here ONE elected thread per CTA drives the Tensor Memory
// Accelerator: `cp.async.bulk.shared::cluster.global` pulls 32 KiB chunks of the peer's arena straight into shared memory
// (completion on an mbarrier via complete_tx), `cp.async.bulk.global.shared::cta` writes them to local HBM, four stages
// deep per CTA — 148 CTAs keep ~19 MB in flight with no register or LSU pressure. SASS: UBLKCP / SYNCS.
//
// Reference counterpart: the TCP request/response of `P2PEndpoint` (srcs/go/rchannel/handler/p2p.go:36-120).
#include "kernels.h"
#include "signal.cuh"

namespace kf {

namespace {

constexpr uint32_t kChunk = 32 * 1024;
constexpr int kStages = 4;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)), "l"(gsrc),
                 "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tma_store(void* gdst, const void* smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(smem_src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

struct TmaPullArgs {
    int target;
    PubStore ps;
    void* dst;
    size_t nbytes;
    uint32_t* ok; // [0] = version read (0: nothing published), [1] += torn reads, [2] = this copy is torn, [3] scratch
    int retry;    // second attempt: runs only when ok[2] is set
    const int* target_dev; // device word holding the target rank (null: `target`)
};

__global__ void __launch_bounds__(32) pull_tma_kernel(const __grid_constant__ MeshArgs m, const __grid_constant__ TmaPullArgs a) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t full[kStages];
    __shared__ uint32_t s_ver;
    if (a.retry && ld_relaxed_sys(a.ok + 2) == 0u) return; // the first attempt was clean
    const int target = a.target_dev ? *a.target_dev : a.target;
    const uint32_t* peer_ver = reinterpret_cast<const uint32_t*>(reinterpret_cast<const char*>(m.data[target]) + a.ps.ver_off);
    if (threadIdx.x == 0) {
        s_ver = ld_acquire_sys(peer_ver);
        for (int s = 0; s < kStages; ++s) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const uint32_t vr = s_ver;
    if (blockIdx.x == 0 && threadIdx.x == 0) a.ok[0] = vr;
    if (vr == 0 || threadIdx.x != 0) return;
    const char* src = reinterpret_cast<const char*>(m.data[target]) + a.ps.pub_off[vr % (uint32_t)a.ps.nbuf];
    char* dst = reinterpret_cast<char*>(a.dst);
    const size_t nchunks = (a.nbytes + kChunk - 1) / kChunk;
    const size_t mine = nchunks > blockIdx.x ? (nchunks - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    auto chunk_off = [&](size_t k) { return (blockIdx.x + k * gridDim.x) * (size_t)kChunk; };
    auto chunk_len = [&](size_t k) {
        const size_t off = chunk_off(k);
        return (uint32_t)(a.nbytes - off < kChunk ? a.nbytes - off : kChunk);
    };
    for (size_t k = 0; k < mine && k < (size_t)kStages; ++k) { // prologue: fill the pipeline
        mbar_expect_tx(&full[k], chunk_len(k));
        tma_load(smem + k * kChunk, src + chunk_off(k), chunk_len(k), &full[k]);
    }
    for (size_t k = 0; k < mine; ++k) {
        const int s = (int)(k % kStages);
        mbar_wait(&full[s], (uint32_t)((k / kStages) & 1));
        tma_store(dst + chunk_off(k), smem + (size_t)s * kChunk, chunk_len(k));
        tma_commit();
        const size_t nk = k + kStages;
        if (nk < mine) {
            tma_wait_read_all(); // the store has finished reading slot s; the other slots' loads are still in flight
            mbar_expect_tx(&full[s], chunk_len(nk));
            tma_load(smem + (size_t)s * kChunk, src + chunk_off(nk), chunk_len(nk), &full[s]);
        }
    }
    tma_wait_all();
    const uint32_t wr = ld_acquire_sys(peer_ver + 1);
    if ((int32_t)(wr - (vr + (uint32_t)a.ps.nbuf)) >= 0) { // the buffer we copied was being rewritten
        atomicAdd(&a.ok[1], 1u);
        atomicExch(a.retry ? &a.ok[3] : &a.ok[2], 1u);
    }
}

} // namespace

void launch_pull_tma(const MeshArgs& m, int target, const PubStore& ps, void* dst, size_t nbytes, uint32_t* ok, int retry, int blocks,
                     cudaStream_t st, const int* target_dev) {
    if (ps.nbuf < 2 || ps.nbuf > 3) throw std::runtime_error("kungfu_b200: the published-model store has 2 or 3 buffers");
    if (nbytes % 16) throw std::runtime_error("kungfu_b200: TMA pull needs a 16-byte multiple");
    if (!target_dev && (target < 0 || target >= m.nranks)) throw std::runtime_error("kungfu_b200: bad pull target");
    if (blocks < 1) blocks = 1;
    static bool configured = false;
    const size_t smem = (size_t)kChunk * kStages;
    if (!configured) {
        KF_CUDA_CHECK(cudaFuncSetAttribute(pull_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    TmaPullArgs a{target, ps, dst, nbytes, ok, retry, target_dev};
    pull_tma_kernel<<<blocks, 32, smem, st>>>(m, a);
    count_launch();
    KF_CUDA_CHECK(cudaGetLastError());
    if (retry) launch_pull_settle(ok, st);
}

} // namespace kf
