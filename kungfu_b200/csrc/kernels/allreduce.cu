// kungfu_b200 — K1/K6/K10/K11: all-reduce kernels over peer-mapped HBM (NVLink 5 / NVSwitch).
//
// Three algorithms, picked per message size by the host-side dispatcher (see ops/allreduce.py):
//   one-shot  : every rank loads the message from all peers and reduces locally. Latency path.
//   two-shot  : reduce-scatter by P2P loads into the slice owner, then all-gather by P2P stores.
//               In place on the symmetric arena; no barrier between the two phases is needed because
//               slice r of every buffer is read only by rank r and written only by rank r.
//   NVLS      : `multimem.ld_reduce` (switch fan-in) + `multimem.st` (switch fan-out) on the slice.
// Fused epilogues: scale (1/np averaging, TF S-SGD `g / np`, optimizers/sync_sgd.py:103-104) and
// squared-norm side reductions (gradient noise scale / variance monitors, optimizers/
// grad_noise_scale.py:11-88, ops/monitor.py:6-17) without an extra pass over the gradients.
//
// Reference call sites replaced: ncclAllReduce (srcs/cpp/src/nccl/gpu_collective.cpp:146-166) and the
// D2H → CPU all-reduce → H2D staging of the torch binding (srcs/cpp/src/torch/ops/cuda/
// collective.cpp:31-54).
#include "multimem.cuh"
#include "signal.cuh"
#include "kernels.h"

#include <atomic>
#include <stdexcept>
#include <string>
#include <type_traits>

namespace kf {

static std::atomic<uint64_t> g_launches{0};
void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }
uint64_t launch_count() { return g_launches.load(std::memory_order_relaxed); }

struct ArArgs {
    size_t off;      // byte offset of the message inside every rank's arena
    size_t out_off;  // sliced algorithms: byte offset the result is delivered to on every rank (== off: in place)
    size_t nvec;     // number of 16-byte vectors
    void* out;       // one-shot: local output (may be any device memory); others: unused
    float scale;     // applied to float-accumulated types when flags & kArScale
    float* sumsq;    // [0] += sum_r |x_r[slice]|^2 (or |x_local|^2 for NVLS), [1] += |result[slice]|^2
    int flags;
};

template <typename T> struct IsFloatAcc { static constexpr bool value = std::is_floating_point<typename Elem<T>::Acc>::value; };

template <typename T> __device__ __forceinline__ float vec_sumsq(const AccVec<T>& x) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < Elem<T>::kLanes; ++i) {
        const float f = (float)x.a[i];
        s = fmaf(f, f, s);
    }
    return s;
}

template <typename T> __device__ __forceinline__ void apply_scale(AccVec<T>& x, float scale) {
    if constexpr (IsFloatAcc<T>::value) {
#pragma unroll
        for (int i = 0; i < Elem<T>::kLanes; ++i) x.a[i] = (typename Elem<T>::Acc)(x.a[i] * scale);
    }
}

// Load vector `v` of the message from every peer (all loads issued before any use so that
// nranks NVLink requests are in flight per thread) and reduce. `pre` accumulates sum_r |x_r|^2.
// CANON: peers are visited in rank order 0..N-1 on every rank, so that ranks which all reduce the same data (one-shot)
// produce bit-identical results; the sliced algorithms have a single reducer per element and rotate for load balance.
template <typename T, int OP, bool SUMSQ, bool CANON = false>
__device__ __forceinline__ AccVec<T> gather_reduce(const MeshArgs& m, size_t byte_off, float& pre) {
    Vec16 buf[kMaxRanks];
#pragma unroll
    for (int i = 0; i < kMaxRanks; ++i) {
        if (i < m.nranks) {
            int r = CANON ? i : m.rank + i;
            if (r >= m.nranks) r -= m.nranks;
            buf[i] = ld_stream(reinterpret_cast<const char*>(m.data[r]) + byte_off);
        }
    }
    AccVec<T> acc = unpack<T>(buf[0]);
    if constexpr (SUMSQ) pre += vec_sumsq<T>(acc);
#pragma unroll
    for (int i = 1; i < kMaxRanks; ++i) {
        if (i < m.nranks) {
            AccVec<T> x = unpack<T>(buf[i]);
            if constexpr (SUMSQ) pre += vec_sumsq<T>(x);
            red_into<OP, T>(acc, x);
        }
    }
    return acc;
}

__device__ __forceinline__ void slice_bounds(size_t nvec, int nranks, int rank, size_t& s0, size_t& s1) {
    const size_t per = (nvec + nranks - 1) / nranks;
    s0 = (size_t)rank * per;
    if (s0 > nvec) s0 = nvec;
    s1 = s0 + per;
    if (s1 > nvec) s1 = nvec;
}

template <bool SUMSQ> __device__ __forceinline__ void flush_sumsq(const ArArgs& a, float pre, float post) {
    if constexpr (SUMSQ) {
        __shared__ float red[32];
        const float p0 = block_sum(pre, red);
        const float p1 = block_sum(post, red);
        if (threadIdx.x == 0) {
            atomicAdd(&a.sumsq[0], p0);
            atomicAdd(&a.sumsq[1], p1);
        }
    }
}

// --------------------------------------------------------------------------------------------
// one-shot
// --------------------------------------------------------------------------------------------
template <typename T, int OP, bool SUMSQ>
__global__ void __launch_bounds__(512) allreduce_oneshot_kernel(const __grid_constant__ MeshArgs m,
                                                                const __grid_constant__ ArArgs a) {
    uint32_t seq = load_seq(m);
    block_barrier<false>(m, seq);
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    float pre = 0.f, post = 0.f;
    for (size_t v = tid; v < a.nvec; v += stride) {
        AccVec<T> acc = gather_reduce<T, OP, SUMSQ, true>(m, a.off + v * 16, pre);
        if (a.flags & kArScale) apply_scale<T>(acc, a.scale);
        if constexpr (SUMSQ) post += vec_sumsq<T>(acc);
        st_plain(reinterpret_cast<char*>(a.out) + v * 16, pack<T>(acc));
    }
    // one-shot computes the full result on every rank: scale the side sums so that summing them over
    // ranks gives the same totals as the sliced algorithms.
    if constexpr (SUMSQ) {
        pre /= (float)m.nranks;
        post /= (float)m.nranks;
    }
    flush_sumsq<SUMSQ>(a, pre, post);
    block_barrier<false>(m, seq); // peers are done reading my staging slot
    store_seq(m, seq);
}

// --------------------------------------------------------------------------------------------
// two-shot, P2P loads (reduce-scatter) + P2P stores (all-gather), in place
// --------------------------------------------------------------------------------------------
template <typename T, int OP, bool SUMSQ>
__global__ void __launch_bounds__(512) allreduce_twoshot_kernel(const __grid_constant__ MeshArgs m,
                                                                const __grid_constant__ ArArgs a) {
    const bool gated = (a.flags & kArGated) != 0;
    uint32_t seq = 0;
    if (!gated) {
        seq = load_seq(m);
        block_barrier<false>(m, seq);
    }
    size_t s0, s1;
    slice_bounds(a.nvec, m.nranks, m.rank, s0, s1);
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    float pre = 0.f, post = 0.f;
    for (size_t v = s0 + tid; v < s1; v += stride) {
        const size_t bo = a.off + v * 16;
        const size_t oo = a.out_off + v * 16;
        AccVec<T> acc = gather_reduce<T, OP, SUMSQ>(m, bo, pre);
        if (a.flags & kArScale) apply_scale<T>(acc, a.scale);
        if constexpr (SUMSQ) post += vec_sumsq<T>(acc);
        const Vec16 res = pack<T>(acc);
#pragma unroll
        for (int i = 0; i < kMaxRanks; ++i) {
            if (i < m.nranks) {
                int r = m.rank + i;
                if (r >= m.nranks) r -= m.nranks;
                st_stream(reinterpret_cast<char*>(m.data[r]) + oo, res);
            }
        }
    }
    flush_sumsq<SUMSQ>(a, pre, post);
    if (gated) {
        signal_done(m, 0);
        return;
    }
    block_barrier<true>(m, seq); // all slices have landed in my buffer / my stores are visible
    store_seq(m, seq);
}

// --------------------------------------------------------------------------------------------
// NVLS: switch-side reduction and multicast
// --------------------------------------------------------------------------------------------
template <typename T, int OP, bool SUMSQ>
__global__ void __launch_bounds__(512) allreduce_nvls_kernel(const __grid_constant__ MeshArgs m,
                                                             const __grid_constant__ ArArgs a) {
    const bool gated = (a.flags & kArGated) != 0;
    uint32_t seq = gated ? 0u : load_seq(m);
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    float pre = 0.f, post = 0.f;
    if (SUMSQ && !(a.flags & kArSumSqPost)) {
        // |x_local|^2 over the whole local message. Block b touches exactly the vectors that the
        // peers' block b will later overwrite, and does so before it signals — no race with the
        // in-place multicast stores.
        for (int r = 0; r < m.nranks; ++r) {
            size_t q0, q1;
            slice_bounds(a.nvec, m.nranks, r, q0, q1);
            for (size_t v = q0 + tid; v < q1; v += stride) {
                const Vec16 x = ld_stream(reinterpret_cast<const char*>(m.data[m.rank]) + a.off + v * 16);
                pre += vec_sumsq<T>(unpack<T>(x));
            }
        }
    }
    if (!gated) block_barrier<false>(m, seq);
    size_t s0, s1;
    slice_bounds(a.nvec, m.nranks, m.rank, s0, s1);
    char* mc = reinterpret_cast<char*>(m.mc) + a.off;
    char* mco = reinterpret_cast<char*>(m.mc) + a.out_off;
    // kU switch reductions in flight per thread: a multimem.ld_reduce is a ~3 us round trip through the NVSwitch, so the bytes
    // in flight (threads x kU x 16 B), not the instruction rate, set the bandwidth — one load at a time left 32 CTAs at ~90 GB/s
    constexpr int kU = 4;
    for (size_t v0 = s0 + tid; v0 < s1; v0 += stride * kU) {
        Vec16 xs[kU];
#pragma unroll
        for (int u = 0; u < kU; ++u) {
            const size_t v = v0 + (size_t)u * stride;
            if (v < s1) xs[u] = Multimem<T, OP>::ld_reduce(mc + v * 16);
        }
#pragma unroll
        for (int u = 0; u < kU; ++u) {
            const size_t v = v0 + (size_t)u * stride;
            if (v >= s1) break;
            Vec16 x = xs[u];
            if ((a.flags & kArScale) || SUMSQ) {
                AccVec<T> acc = unpack<T>(x);
                if (a.flags & kArScale) apply_scale<T>(acc, a.scale);
                if constexpr (SUMSQ) post += vec_sumsq<T>(acc);
                x = pack<T>(acc);
            }
            multimem_st(mco + v * 16, x);
        }
    }
    flush_sumsq<SUMSQ>(a, pre, post);
    if (gated) {
        signal_done(m, 0);
        return;
    }
    block_barrier<true>(m, seq);
    store_seq(m, seq);
}


// --------------------------------------------------------------------------------------------
// LL one-shot: flag-carrying pushes, no barrier at all (latency path for small messages)
// --------------------------------------------------------------------------------------------
// Every rank PUSHES its message into a slot of every peer as 16-byte lines {data, flag, data, flag} (8 payload bytes per
// line, each 8-byte half validated by its own flag word — no reliance on 16-byte store atomicity), then polls its own slots
// until the lines of all N ranks carry this operation's flag, reduces them in rank order (bit-identical results everywhere)
// and writes the result straight into the user tensor. One NVLink one-way trip instead of barrier + load round trip + barrier.
// Two slot sets alternate by operation parity: a rank can start operation k+2 (which reuses the slots of k) only after it has
// received every peer's data of k+1, and a peer sends that only after it has finished reading k — so no trailing barrier is
// needed either. The operation counter lives in m.seq[kMaxBlocks + 1]; the last block to finish bumps it (ticket in +2), the
// next launch on the stream reads the new value.
struct LlArgs {
    const void* in;   // local input (any device memory, 8-byte aligned)
    void* out;        // local output
    size_t slot_off;  // arena offset of the slot area: [2 parities][kMaxRanks][max_lines] 16-byte lines
    size_t max_lines; // capacity of one (parity, rank) slot
    size_t nlines;    // 8-byte payload groups in this message
    float scale;
    int flags;
};

__device__ __forceinline__ void st_volatile_v4(void* p, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.volatile.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void ld_volatile_v4(const void* p, uint32_t& a, uint32_t& b, uint32_t& c, uint32_t& d) {
    asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "l"(p) : "memory");
}

template <typename T> struct Half8 { // 8 payload bytes as elements of T with their accumulator type
    static constexpr int kN = 8 / sizeof(T);
    typename Elem<T>::Acc a[kN];
};
template <typename T> __device__ __forceinline__ Half8<T> unpack_half(uint32_t lo, uint32_t hi) {
    uint32_t w[2] = {lo, hi};
    Half8<T> r;
    const T* e = reinterpret_cast<const T*>(w);
#pragma unroll
    for (int i = 0; i < Half8<T>::kN; ++i) r.a[i] = Elem<T>::up(e[i]);
    return r;
}

template <typename T, int OP>
__global__ void __launch_bounds__(512) allreduce_ll_kernel(const __grid_constant__ MeshArgs m, const __grid_constant__ LlArgs a) {
    const uint32_t op = m.seq[kMaxBlocks + 1] + 1u; // flag of this operation (never 0: slots start zeroed)
    const size_t par = op & 1u;
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const size_t slot_lines = a.max_lines;
    // ---- push my lines to every rank (own slot included: the reduce below treats all N sources alike)
    const uint2* in = reinterpret_cast<const uint2*>(a.in);
    for (size_t i = tid; i < a.nlines; i += stride) {
        const uint2 d = in[i];
#pragma unroll
        for (int k = 0; k < kMaxRanks; ++k) {
            if (k < m.nranks) {
                int r = m.rank + k;
                if (r >= m.nranks) r -= m.nranks;
                char* dst = reinterpret_cast<char*>(m.data[r]) + a.slot_off + ((par * kMaxRanks + (size_t)m.rank) * slot_lines + i) * 16;
                st_volatile_v4(dst, d.x, op, d.y, op);
            }
        }
    }
    // ---- poll + reduce in rank order
    uint2* out = reinterpret_cast<uint2*>(a.out);
    const char* mine = reinterpret_cast<const char*>(m.data[m.rank]) + a.slot_off + par * kMaxRanks * slot_lines * 16;
    for (size_t i = tid; i < a.nlines; i += stride) {
        Half8<T> acc;
#pragma unroll
        for (int r = 0; r < kMaxRanks; ++r) {
            if (r < m.nranks) {
                const char* src = mine + ((size_t)r * slot_lines + i) * 16;
                uint32_t d0, f0, d1, f1;
                ld_volatile_v4(src, d0, f0, d1, f1);
                if (f0 != op || f1 != op) {
                    const unsigned long long t0 = global_ns();
                    uint32_t iters = 0;
                    bool dead = false;
                    while (!dead) {
                        ld_volatile_v4(src, d0, f0, d1, f1);
                        if (f0 == op && f1 == op) break;
                        if (((++iters) & 0x3ff) != 0) continue;
                        // (host-mapped status word: a PCIe read, so only every 1024 polls)
                        if (m.status && *((volatile uint32_t*)&m.status[0]) != 0u) dead = true;
                        else if (global_ns() - t0 > m.timeout_ns) {
                            if (m.status) {
                                m.status[1] = ((uint32_t)m.rank << 16) | (uint32_t)r;
                                m.status[3] = op;
                                m.status[4] = f0;
                                m.status[2] = m.status[2] + 1u;
                                *((volatile uint32_t*)&m.status[0]) = kErrBarrierTimeout;
                                __threadfence_system();
                            }
                            dead = true;
                        }
                    }
                }
                const Half8<T> x = unpack_half<T>(d0, d1);
                if (r == 0) acc = x;
                else {
#pragma unroll
                    for (int j = 0; j < Half8<T>::kN; ++j) acc.a[j] = red_apply<OP>(acc.a[j], x.a[j]);
                }
            }
        }
        if constexpr (IsFloatAcc<T>::value) {
            if (a.flags & kArScale) {
#pragma unroll
                for (int j = 0; j < Half8<T>::kN; ++j) acc.a[j] = (typename Elem<T>::Acc)(acc.a[j] * a.scale);
            }
        }
        uint32_t w[2];
        T* e = reinterpret_cast<T*>(w);
#pragma unroll
        for (int j = 0; j < Half8<T>::kN; ++j) e[j] = Elem<T>::down(acc.a[j]);
        out[i] = make_uint2(w[0], w[1]);
    }
    // ---- last block out advances the operation counter
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const uint32_t prev = atomicAdd(&m.seq[kMaxBlocks + 2], 1u);
        if (prev == gridDim.x - 1) {
            m.seq[kMaxBlocks + 2] = 0u;
            m.seq[kMaxBlocks + 1] = op;
        }
    }
}

template <typename T> static void launch_ll_op(int op, const MeshArgs& m, const LlArgs& a, int blocks, int threads, cudaStream_t st) {
    switch (op) {
    case kSum: allreduce_ll_kernel<T, kSum><<<blocks, threads, 0, st>>>(m, a); break;
    case kMin: allreduce_ll_kernel<T, kMin><<<blocks, threads, 0, st>>>(m, a); break;
    case kMax: allreduce_ll_kernel<T, kMax><<<blocks, threads, 0, st>>>(m, a); break;
    case kProd: allreduce_ll_kernel<T, kProd><<<blocks, threads, 0, st>>>(m, a); break;
    default: throw std::runtime_error("kungfu_b200: bad reduce op");
    }
}

size_t allreduce_ll_slot_bytes(size_t max_payload_bytes) { return 2 * (size_t)kMaxRanks * ((max_payload_bytes + 7) / 8) * 16; }

void launch_allreduce_ll(const MeshArgs& m, const void* in, void* out, size_t nbytes, size_t slot_off, size_t max_payload_bytes, int dtype,
                         int op, float scale, int flags, cudaStream_t st) {
    if (nbytes == 0) return;
    if (nbytes % 8 != 0 || ((uintptr_t)in % 8) || ((uintptr_t)out % 8)) throw std::runtime_error("kungfu_b200: LL all-reduce needs 8-byte aligned, 8-byte padded messages");
    if (nbytes > max_payload_bytes) throw std::runtime_error("kungfu_b200: message exceeds the LL slot");
    if (slot_off % 16) throw std::runtime_error("kungfu_b200: LL slot area must be 16-byte aligned");
    LlArgs a{in, out, slot_off, (max_payload_bytes + 7) / 8, nbytes / 8, scale, flags};
    const int threads = a.nlines <= 128 ? 128 : 512;
    int blocks = (int)((a.nlines + threads - 1) / threads);
    if (blocks > 16) blocks = 16; // enough stores in flight for 64 KiB; more CTAs only add scheduling latency
    if (blocks < 1) blocks = 1;
    switch (dtype) {
    case kF32: launch_ll_op<float>(op, m, a, blocks, threads, st); break;
    case kBF16: launch_ll_op<__nv_bfloat16>(op, m, a, blocks, threads, st); break;
    case kF16: launch_ll_op<__half>(op, m, a, blocks, threads, st); break;
    case kF64: launch_ll_op<double>(op, m, a, blocks, threads, st); break;
    case kI32: launch_ll_op<int32_t>(op, m, a, blocks, threads, st); break;
    case kI64: launch_ll_op<int64_t>(op, m, a, blocks, threads, st); break;
    case kU8: launch_ll_op<uint8_t>(op, m, a, blocks, threads, st); break;
    case kI8: launch_ll_op<int8_t>(op, m, a, blocks, threads, st); break;
    case kU32: launch_ll_op<uint32_t>(op, m, a, blocks, threads, st); break;
    case kU64: launch_ll_op<uint64_t>(op, m, a, blocks, threads, st); break;
    case kI16: launch_ll_op<int16_t>(op, m, a, blocks, threads, st); break;
    case kU16: launch_ll_op<uint16_t>(op, m, a, blocks, threads, st); break;
    default: throw std::runtime_error("kungfu_b200: unsupported dtype");
    }
    count_launch();
    KF_CUDA_CHECK(cudaGetLastError());
}

// --------------------------------------------------------------------------------------------
// host dispatch
// --------------------------------------------------------------------------------------------
template <typename T, int OP, bool SUMSQ>
static void launch_typed(int algo, const MeshArgs& m, const ArArgs& a, int blocks, int threads, cudaStream_t st) {
    if (algo == kAlgoOneShot) {
        allreduce_oneshot_kernel<T, OP, SUMSQ><<<blocks, threads, 0, st>>>(m, a);
    } else if (algo == kAlgoTwoShot) {
        allreduce_twoshot_kernel<T, OP, SUMSQ><<<blocks, threads, 0, st>>>(m, a);
    } else if (algo == kAlgoNvls) {
        if constexpr (Multimem<T, OP>::kSupported) {
            allreduce_nvls_kernel<T, OP, SUMSQ><<<blocks, threads, 0, st>>>(m, a);
        } else {
            throw std::runtime_error("kungfu_b200: NVLS all-reduce unsupported for this dtype/op");
        }
    } else {
        throw std::runtime_error("kungfu_b200: unknown all-reduce algorithm");
    }
    count_launch();
    KF_CUDA_CHECK(cudaGetLastError());
}

template <typename T, bool SUMSQ>
static void launch_op(int op, int algo, const MeshArgs& m, const ArArgs& a, int blocks, int threads, cudaStream_t st) {
    switch (op) {
    case kSum: launch_typed<T, kSum, SUMSQ>(algo, m, a, blocks, threads, st); break;
    case kMin: launch_typed<T, kMin, false>(algo, m, a, blocks, threads, st); break;
    case kMax: launch_typed<T, kMax, false>(algo, m, a, blocks, threads, st); break;
    case kProd: launch_typed<T, kProd, false>(algo, m, a, blocks, threads, st); break;
    default: throw std::runtime_error("kungfu_b200: bad reduce op");
    }
}

bool nvls_supported(int dtype, int op) {
    switch (dtype) {
    case kF32: return op == kSum;
    case kBF16: case kF16: return op == kSum || op == kMin || op == kMax;
    default: return false;
    }
}

void launch_allreduce(int algo, const MeshArgs& m, size_t off, void* out, size_t nbytes, int dtype, int op, float scale,
                      float* sumsq, int flags, int blocks, int threads, cudaStream_t st, size_t out_off) {
    if (out_off == (size_t)-1) out_off = off;
    if (out_off % 16 != 0) throw std::runtime_error("kungfu_b200: all-reduce output offset must be 16-byte aligned");
    if ((flags & kArGated) && algo == kAlgoOneShot) throw std::runtime_error("kungfu_b200: the one-shot all-reduce has no split-phase form");
    if (nbytes % 16 != 0) throw std::runtime_error("kungfu_b200: all-reduce message must be padded to 16 bytes");
    if (off % 16 != 0) throw std::runtime_error("kungfu_b200: all-reduce offset must be 16-byte aligned");
    if (blocks < 1 || blocks > kMaxBlocks) throw std::runtime_error("kungfu_b200: bad block count");
    if (threads < m.nranks || threads > 512) throw std::runtime_error("kungfu_b200: bad thread count");
    if (algo == kAlgoNvls && m.mc == nullptr) throw std::runtime_error("kungfu_b200: NVLS requested but no multicast mapping");
    if (algo == kAlgoOneShot && out == nullptr) throw std::runtime_error("kungfu_b200: one-shot needs an output buffer");
    ArArgs a;
    a.off = off;
    a.out_off = out_off;
    a.nvec = nbytes / 16;
    a.out = out;
    a.scale = scale;
    a.sumsq = sumsq;
    a.flags = flags;
    const bool sq = (flags & kArSumSq) != 0;
    if (sq && (sumsq == nullptr || op != kSum)) throw std::runtime_error("kungfu_b200: sumsq epilogue needs op=sum and an output");
#define KF_DISPATCH(T)                                                                                                  \
    do {                                                                                                                \
        if (sq) launch_op<T, true>(op, algo, m, a, blocks, threads, st);                                                \
        else launch_op<T, false>(op, algo, m, a, blocks, threads, st);                                                  \
    } while (0)
    switch (dtype) {
    case kF32: KF_DISPATCH(float); break;
    case kBF16: KF_DISPATCH(__nv_bfloat16); break;
    case kF16: KF_DISPATCH(__half); break;
    case kF64: KF_DISPATCH(double); break;
    case kI32: KF_DISPATCH(int32_t); break;
    case kI64: KF_DISPATCH(int64_t); break;
    case kU8: KF_DISPATCH(uint8_t); break;
    case kI8: KF_DISPATCH(int8_t); break;
    case kU32: KF_DISPATCH(uint32_t); break;
    case kU64: KF_DISPATCH(uint64_t); break;
    case kI16: KF_DISPATCH(int16_t); break;
    case kU16: KF_DISPATCH(uint16_t); break;
    default: throw std::runtime_error("kungfu_b200: unsupported dtype");
    }
#undef KF_DISPATCH
}

} // namespace kf
