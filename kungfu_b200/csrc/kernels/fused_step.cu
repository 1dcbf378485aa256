// kungfu_b200 — K1b/K8/K9: optimizer steps fused with their collective, one kernel each.
//
//  ssgd_fused_kernel : reduce-scatter(gradients) -> optimizer update on the owned shard (fp32 master
//                      weights + optimizer state are SHARDED across ranks) -> all-gather(parameters).
//                      Replaces all-reduce + `g / np` + optimizer (optimizers/sync_sgd.py:78-109,
//                      torch/optimizers/sync_sgd.py:12-26) and moves the same NVLink bytes as a bare
//                      all-reduce.
//  sma_fused_kernel  : all-reduce(model)/np -> v <- (1-a) v + a avg -> local optimizer step with the
//                      LOCAL gradients (optimizers/sma_sgd.py:60-67), one pass over the arena.
//  pair_avg_kernel   : asynchronous pull of ONE peer's published model over NVLink, v <- 0.5 (v + other),
//                      local step, publish (optimizers/async_sgd.py:110-142). The blob store + RWMutex of
//                      the reference (srcs/go/store/store.go:14-59, rchannel/handler/p2p.go:94-120)
//                      becomes a double-buffered device arena with version words; the target is never
//                      blocked by a reader.
#include "kernels.h"
#include "multimem.cuh"
#include "optim.cuh"
#include "signal.cuh"

namespace kf {

// ------------------------------------------------------------------------------------------------
// helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void slice_bounds2(size_t nvec, int nranks, int rank, size_t& s0, size_t& s1) {
    const size_t per = (nvec + nranks - 1) / nranks;
    s0 = (size_t)rank * per;
    if (s0 > nvec) s0 = nvec;
    s1 = s0 + per;
    if (s1 > nvec) s1 = nvec;
}

// load / store LANES consecutive elements of runtime dtype (f32 / bf16 / f16) at element index idx
template <int LANES> __device__ __forceinline__ void load_lanes_rt(const void* base, int dtype, size_t idx, float* out) {
    if (dtype == kF32) {
        const float* p = reinterpret_cast<const float*>(base) + idx;
#pragma unroll
        for (int i = 0; i < LANES / 4; ++i) {
            const float4 q = *reinterpret_cast<const float4*>(p + 4 * i);
            out[4 * i + 0] = q.x;
            out[4 * i + 1] = q.y;
            out[4 * i + 2] = q.z;
            out[4 * i + 3] = q.w;
        }
    } else {
        const uint16_t* p = reinterpret_cast<const uint16_t*>(base) + idx;
#pragma unroll
        for (int i = 0; i < LANES / 4; ++i) {
            const uint2 q = *reinterpret_cast<const uint2*>(p + 4 * i);
            const uint32_t w[2] = {q.x, q.y};
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                if (dtype == kBF16) {
                    const __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162*>(&w[j]);
                    out[4 * i + 2 * j + 0] = __low2float(h);
                    out[4 * i + 2 * j + 1] = __high2float(h);
                } else {
                    const __half2 h = *reinterpret_cast<const __half2*>(&w[j]);
                    out[4 * i + 2 * j + 0] = __low2float(h);
                    out[4 * i + 2 * j + 1] = __high2float(h);
                }
            }
        }
    }
}
template <int LANES> __device__ __forceinline__ void store_lanes_rt(void* base, int dtype, size_t idx, const float* in) {
    if (dtype == kF32) {
        float* p = reinterpret_cast<float*>(base) + idx;
#pragma unroll
        for (int i = 0; i < LANES / 4; ++i)
            *reinterpret_cast<float4*>(p + 4 * i) = make_float4(in[4 * i], in[4 * i + 1], in[4 * i + 2], in[4 * i + 3]);
    } else {
        uint16_t* p = reinterpret_cast<uint16_t*>(base) + idx;
#pragma unroll
        for (int i = 0; i < LANES / 4; ++i) {
            uint2 q;
            if (dtype == kBF16) {
                const __nv_bfloat162 a = __floats2bfloat162_rn(in[4 * i], in[4 * i + 1]);
                const __nv_bfloat162 b = __floats2bfloat162_rn(in[4 * i + 2], in[4 * i + 3]);
                q.x = *reinterpret_cast<const uint32_t*>(&a);
                q.y = *reinterpret_cast<const uint32_t*>(&b);
            } else {
                const __half2 a = __floats2half2_rn(in[4 * i], in[4 * i + 1]);
                const __half2 b = __floats2half2_rn(in[4 * i + 2], in[4 * i + 3]);
                q.x = *reinterpret_cast<const uint32_t*>(&a);
                q.y = *reinterpret_cast<const uint32_t*>(&b);
            }
            *reinterpret_cast<uint2*>(p + 4 * i) = q;
        }
    }
}

template <typename T> __device__ __forceinline__ AccVec<T> gather_sum(const MeshArgs& m, size_t byte_off, float* pre) {
    Vec16 buf[kMaxRanks];
#pragma unroll
    for (int i = 0; i < kMaxRanks; ++i) {
        if (i < m.nranks) {
            int r = m.rank + i;
            if (r >= m.nranks) r -= m.nranks;
            buf[i] = ld_stream(reinterpret_cast<const char*>(m.data[r]) + byte_off);
        }
    }
    AccVec<T> acc = unpack<T>(buf[0]);
    if (pre) {
#pragma unroll
        for (int j = 0; j < Elem<T>::kLanes; ++j) *pre = fmaf((float)acc.a[j], (float)acc.a[j], *pre);
    }
#pragma unroll
    for (int i = 1; i < kMaxRanks; ++i) {
        if (i < m.nranks) {
            const AccVec<T> x = unpack<T>(buf[i]);
#pragma unroll
            for (int j = 0; j < Elem<T>::kLanes; ++j) {
                if (pre) *pre = fmaf((float)x.a[j], (float)x.a[j], *pre);
                acc.a[j] += x.a[j];
            }
        }
    }
    return acc;
}

__device__ __forceinline__ void scatter_store(const MeshArgs& m, size_t byte_off, const Vec16& res) {
#pragma unroll
    for (int i = 0; i < kMaxRanks; ++i) {
        if (i < m.nranks) {
            int r = m.rank + i;
            if (r >= m.nranks) r -= m.nranks;
            st_stream(reinterpret_cast<char*>(m.data[r]) + byte_off, res);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// S-SGD fused: RS(grads) + optimizer(shard) + AG(params)
// ------------------------------------------------------------------------------------------------
struct SsgdArgs {
    size_t grad_off, param_off, nvec;
    float* master; // fp32 master shard of the owned slice (nullptr: parameters are fp32 and live in the arena)
    float* s1;     // momentum / exp_avg shard (nullptr: stateless)
    float* s2;     // exp_avg_sq shard
    const float* hyper;
    float* sumsq;
    int flags;
};

template <typename T, int KIND, bool NVLS>
__global__ void __launch_bounds__(512) ssgd_fused_kernel(const __grid_constant__ MeshArgs m,
                                                         const __grid_constant__ SsgdArgs a) {
    constexpr int L = Elem<T>::kLanes;
    const bool gated = (a.flags & kArGated) != 0;
    uint32_t seq = gated ? 0u : load_seq(m);
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const bool sumsq = (a.flags & kArSumSq) != 0;
    float pre = 0.f, post = 0.f;
    if (NVLS && sumsq && !(a.flags & kArSumSqPost)) {
        for (int r = 0; r < m.nranks; ++r) {
            size_t q0, q1;
            slice_bounds2(a.nvec, m.nranks, r, q0, q1);
            for (size_t v = q0 + tid; v < q1; v += stride) {
                const AccVec<T> x =
                    unpack<T>(ld_stream(reinterpret_cast<const char*>(m.data[m.rank]) + a.grad_off + v * 16));
#pragma unroll
                for (int j = 0; j < L; ++j) pre = fmaf((float)x.a[j], (float)x.a[j], pre);
            }
        }
    }
    if (!gated) block_barrier<false>(m, seq);
    const Hyper h = load_hyper(a.hyper);
    const float gs_scale = h.v[kHpGradScale];
    size_t s0, s1;
    slice_bounds2(a.nvec, m.nranks, m.rank, s0, s1);
    // kU gradient reductions in flight per thread (switch round trip ~3 us: bytes in flight set the bandwidth), then the
    // optimizer math on each of them while the next batch's loads are already queued behind
    constexpr int kU = NVLS ? 4 : 2;
    for (size_t v0 = s0 + tid; v0 < s1; v0 += stride * kU) {
        AccVec<T> gs[kU];
#pragma unroll
        for (int u = 0; u < kU; ++u) {
            const size_t v = v0 + (size_t)u * stride;
            if (v < s1) {
                if constexpr (NVLS) gs[u] = unpack<T>(Multimem<T, kSum>::ld_reduce(reinterpret_cast<const char*>(m.mc) + a.grad_off + v * 16));
                else gs[u] = gather_sum<T>(m, a.grad_off + v * 16, (sumsq && !(a.flags & kArSumSqPost)) ? &pre : nullptr);
            }
        }
#pragma unroll
        for (int u = 0; u < kU; ++u) {
            const size_t v = v0 + (size_t)u * stride;
            if (v >= s1) break;
            const AccVec<T>& g = gs[u];
            const size_t idx = (v - s0) * L;
            F32Lanes<L> w, m1, m2;
            if (a.master) w.load(a.master + idx);
            else {
                const AccVec<T> p = unpack<T>(ld_plain(reinterpret_cast<const char*>(m.data[m.rank]) + a.param_off + v * 16));
#pragma unroll
                for (int j = 0; j < L; ++j) w.f[j] = (float)p.a[j];
            }
            if (a.s1) m1.load(a.s1 + idx);
            else m1.zero();
            if (KIND == kOptAdam && a.s2) m2.load(a.s2 + idx);
            else m2.zero();
            AccVec<T> out;
#pragma unroll
            for (int j = 0; j < L; ++j) {
                const float gj = (float)g.a[j] * gs_scale;
                post = fmaf(gj, gj, post);
                opt_update<KIND>(w.f[j], gj, m1.f[j], m2.f[j], h);
                out.a[j] = (typename Elem<T>::Acc)w.f[j];
            }
            if (a.master) w.store(a.master + idx);
            if (a.s1) m1.store(a.s1 + idx);
            if (KIND == kOptAdam && a.s2) m2.store(a.s2 + idx);
            const Vec16 res = pack<T>(out);
            if constexpr (NVLS) multimem_st(reinterpret_cast<char*>(m.mc) + a.param_off + v * 16, res);
            else scatter_store(m, a.param_off + v * 16, res);
        }
    }
    if (sumsq) {
        __shared__ float red[32];
        const float p0 = block_sum(pre, red);
        const float p1 = block_sum(post, red);
        if (threadIdx.x == 0) {
            atomicAdd(&a.sumsq[0], p0);
            atomicAdd(&a.sumsq[1], p1);
        }
    }
    if (gated) {
        signal_done(m, 0);
        return;
    }
    block_barrier<true>(m, seq);
    store_seq(m, seq);
}

template <typename T, int KIND>
static void launch_ssgd_typed(int algo, const MeshArgs& m, const SsgdArgs& a, int blocks, int threads, cudaStream_t st) {
    if (algo == kAlgoNvls) {
        if constexpr (Multimem<T, kSum>::kSupported) ssgd_fused_kernel<T, KIND, true><<<blocks, threads, 0, st>>>(m, a);
        else throw std::runtime_error("kungfu_b200: NVLS unsupported for dtype");
    } else {
        ssgd_fused_kernel<T, KIND, false><<<blocks, threads, 0, st>>>(m, a);
    }
    count_launch();
    KF_CUDA_CHECK(cudaGetLastError());
}

static void check_launch_shape(const MeshArgs& m, size_t nbytes, int blocks, int threads) {
    if (nbytes % 16 != 0) throw std::runtime_error("kungfu_b200: region must be padded to 16 bytes");
    if (blocks < 1 || blocks > kMaxBlocks) throw std::runtime_error("kungfu_b200: bad block count");
    if (threads < 32 || threads > 512 || threads < m.nranks) throw std::runtime_error("kungfu_b200: bad thread count");
}

void launch_ssgd_fused(int algo, const MeshArgs& m, size_t grad_off, size_t param_off, int dtype, size_t nbytes,
                       float* master_shard, float* s1_shard, float* s2_shard, const float* hyper, int opt_kind,
                       float* sumsq, int flags, int blocks, int threads, cudaStream_t st) {
    check_launch_shape(m, nbytes, blocks, threads);
    if (grad_off % 16 || param_off % 16) throw std::runtime_error("kungfu_b200: offsets must be 16-byte aligned");
    if (algo == kAlgoNvls && !m.mc) throw std::runtime_error("kungfu_b200: NVLS requested but no multicast mapping");
    if ((flags & kArSumSq) && !sumsq) throw std::runtime_error("kungfu_b200: sumsq output missing");
    if (dtype != kF32 && master_shard == nullptr) throw std::runtime_error("kungfu_b200: 16-bit parameters need an fp32 master shard");
    SsgdArgs a{grad_off, param_off, nbytes / 16, master_shard, s1_shard, s2_shard, hyper, sumsq, flags};
    if (dtype == kF32) {
        if (opt_kind == kOptSgd) launch_ssgd_typed<float, kOptSgd>(algo, m, a, blocks, threads, st);
        else launch_ssgd_typed<float, kOptAdam>(algo, m, a, blocks, threads, st);
    } else if (dtype == kBF16) {
        if (opt_kind == kOptSgd) launch_ssgd_typed<__nv_bfloat16, kOptSgd>(algo, m, a, blocks, threads, st);
        else launch_ssgd_typed<__nv_bfloat16, kOptAdam>(algo, m, a, blocks, threads, st);
    } else if (dtype == kF16) {
        if (opt_kind == kOptSgd) launch_ssgd_typed<__half, kOptSgd>(algo, m, a, blocks, threads, st);
        else launch_ssgd_typed<__half, kOptAdam>(algo, m, a, blocks, threads, st);
    } else {
        throw std::runtime_error("kungfu_b200: fused S-SGD supports f32/bf16/f16 parameters");
    }
}

// ------------------------------------------------------------------------------------------------
// SMA fused
// ------------------------------------------------------------------------------------------------
struct SmaArgs {
    size_t param_off, avg_off, nvec;
    const void* grads;
    int grad_dtype;
    float* s1;
    float* s2;
    void* model_out;
    int model_dtype;
    const float* hyper;
};

// blend + local step on vector v: p <- p + alpha (avg - p), then the optimizer update with the LOCAL gradient
template <typename T, int KIND>
__device__ __forceinline__ void sma_blend_vec(char* P, const char* A, size_t v, const SmaArgs& a, const Hyper& h) {
    constexpr int L = Elem<T>::kLanes;
    const float alpha = h.v[kHpAlpha];
    const AccVec<T> p = unpack<T>(ld_plain(P + v * 16));
    const AccVec<T> av = unpack<T>(ld_plain(A + v * 16));
    const size_t idx = v * L;
    float g[L];
    if (a.grads) load_lanes_rt<L>(a.grads, a.grad_dtype, idx, g);
    F32Lanes<L> m1, m2;
    if (a.s1) m1.load(a.s1 + idx);
    else m1.zero();
    if (KIND == kOptAdam && a.s2) m2.load(a.s2 + idx);
    else m2.zero();
    AccVec<T> out;
    float wf[L];
#pragma unroll
    for (int j = 0; j < L; ++j) {
        float w = fmaf(alpha, (float)av.a[j] - (float)p.a[j], (float)p.a[j]);
        if (a.grads) opt_update<KIND>(w, g[j] * h.v[kHpGradScale], m1.f[j], m2.f[j], h);
        wf[j] = w;
        out.a[j] = (typename Elem<T>::Acc)w;
    }
    st_plain(P + v * 16, pack<T>(out));
    if (a.grads && a.s1) m1.store(a.s1 + idx);
    if (a.grads && KIND == kOptAdam && a.s2) m2.store(a.s2 + idx);
    if (a.model_out) store_lanes_rt<L>(a.model_out, a.model_dtype, idx, wf);
}

// LOCAL blend + step over a whole region (the overlapped SMA: the average was produced earlier by an out-of-place
// all-reduce that ran under forward/backward)
template <typename T, int KIND>
__global__ void __launch_bounds__(512) sma_blend_kernel(char* P, const char* A, const __grid_constant__ SmaArgs a) {
    const Hyper h = load_hyper(a.hyper);
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t v = tid; v < a.nvec; v += stride) sma_blend_vec<T, KIND>(P, A, v, a, h);
}

template <typename T, int KIND, bool NVLS>
__global__ void __launch_bounds__(512) sma_fused_kernel(const __grid_constant__ MeshArgs m,
                                                        const __grid_constant__ SmaArgs a) {
    constexpr int L = Elem<T>::kLanes;
    uint32_t seq = load_seq(m);
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    block_barrier<false>(m, seq);
    const Hyper h = load_hyper(a.hyper);
    const float inv_n = 1.f / (float)m.nranks;
    size_t s0, s1;
    slice_bounds2(a.nvec, m.nranks, m.rank, s0, s1);
    // phase 1+2: average of the owned slice, delivered into every rank's scratch region
    for (size_t v = s0 + tid; v < s1; v += stride) {
        AccVec<T> acc;
        if constexpr (NVLS) acc = unpack<T>(Multimem<T, kSum>::ld_reduce(reinterpret_cast<const char*>(m.mc) + a.param_off + v * 16));
        else acc = gather_sum<T>(m, a.param_off + v * 16, nullptr);
#pragma unroll
        for (int j = 0; j < L; ++j) acc.a[j] = (typename Elem<T>::Acc)((float)acc.a[j] * inv_n);
        const Vec16 res = pack<T>(acc);
        if constexpr (NVLS) multimem_st(reinterpret_cast<char*>(m.mc) + a.avg_off + v * 16, res);
        else scatter_store(m, a.avg_off + v * 16, res);
    }
    block_barrier<true>(m, seq);
    // phase 3: blend + local step. Block b touches, inside slice r, exactly the vectors that block b
    // of rank r read (phase 1) and wrote (phase 2) — the only blocks it has synchronised with.
    char* P = reinterpret_cast<char*>(m.data[m.rank]) + a.param_off;
    const char* A = reinterpret_cast<const char*>(m.data[m.rank]) + a.avg_off;
    for (int r = 0; r < m.nranks; ++r) {
        size_t q0, q1;
        slice_bounds2(a.nvec, m.nranks, r, q0, q1);
        for (size_t v = q0 + tid; v < q1; v += stride) sma_blend_vec<T, KIND>(P, A, v, a, h);
    }
    store_seq(m, seq);
}

template <typename T, int KIND>
static void launch_sma_typed(int algo, const MeshArgs& m, const SmaArgs& a, int blocks, int threads, cudaStream_t st) {
    if (algo == kAlgoNvls) {
        if constexpr (Multimem<T, kSum>::kSupported) sma_fused_kernel<T, KIND, true><<<blocks, threads, 0, st>>>(m, a);
        else throw std::runtime_error("kungfu_b200: NVLS unsupported for dtype");
    } else {
        sma_fused_kernel<T, KIND, false><<<blocks, threads, 0, st>>>(m, a);
    }
    count_launch();
    KF_CUDA_CHECK(cudaGetLastError());
}

void launch_sma_fused(int algo, const MeshArgs& m, size_t param_off, size_t avg_off, int dtype, size_t nbytes,
                      const void* grads, int grad_dtype, float* s1, float* s2, void* model_out, int model_dtype,
                      const float* hyper, int opt_kind, int blocks, int threads, cudaStream_t st) {
    check_launch_shape(m, nbytes, blocks, threads);
    if (param_off % 16 || avg_off % 16) throw std::runtime_error("kungfu_b200: offsets must be 16-byte aligned");
    if (algo == kAlgoNvls && !m.mc) throw std::runtime_error("kungfu_b200: NVLS requested but no multicast mapping");
    SmaArgs a{param_off, avg_off, nbytes / 16, grads, grad_dtype, s1, s2, model_out, model_dtype, hyper};
    if (dtype == kF32) {
        if (opt_kind == kOptSgd) launch_sma_typed<float, kOptSgd>(algo, m, a, blocks, threads, st);
        else launch_sma_typed<float, kOptAdam>(algo, m, a, blocks, threads, st);
    } else if (dtype == kBF16) {
        if (opt_kind == kOptSgd) launch_sma_typed<__nv_bfloat16, kOptSgd>(algo, m, a, blocks, threads, st);
        else launch_sma_typed<__nv_bfloat16, kOptAdam>(algo, m, a, blocks, threads, st);
    } else {
        throw std::runtime_error("kungfu_b200: fused SMA supports f32/bf16 parameters");
    }
}

void launch_sma_blend(void* params, const void* avg, int dtype, size_t nbytes, const void* grads, int grad_dtype, float* s1, float* s2,
                      const float* hyper, int opt_kind, int blocks, int threads, cudaStream_t st) {
    if (nbytes % 16 != 0) throw std::runtime_error("kungfu_b200: region must be padded to 16 bytes");
    if (blocks < 1 || threads < 32 || threads > 512) throw std::runtime_error("kungfu_b200: bad launch shape");
    SmaArgs a{0, 0, nbytes / 16, grads, grad_dtype, s1, s2, nullptr, dtype, hyper};
    char* P = reinterpret_cast<char*>(params);
    const char* A = reinterpret_cast<const char*>(avg);
    if (dtype == kF32) {
        if (opt_kind == kOptSgd) sma_blend_kernel<float, kOptSgd><<<blocks, threads, 0, st>>>(P, A, a);
        else sma_blend_kernel<float, kOptAdam><<<blocks, threads, 0, st>>>(P, A, a);
    } else if (dtype == kBF16) {
        if (opt_kind == kOptSgd) sma_blend_kernel<__nv_bfloat16, kOptSgd><<<blocks, threads, 0, st>>>(P, A, a);
        else sma_blend_kernel<__nv_bfloat16, kOptAdam><<<blocks, threads, 0, st>>>(P, A, a);
    } else {
        throw std::runtime_error("kungfu_b200: SMA blend supports f32/bf16 parameters");
    }
    count_launch();
    KF_CUDA_CHECK(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------
// PairAveraging: pull + average + step + publish
// ------------------------------------------------------------------------------------------------
struct PairArgs {
    int target;
    PubStore ps;
    void* params;
    size_t nvec;
    const void* grads;
    int grad_dtype;
    float* s1;
    float* s2;
    float* master;              // pair_avg: fp32 master copy of 16-bit parameters (nullptr: update the parameters in place)
    const float* hyper;
    int do_step;
    uint32_t* done_counter;
    void* dst;                  // pull: destination
    uint32_t* ok;               // pull: [0] = version read, [1] += torn reads, [2] = this copy is torn
    int retry;                  // pull: second attempt, runs only when ok[2] is set
    const int* target_dev;      // device word holding the target rank (CUDA graphs: the target changes between replays); null: `target`
    const void* pulled;         // pair_avg: private copy of the target's model (nullptr: read the target over NVLink)
    const uint32_t* pulled_ok;  // pair_avg: the `ok` words of the pull that filled `pulled`
};

__device__ __forceinline__ uint32_t* ver_words(const MeshArgs& m, int rank, size_t ver_off) {
    return reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(m.data[rank]) + ver_off);
}

// Marks the start of a publish (all blocks write the same value) and returns (version to read from
// the target, version this rank is about to publish).
__device__ __forceinline__ void pair_versions(const MeshArgs& m, int target, size_t ver_off, bool writer, uint32_t* sh) {
    if (threadIdx.x == 0) {
        uint32_t* my_ver = ver_words(m, m.rank, ver_off);
        sh[0] = 0u;
        if (target >= 0) sh[0] = ld_acquire_sys(ver_words(m, target, ver_off));
        if (writer) {
            sh[1] = ld_relaxed_sys(my_ver) + 1;
            st_release_sys(my_ver + 1, sh[1]);
        }
    }
    __syncthreads();
}

// Last block to finish flips the published version (threadFenceReduction pattern, system scope so
// that a remote reader that acquires the version word also sees the payload).
__device__ __forceinline__ void pair_commit(const MeshArgs& m, size_t ver_off, uint32_t next, uint32_t* done_counter) {
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t prev = atomicAdd(done_counter, 1u);
        if (prev == gridDim.x - 1) {
            *done_counter = 0;
            __threadfence_system();
            st_release_sys(ver_words(m, m.rank, ver_off), next);
        }
    }
}

// the buffer holding version vr has been handed back to the writer once it starts on version vr + nbuf
__device__ __forceinline__ bool pair_torn(const MeshArgs& m, int target, const PubStore& ps, uint32_t vr) {
    const uint32_t wr = ld_acquire_sys(ver_words(m, target, ps.ver_off) + 1);
    return (int32_t)(wr - (vr + (uint32_t)ps.nbuf)) >= 0;
}

template <typename T, int KIND>
__global__ void __launch_bounds__(512) pair_avg_kernel(const __grid_constant__ MeshArgs m,
                                                       const __grid_constant__ PairArgs a) {
    constexpr int L = Elem<T>::kLanes;
    __shared__ uint32_t sh[2];
    const bool remote = a.pulled == nullptr;
    const int target = a.target_dev ? *a.target_dev : a.target;
    pair_versions(m, remote ? target : -1, a.ps.ver_off, true, sh);
    uint32_t vr = sh[0];
    if (!remote) vr = (a.pulled_ok[2] == 0u) ? a.pulled_ok[0] : 0u; // a torn private copy is not averaged in
    const uint32_t next = sh[1];
    const char* src = remote ? reinterpret_cast<const char*>(m.data[target]) + a.ps.pub_off[vr % (uint32_t)a.ps.nbuf]
                             : reinterpret_cast<const char*>(a.pulled);
    char* pub = reinterpret_cast<char*>(m.data[m.rank]) + a.ps.pub_off[next % (uint32_t)a.ps.nbuf];
    char* P = reinterpret_cast<char*>(a.params);
    const Hyper h = load_hyper(a.hyper);
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    // four reads of the other model are issued per thread before any of them is consumed: a peer load has ~2-3 us of
    // latency, so the in-kernel pull needs megabytes in flight to approach link bandwidth
    constexpr int U = 4;
    for (size_t v0 = tid; v0 < a.nvec; v0 += stride * U) {
        Vec16 ovs[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const size_t v = v0 + (size_t)u * stride;
            if (vr != 0 && v < a.nvec) ovs[u] = ld_stream(src + v * 16);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const size_t v = v0 + (size_t)u * stride;
            if (v >= a.nvec) break;
            AccVec<T> p = unpack<T>(ld_plain(P + v * 16));
            const size_t idx = v * L;
            float g[L];
            if (a.do_step) load_lanes_rt<L>(a.grads, a.grad_dtype, idx, g);
            F32Lanes<L> m1, m2, mw;
            if (a.do_step && a.s1) m1.load(a.s1 + idx);
            else m1.zero();
            if (a.do_step && KIND == kOptAdam && a.s2) m2.load(a.s2 + idx);
            else m2.zero();
            if (a.master) { // the fp32 master is the model; the 16-bit parameters are its rounded image
                mw.load(a.master + idx);
#pragma unroll
                for (int j = 0; j < L; ++j) p.a[j] = (typename Elem<T>::Acc)mw.f[j];
            }
            if (vr != 0) {
                const AccVec<T> o = unpack<T>(ovs[u]);
#pragma unroll
                for (int j = 0; j < L; ++j) p.a[j] = (typename Elem<T>::Acc)(0.5f * ((float)p.a[j] + (float)o.a[j]));
            }
            if (a.do_step) {
#pragma unroll
                for (int j = 0; j < L; ++j) {
                    float w = (float)p.a[j];
                    opt_update<KIND>(w, g[j] * h.v[kHpGradScale], m1.f[j], m2.f[j], h);
                    p.a[j] = (typename Elem<T>::Acc)w;
                }
                if (a.s1) m1.store(a.s1 + idx);
                if (KIND == kOptAdam && a.s2) m2.store(a.s2 + idx);
            }
            if (a.master) {
#pragma unroll
                for (int j = 0; j < L; ++j) mw.f[j] = (float)p.a[j];
                mw.store(a.master + idx);
            }
            const Vec16 res = pack<T>(p);
            st_plain(P + v * 16, res);
            st_stream(pub + v * 16, res);
        }
    }
    __syncthreads();
    if (remote && threadIdx.x == 0 && vr != 0 && pair_torn(m, target, a.ps, vr) && m.status)
        *((volatile uint32_t*)&m.status[5]) = m.status[5] + 1u;
    pair_commit(m, a.ps.ver_off, next, a.done_counter);
}

__global__ void __launch_bounds__(512) publish_kernel(const __grid_constant__ MeshArgs m,
                                                      const __grid_constant__ PairArgs a) {
    __shared__ uint32_t sh[2];
    pair_versions(m, -1, a.ps.ver_off, true, sh);
    const uint32_t next = sh[1];
    char* pub = reinterpret_cast<char*>(m.data[m.rank]) + a.ps.pub_off[next % (uint32_t)a.ps.nbuf];
    const char* S = reinterpret_cast<const char*>(a.params);
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t v = tid; v < a.nvec; v += stride) st_stream(pub + v * 16, ld_plain(S + v * 16));
    pair_commit(m, a.ps.ver_off, next, a.done_counter);
}

__global__ void __launch_bounds__(512) pull_kernel(const __grid_constant__ MeshArgs m,
                                                   const __grid_constant__ PairArgs a) {
    if (a.retry && ld_relaxed_sys(a.ok + 2) == 0u) return; // the first attempt was clean
    __shared__ uint32_t sh[2];
    const int target = a.target_dev ? *a.target_dev : a.target;
    pair_versions(m, target, a.ps.ver_off, false, sh);
    const uint32_t vr = sh[0];
    if (blockIdx.x == 0 && threadIdx.x == 0) a.ok[0] = vr;
    if (vr == 0) return;
    const char* src = reinterpret_cast<const char*>(m.data[target]) + a.ps.pub_off[vr % (uint32_t)a.ps.nbuf];
    char* D = reinterpret_cast<char*>(a.dst);
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t v0 = tid; v0 < a.nvec; v0 += stride * 4) {
        Vec16 t[4];
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (v0 + (size_t)u * stride < a.nvec) t[u] = ld_stream(src + (v0 + (size_t)u * stride) * 16);
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (v0 + (size_t)u * stride < a.nvec) st_plain(D + (v0 + (size_t)u * stride) * 16, t[u]);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        // ok[2] of a retry: set by the first attempt, cleared below only if EVERY block of this attempt was clean — blocks
        // that see a torn copy again leave it set (and count), clean blocks never write it
        if (pair_torn(m, target, a.ps, vr)) {
            atomicAdd(&a.ok[1], 1u);
            if (!a.retry) atomicExch(&a.ok[2], 1u);
            else atomicExch(&a.ok[3], 1u);
        }
    }
}

// After a retry: ok[2] <- ok[3] (still torn?), ok[3] <- 0. One thread; keeps the flag protocol free of inter-block races.
__global__ void pull_settle_kernel(uint32_t* ok) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        ok[2] = ok[3];
        ok[3] = 0u;
    }
}

static void check_pub(const MeshArgs& m, const PubStore& ps) {
    if (ps.nbuf < 2 || ps.nbuf > 3) throw std::runtime_error("kungfu_b200: the published-model store has 2 or 3 buffers");
    for (int i = 0; i < ps.nbuf; ++i)
        if (ps.pub_off[i] % 16) throw std::runtime_error("kungfu_b200: publish buffers must be 16-byte aligned");
    if (ps.ver_off % 16) throw std::runtime_error("kungfu_b200: version words must be 16-byte aligned");
}

void launch_pair_avg(const MeshArgs& m, int target, const PubStore& ps, void* params, int dtype, size_t nbytes, const void* grads,
                     int grad_dtype, float* s1, float* s2, const float* hyper, int opt_kind, int do_step, uint32_t* done_counter,
                     const void* pulled, const uint32_t* pulled_ok, int blocks, int threads, cudaStream_t st, const int* target_dev,
                     float* master) {
    check_launch_shape(m, nbytes, blocks, threads);
    check_pub(m, ps);
    if (!target_dev && (target < 0 || target >= m.nranks)) throw std::runtime_error("kungfu_b200: bad pair-averaging target");
    if (do_step && !grads) throw std::runtime_error("kungfu_b200: fused pair-averaging step needs gradients");
    if (pulled && !pulled_ok) throw std::runtime_error("kungfu_b200: a private copy needs the status words of its pull");
    PairArgs a{};
    a.target = target;
    a.ps = ps;
    a.params = params;
    a.nvec = nbytes / 16;
    a.grads = grads;
    a.grad_dtype = grad_dtype;
    a.s1 = s1;
    a.s2 = s2;
    a.hyper = hyper;
    a.do_step = do_step;
    a.done_counter = done_counter;
    a.pulled = pulled;
    a.pulled_ok = pulled_ok;
    a.target_dev = target_dev;
    a.master = dtype == kF32 ? nullptr : master;
    if (dtype == kF32) {
        if (opt_kind == kOptSgd) pair_avg_kernel<float, kOptSgd><<<blocks, threads, 0, st>>>(m, a);
        else pair_avg_kernel<float, kOptAdam><<<blocks, threads, 0, st>>>(m, a);
    } else if (dtype == kBF16) {
        if (opt_kind == kOptSgd) pair_avg_kernel<__nv_bfloat16, kOptSgd><<<blocks, threads, 0, st>>>(m, a);
        else pair_avg_kernel<__nv_bfloat16, kOptAdam><<<blocks, threads, 0, st>>>(m, a);
    } else {
        throw std::runtime_error("kungfu_b200: pair averaging supports f32/bf16 parameters");
    }
    count_launch();
    KF_CUDA_CHECK(cudaGetLastError());
}

void launch_publish(const MeshArgs& m, const PubStore& ps, const void* src, size_t nbytes, uint32_t* done_counter, int blocks,
                    int threads, cudaStream_t st) {
    check_launch_shape(m, nbytes, blocks, threads);
    check_pub(m, ps);
    PairArgs a{};
    a.target = -1;
    a.ps = ps;
    a.params = const_cast<void*>(src);
    a.nvec = nbytes / 16;
    a.done_counter = done_counter;
    publish_kernel<<<blocks, threads, 0, st>>>(m, a);
    count_launch();
    KF_CUDA_CHECK(cudaGetLastError());
}

void launch_pull_settle(uint32_t* ok, cudaStream_t st) {
    pull_settle_kernel<<<1, 32, 0, st>>>(ok);
    count_launch();
    KF_CUDA_CHECK(cudaGetLastError());
}

void launch_pull(const MeshArgs& m, int target, const PubStore& ps, void* dst, size_t nbytes, uint32_t* ok, int retry, int blocks,
                 int threads, cudaStream_t st, const int* target_dev) {
    check_launch_shape(m, nbytes, blocks, threads);
    check_pub(m, ps);
    if (!target_dev && (target < 0 || target >= m.nranks)) throw std::runtime_error("kungfu_b200: bad pull target");
    PairArgs a{};
    a.target = target;
    a.ps = ps;
    a.nvec = nbytes / 16;
    a.dst = dst;
    a.ok = ok;
    a.retry = retry;
    a.target_dev = target_dev;
    pull_kernel<<<blocks, threads, 0, st>>>(m, a);
    count_launch();
    KF_CUDA_CHECK(cudaGetLastError());
    if (retry) launch_pull_settle(ok, st);
}

} // namespace kf
