// kungfu_b200 — single-GPU building blocks on the flat arenas: optimizer over a flat parameter
// arena (one launch for the whole model instead of one per tensor), casts, squared norms, the
// gradient-noise-scale finalisation, and the input normalisation kernel used by the e2e path.
//
// Reference counterparts: TF stock element-wise kernels used as "fused ops" by the optimizers
// (SURVEY.md §2.3 K7–K11), `KungfuNoiseScale` (srcs/cpp/src/tensorflow/ops/cpu/collective.cpp:212-258),
// `ExponentialMovingAverage` (srcs/cpp/include/kungfu/utils/ema.hpp:19-28).
#include "kernels.h"
#include "optim.cuh"

namespace kf {

static inline int grid_for(size_t work_items, int threads, int max_blocks = 148 * 8) {
    size_t b = (work_items + threads - 1) / threads;
    if (b < 1) b = 1;
    if (b > (size_t)max_blocks) b = max_blocks;
    return (int)b;
}

__device__ __forceinline__ float load_as_float(const void* p, int dtype, size_t i) {
    if (dtype == kF32) return reinterpret_cast<const float*>(p)[i];
    if (dtype == kBF16) return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[i]);
    if (dtype == kF16) return __half2float(reinterpret_cast<const __half*>(p)[i]);
    return 0.f;
}
__device__ __forceinline__ void store_from_float(void* p, int dtype, size_t i, float v) {
    if (dtype == kF32) reinterpret_cast<float*>(p)[i] = v;
    else if (dtype == kBF16) reinterpret_cast<__nv_bfloat16*>(p)[i] = __float2bfloat16_rn(v);
    else if (dtype == kF16) reinterpret_cast<__half*>(p)[i] = __float2half_rn(v);
}

// 4 elements per thread, vectorised for the three float dtypes.
__device__ __forceinline__ void load4(const void* p, int dtype, size_t i, float* o) {
    if (dtype == kF32) {
        const float4 q = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p) + i);
        o[0] = q.x; o[1] = q.y; o[2] = q.z; o[3] = q.w;
    } else {
        const uint2 q = *reinterpret_cast<const uint2*>(reinterpret_cast<const uint16_t*>(p) + i);
        if (dtype == kBF16) {
            const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&q.x);
            const __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162*>(&q.y);
            o[0] = __low2float(a); o[1] = __high2float(a); o[2] = __low2float(b); o[3] = __high2float(b);
        } else {
            const __half2 a = *reinterpret_cast<const __half2*>(&q.x);
            const __half2 b = *reinterpret_cast<const __half2*>(&q.y);
            o[0] = __low2float(a); o[1] = __high2float(a); o[2] = __low2float(b); o[3] = __high2float(b);
        }
    }
}
__device__ __forceinline__ void store4(void* p, int dtype, size_t i, const float* v) {
    if (dtype == kF32) {
        *reinterpret_cast<float4*>(reinterpret_cast<float*>(p) + i) = make_float4(v[0], v[1], v[2], v[3]);
    } else {
        uint2 q;
        if (dtype == kBF16) {
            const __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]);
            const __nv_bfloat162 b = __floats2bfloat162_rn(v[2], v[3]);
            q.x = *reinterpret_cast<const uint32_t*>(&a);
            q.y = *reinterpret_cast<const uint32_t*>(&b);
        } else {
            const __half2 a = __floats2half2_rn(v[0], v[1]);
            const __half2 b = __floats2half2_rn(v[2], v[3]);
            q.x = *reinterpret_cast<const uint32_t*>(&a);
            q.y = *reinterpret_cast<const uint32_t*>(&b);
        }
        *reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(p) + i) = q;
    }
}

// ---- hyper tick: advance the step counter / bias corrections on device (graph-capturable) -----------
__global__ void hyper_tick_kernel(float* hp) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        const float step = hp[kHpStep] + 1.f;
        hp[kHpStep] = step;
        hp[kHpFirstStep] = step == 1.f ? 1.f : 0.f;
        hp[kHpBiasC1] = 1.f - powf(hp[kHpBeta1], step);
        hp[kHpBiasC2] = 1.f - powf(hp[kHpBeta2], step);
    }
}
void launch_hyper_tick(float* hyper, cudaStream_t st) {
    hyper_tick_kernel<<<1, 32, 0, st>>>(hyper);
    count_launch();
    KF_CUDA_CHECK(cudaGetLastError());
}

// ---- flat optimizer over the whole arena ------------------------------------------------------------
template <int KIND>
__global__ void __launch_bounds__(256) flat_opt_kernel(void* params, int dtype, const void* grads, int grad_dtype, float* master,
                                                       float* s1, float* s2, const float* hyper, size_t nelem) {
    const Hyper h = load_hyper(hyper);
    const float gs = h.v[kHpGradScale];
    const size_t n4 = nelem / 4;
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t q = tid; q < n4; q += stride) {
        const size_t i = q * 4;
        float w[4], g[4], a[4] = {0.f, 0.f, 0.f, 0.f}, b[4] = {0.f, 0.f, 0.f, 0.f};
        if (master) load4(master, kF32, i, w);
        else load4(params, dtype, i, w);
        load4(grads, grad_dtype, i, g);
        if (s1) load4(s1, kF32, i, a);
        if (KIND == kOptAdam && s2) load4(s2, kF32, i, b);
#pragma unroll
        for (int j = 0; j < 4; ++j) opt_update<KIND>(w[j], g[j] * gs, a[j], b[j], h);
        if (master) store4(master, kF32, i, w);
        store4(params, dtype, i, w);
        if (s1) store4(s1, kF32, i, a);
        if (KIND == kOptAdam && s2) store4(s2, kF32, i, b);
    }
    // tail (nelem % 4)
    for (size_t i = n4 * 4 + tid; i < nelem; i += stride) {
        float w = master ? master[i] : load_as_float(params, dtype, i);
        float a = s1 ? s1[i] : 0.f, b = (KIND == kOptAdam && s2) ? s2[i] : 0.f;
        opt_update<KIND>(w, load_as_float(grads, grad_dtype, i) * gs, a, b, h);
        if (master) master[i] = w;
        store_from_float(params, dtype, i, w);
        if (s1) s1[i] = a;
        if (KIND == kOptAdam && s2) s2[i] = b;
    }
}

void launch_flat_opt(void* params, int dtype, const void* grads, int grad_dtype, float* master, float* s1, float* s2,
                     const float* hyper, int opt_kind, size_t nelem, cudaStream_t st) {
    if (nelem == 0) return;
    const int threads = 256;
    const int blocks = grid_for((nelem + 3) / 4, threads);
    if (opt_kind == kOptSgd)
        flat_opt_kernel<kOptSgd><<<blocks, threads, 0, st>>>(params, dtype, grads, grad_dtype, master, s1, s2, hyper, nelem);
    else
        flat_opt_kernel<kOptAdam><<<blocks, threads, 0, st>>>(params, dtype, grads, grad_dtype, master, s1, s2, hyper, nelem);
    count_launch();
    KF_CUDA_CHECK(cudaGetLastError());
}

// ---- cast / scale -----------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) cast_kernel(const void* src, int sdt, void* dst, int ddt, size_t nelem, float scale) {
    const size_t n4 = nelem / 4;
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t q = tid; q < n4; q += stride) {
        float v[4];
        load4(src, sdt, q * 4, v);
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] *= scale;
        store4(dst, ddt, q * 4, v);
    }
    for (size_t i = n4 * 4 + tid; i < nelem; i += stride) store_from_float(dst, ddt, i, load_as_float(src, sdt, i) * scale);
}
void launch_cast(const void* src, int src_dtype, void* dst, int dst_dtype, size_t nelem, float scale, cudaStream_t st) {
    if (nelem == 0) return;
    cast_kernel<<<grid_for((nelem + 3) / 4, 256), 256, 0, st>>>(src, src_dtype, dst, dst_dtype, nelem, scale);
    count_launch();
    KF_CUDA_CHECK(cudaGetLastError());
}

__global__ void __launch_bounds__(256) axpby_kernel(void* y, const void* x, int dt, size_t nelem, float a, float b) {
    const size_t n4 = nelem / 4;
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t q = tid; q < n4; q += stride) {
        float xv[4], yv[4];
        load4(x, dt, q * 4, xv);
        load4(y, dt, q * 4, yv);
#pragma unroll
        for (int j = 0; j < 4; ++j) yv[j] = fmaf(a, xv[j], b * yv[j]);
        store4(y, dt, q * 4, yv);
    }
    for (size_t i = n4 * 4 + tid; i < nelem; i += stride)
        store_from_float(y, dt, i, fmaf(a, load_as_float(x, dt, i), b * load_as_float(y, dt, i)));
}
void launch_axpby(void* y, const void* x, int dtype, size_t nelem, float a, float b, cudaStream_t st) {
    if (nelem == 0) return;
    axpby_kernel<<<grid_for((nelem + 3) / 4, 256), 256, 0, st>>>(y, x, dtype, nelem, a, b);
    count_launch();
    KF_CUDA_CHECK(cudaGetLastError());
}

// ---- squared norms (gns_reduce: stage 1 per-block partials, stage 2 atomics into the accumulator) ---
template <bool DIFF>
__global__ void __launch_bounds__(512) sumsq_kernel(const void* a, const void* b, int dt, size_t nelem, float* out) {
    __shared__ float red[32];
    const size_t n4 = nelem / 4;
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    float s0 = 0.f, s1 = 0.f; // two accumulators for ILP
    for (size_t q = tid; q < n4; q += stride) {
        float v[4];
        load4(a, dt, q * 4, v);
        if constexpr (DIFF) {
            float w[4];
            load4(b, dt, q * 4, w);
#pragma unroll
            for (int j = 0; j < 4; ++j) v[j] -= w[j];
        }
        s0 = fmaf(v[0], v[0], s0);
        s1 = fmaf(v[1], v[1], s1);
        s0 = fmaf(v[2], v[2], s0);
        s1 = fmaf(v[3], v[3], s1);
    }
    for (size_t i = n4 * 4 + tid; i < nelem; i += stride) {
        float v = load_as_float(a, dt, i);
        if constexpr (DIFF) v -= load_as_float(b, dt, i);
        s0 = fmaf(v, v, s0);
    }
    const float t = block_sum(s0 + s1, red);
    if (threadIdx.x == 0) atomicAdd(out, t);
}
void launch_sumsq(const void* src, int dtype, size_t nelem, float* out, cudaStream_t st) {
    if (nelem == 0) return;
    sumsq_kernel<false><<<grid_for((nelem + 3) / 4, 512, 148 * 2), 512, 0, st>>>(src, nullptr, dtype, nelem, out);
    count_launch();
    KF_CUDA_CHECK(cudaGetLastError());
}
void launch_sumsq_diff(const void* a, const void* b, int dtype, size_t nelem, float* out, cudaStream_t st) {
    if (nelem == 0) return;
    sumsq_kernel<true><<<grid_for((nelem + 3) / 4, 512, 148 * 2), 512, 0, st>>>(a, b, dtype, nelem, out);
    count_launch();
    KF_CUDA_CHECK(cudaGetLastError());
}

// ---- gradient noise scale: estimators + EMAs + ratio, all on device ---------------------------------
// G_biased = (B_big |g_big|^2 - B_small |g_small|^2) / (B_big - B_small)
// S_biased = (|g_small|^2 - |g_big|^2) / (1/B_small - 1/B_big)
// EMA: first sample passes through, then v = a v + (1 - a) x  (utils/ema.hpp:19-28)
__global__ void gns_finalize_kernel(const float* sums, float* state, float* out, float b_small, float b_big, float alpha,
                                    float nranks) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const float g_small = sums[0] / nranks; // mean over ranks of |g_r|^2
    const float g_big = sums[1];
    float G = 0.f, S = 0.f;
    if (b_big != b_small) {
        G = (b_big * g_big - b_small * g_small) / (b_big - b_small);
        S = (g_small - g_big) / (1.f / b_small - 1.f / b_big);
    }
    float ge, se;
    if (state[2] == 0.f) {
        ge = G;
        se = S;
        state[2] = 1.f;
    } else {
        ge = alpha * state[0] + (1.f - alpha) * G;
        se = alpha * state[1] + (1.f - alpha) * S;
    }
    state[0] = ge;
    state[1] = se;
    out[0] = se / ge;
    out[1] = G;
    out[2] = S;
}
void launch_gns_finalize(const float* sums, float* state, float* out, float batch_small, float batch_big, float alpha,
                         float nranks, cudaStream_t st) {
    gns_finalize_kernel<<<1, 32, 0, st>>>(sums, state, out, batch_small, batch_big, alpha, nranks);
    count_launch();
    KF_CUDA_CHECK(cudaGetLastError());
}

// ---- uint8 NHWC -> normalised bf16/fp32 NHWC ---------------------------------------------------------
// 16 bytes of pixels (5.33 RGB pixels) per thread would straddle channels; process 4 pixels = 12 bytes
// per thread with three 4-byte loads, writing 12 outputs.
__global__ void __launch_bounds__(256) u8_normalize_kernel(const uint8_t* src, void* dst, int ddt, size_t npix,
                                                           const float* mean3, const float* invstd3) {
    const float m0 = mean3[0], m1 = mean3[1], m2 = mean3[2];
    const float i0 = invstd3[0], i1 = invstd3[1], i2 = invstd3[2];
    const size_t ngrp = npix / 4;
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t q = tid; q < ngrp; q += stride) {
        const uint32_t* s = reinterpret_cast<const uint32_t*>(src + q * 12);
        const uint32_t w0 = s[0], w1 = s[1], w2 = s[2];
        uint8_t b[12];
        *reinterpret_cast<uint32_t*>(b + 0) = w0;
        *reinterpret_cast<uint32_t*>(b + 4) = w1;
        *reinterpret_cast<uint32_t*>(b + 8) = w2;
        float v[12];
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            v[3 * p + 0] = ((float)b[3 * p + 0] - m0) * i0;
            v[3 * p + 1] = ((float)b[3 * p + 1] - m1) * i1;
            v[3 * p + 2] = ((float)b[3 * p + 2] - m2) * i2;
        }
        store4(dst, ddt, q * 12 + 0, v + 0);
        store4(dst, ddt, q * 12 + 4, v + 4);
        store4(dst, ddt, q * 12 + 8, v + 8);
    }
    for (size_t p = ngrp * 4 + tid; p < npix; p += stride) {
        store_from_float(dst, ddt, p * 3 + 0, ((float)src[p * 3 + 0] - m0) * i0);
        store_from_float(dst, ddt, p * 3 + 1, ((float)src[p * 3 + 1] - m1) * i1);
        store_from_float(dst, ddt, p * 3 + 2, ((float)src[p * 3 + 2] - m2) * i2);
    }
}
void launch_u8_normalize(const uint8_t* src, void* dst, int dst_dtype, size_t npix, const float* mean3,
                         const float* invstd3, cudaStream_t st) {
    if (npix == 0) return;
    u8_normalize_kernel<<<grid_for((npix + 3) / 4, 256), 256, 0, st>>>(src, dst, dst_dtype, npix, mean3, invstd3);
    count_launch();
    KF_CUDA_CHECK(cudaGetLastError());
}

__global__ void fill_u32_kernel(uint32_t* p, uint32_t v, size_t n) {
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = tid; i < n; i += stride) p[i] = v;
}
void launch_fill_u32(uint32_t* p, uint32_t v, size_t n, cudaStream_t st) {
    if (n == 0) return;
    fill_u32_kernel<<<grid_for(n, 256), 256, 0, st>>>(p, v, n);
    count_launch();
    KF_CUDA_CHECK(cudaGetLastError());
}

} // namespace kf

// ---- multi-tensor copy: the stolen gradients of one bucket -> their slots of the flat arena, ONE launch ----------------
// (reference: gradient `fuse` = tf.concat of every gradient, srcs/python/kungfu/tensorflow/ops/__init__.py:29-46)
namespace kf {

struct CopyBatch {
    const void* src[kMaxCopyBatch];
    void* dst[kMaxCopyBatch];
    unsigned long long nbytes[kMaxCopyBatch];
    int n;
    float* sumsq; // optional: [0] += sum of squares of the copied elements
    int dtype;
};

__device__ __forceinline__ float vec16_sumsq(const Vec16& v, int dtype) {
    float s = 0.f;
    if (dtype == kF32) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float f = __uint_as_float(v.w[i]);
            s = fmaf(f, f, s);
        }
    } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float a, c;
            if (dtype == kBF16) {
                const __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162*>(&v.w[i]);
                a = __low2float(h), c = __high2float(h);
            } else {
                const __half2 h = *reinterpret_cast<const __half2*>(&v.w[i]);
                a = __low2float(h), c = __high2float(h);
            }
            s = fmaf(a, a, fmaf(c, c, s));
        }
    }
    return s;
}

__global__ void __launch_bounds__(256) multi_copy_kernel(const __grid_constant__ CopyBatch b) {
    // blockIdx.y = tensor, blockIdx.x strides over its 16-byte vectors; ragged heads/tails go byte-wise
    const int t = blockIdx.y;
    if (t >= b.n) return;
    const char* s = reinterpret_cast<const char*>(b.src[t]);
    char* d = reinterpret_cast<char*>(b.dst[t]);
    const size_t nb = b.nbytes[t];
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    float sq = 0.f;
    if ((((uintptr_t)s | (uintptr_t)d) & 15) == 0) {
        const size_t nv = nb / 16;
        size_t v = tid;
        // four loads in flight per thread: with one, an 8 MiB weight gradient (BERT-large: most buckets are a single tensor) took 13 us
        // = 1.3 TB/s, pure load latency
        for (; v + 3 * stride < nv; v += 4 * stride) {
            const Vec16 x0 = ld_ro_stream(s + v * 16);
            const Vec16 x1 = ld_ro_stream(s + (v + stride) * 16);
            const Vec16 x2 = ld_ro_stream(s + (v + 2 * stride) * 16);
            const Vec16 x3 = ld_ro_stream(s + (v + 3 * stride) * 16);
            if (b.sumsq) sq += vec16_sumsq(x0, b.dtype) + vec16_sumsq(x1, b.dtype) + vec16_sumsq(x2, b.dtype) + vec16_sumsq(x3, b.dtype);
            st_plain(d + v * 16, x0);
            st_plain(d + (v + stride) * 16, x1);
            st_plain(d + (v + 2 * stride) * 16, x2);
            st_plain(d + (v + 3 * stride) * 16, x3);
        }
        for (; v < nv; v += stride) {
            const Vec16 x = ld_ro_stream(s + v * 16);
            if (b.sumsq) sq += vec16_sumsq(x, b.dtype);
            st_plain(d + v * 16, x);
        }
        const size_t esz = b.dtype == kF32 ? 4 : 2;
        for (size_t i = nv * 16 + tid * esz; i + esz <= nb; i += stride * esz) { // ragged tail, element-wise
            float f;
            if (esz == 4) {
                f = *reinterpret_cast<const float*>(s + i);
                *reinterpret_cast<float*>(d + i) = f;
            } else {
                const uint16_t h = *reinterpret_cast<const uint16_t*>(s + i);
                *reinterpret_cast<uint16_t*>(d + i) = h;
                f = b.dtype == kBF16 ? __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(&h)) : __half2float(*reinterpret_cast<const __half*>(&h));
            }
            sq = fmaf(f, f, sq);
        }
    } else {
        for (size_t i = tid; i < nb; i += stride) d[i] = s[i];
        if (b.sumsq && blockIdx.x == 0 && threadIdx.x == 0) { // unaligned hand-over (never seen from autograd): correct, not fast
            const size_t esz = b.dtype == kF32 ? 4 : 2;
            for (size_t i = 0; i + esz <= nb; i += esz) {
                float f;
                if (esz == 4) memcpy(&f, s + i, 4);
                else {
                    uint16_t h;
                    memcpy(&h, s + i, 2);
                    f = b.dtype == kBF16 ? __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(&h)) : __half2float(*reinterpret_cast<const __half*>(&h));
                }
                sq = fmaf(f, f, sq);
            }
        }
    }
    if (b.sumsq) {
        __shared__ float red[32];
        const float t0 = block_sum(sq, red);
        if (threadIdx.x == 0 && t0 != 0.f) atomicAdd(b.sumsq, t0);
    }
}

void launch_multi_copy(const void* const* src, void* const* dst, const size_t* nbytes, int n, cudaStream_t st, float* sumsq, int dtype) {
    if (sumsq && dtype != kF32 && dtype != kBF16 && dtype != kF16) throw std::runtime_error("kungfu_b200: multi_copy sum of squares needs a float dtype");
    for (int base = 0; base < n; base += kMaxCopyBatch) {
        CopyBatch b{};
        b.sumsq = sumsq;
        b.dtype = dtype;
        b.n = n - base < kMaxCopyBatch ? n - base : kMaxCopyBatch;
        size_t big = 0;
        for (int i = 0; i < b.n; ++i) {
            b.src[i] = src[base + i];
            b.dst[i] = dst[base + i];
            b.nbytes[i] = nbytes[base + i];
            if (nbytes[base + i] > big) big = nbytes[base + i];
        }
        // one CTA per 16 KiB of the largest tensor (256 threads x 4 vectors x 16 bytes), at most two waves of the GPU per tensor row
        int gx = (int)((big / 16 + 1023) / 1024);
        if (gx < 1) gx = 1;
        if (gx > 296) gx = 296;
        multi_copy_kernel<<<dim3(gx, b.n), 256, 0, st>>>(b);
        count_launch();
    }
    KF_CUDA_CHECK(cudaGetLastError());
}

} // namespace kf
