// kungfu_b200 — fused BatchNorm (+ residual add) + ReLU for channels_last activations, forward and backward.
//
// ResNet-50 training at batch 256 on a B200 spends ~40 % of the step in memory-bound BatchNorm / ReLU / add kernels
// (profiles/resnet50_step_kernels.md). PyTorch runs them as separate passes: stats, normalise, relu, add in forward;
// relu-backward, reduce, elementwise in backward = 13–16 activation-sized passes per BN layer. These kernels do the same
// math in 8–10 passes: the normalise + add + ReLU of forward is one pass that also emits a 1-bit/element ReLU mask, and
// backward needs only (dy, x, mask) — the saved output is never re-read.
//
// Layout: x is [rows = N*H*W, C] with C contiguous (channels_last), C % 8 == 0. One thread owns one 8-channel vector
// (16 bytes of bf16) and strides over rows, so its per-channel coefficients live in registers. Cross-block reductions
// are two-stage (per-block partials in a scratch buffer, reduced by the tiny finalize kernels): deterministic and free
// of global atomics.
//
// The reference ships no model code; its benchmark uses the stock framework model (benchmarks/system/
// benchmark_kungfu_torch.py:58). This is part of the "model family" layer of the new framework (models/resnet.py BNAct).
#include "common.cuh"
#include "kernels.h"

namespace kf {

namespace {

constexpr int kBnThreads = 256;

template <typename T> struct Vec8 {
    float f[8];
    __device__ __forceinline__ void load(const T* p) {
        const Vec16 v = ld_stream(p);
        const AccVec<T> a = unpack<T>(v);
#pragma unroll
        for (int i = 0; i < 8; ++i) f[i] = a.a[i];
    }
    __device__ __forceinline__ void store(T* p) const {
        AccVec<T> a;
#pragma unroll
        for (int i = 0; i < 8; ++i) a.a[i] = f[i];
        st_stream(p, pack<T>(a));
    }
};

// ---- block-level fold of per-thread channel sums -------------------------------------------------------------------
// Every thread holds sums for one 8-channel vector `cv`; threads of a block that share a vector are folded without atomics
// (shared-memory float atomics are CAS loops: with C = 64 there are 32 contenders per address and the tail of the kernel
// costs more than the streaming part). C/8 is a power of two here (bn_check): lanes sharing a vector are first folded with
// xor-shuffles, the remaining sharers (one per warp or per C/8-thread group) meet in a [groups][2][C] shared array.
constexpr int kBnUnroll = 4; // independent 16-byte loads in flight per thread and per input tensor

__host__ __device__ inline int bn_fold_groups(int cvn) { return cvn < 32 ? kBnThreads / 32 : (cvn <= kBnThreads ? kBnThreads / cvn : 1); }

__device__ __forceinline__ void bn_block_fold(float (&s)[8], float (&q)[8], int cv, int cvn, int C, float* __restrict__ out, float* sm) {
    if (cvn > kBnThreads) { // C > 2048: a block covers a slice of the channels, no two threads share one
        for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) sm[i] = 0.f;
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            sm[cv * 8 + i] = s[i];
            sm[C + cv * 8 + i] = q[i];
        }
        __syncthreads();
        for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) out[i] = sm[i];
        return;
    }
    for (int off = cvn; off < 32; off <<= 1) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            s[i] += __shfl_xor_sync(0xffffffffu, s[i], off);
            q[i] += __shfl_xor_sync(0xffffffffu, q[i], off);
        }
    }
    const int lane = threadIdx.x & 31;
    const int g = cvn < 32 ? (int)(threadIdx.x >> 5) : (int)threadIdx.x / cvn;
    if (cvn >= 32 || lane < cvn) {
        float4* ps = reinterpret_cast<float4*>(sm + (size_t)g * 2 * C + cv * 8);
        float4* pq = reinterpret_cast<float4*>(sm + (size_t)g * 2 * C + C + cv * 8);
        ps[0] = make_float4(s[0], s[1], s[2], s[3]);
        ps[1] = make_float4(s[4], s[5], s[6], s[7]);
        pq[0] = make_float4(q[0], q[1], q[2], q[3]);
        pq[1] = make_float4(q[4], q[5], q[6], q[7]);
    }
    __syncthreads();
    const int groups = bn_fold_groups(cvn);
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) {
        float a = 0.f;
        for (int k = 0; k < groups; ++k) a += sm[(size_t)k * 2 * C + i];
        out[i] = a;
    }
}

// ---- forward: per-block partial sums ------------------------------------------------------------------------
// partial[block][0][c] = sum x, partial[block][1][c] = sum x^2 over the rows handled by the block
template <typename T>
__global__ void __launch_bounds__(kBnThreads) bn_stats_kernel(const T* __restrict__ x, size_t rows, int C, float* __restrict__ partial) {
    extern __shared__ __align__(16) float sm[]; // [groups][2][C]
    const int cvn = C / 8;
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t nthreads = (size_t)gridDim.x * blockDim.x;
    const int cv = (int)(tid % cvn);
    const size_t rstride = nthreads / cvn;
    const uint32_t zero = (uint32_t)(rows >> 62); // always 0, opaque to the compiler (see hold4)
    size_t r = tid / cvn;
    float s[8], q[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) s[i] = q[i] = 0.f;
    const T* px = x + (size_t)cv * 8;
    for (; r + (kBnUnroll - 1) * rstride < rows; r += kBnUnroll * rstride) { // kBnUnroll loads issued before the first use
        Vec16 raw[kBnUnroll];
#pragma unroll
        for (int u = 0; u < kBnUnroll; ++u) raw[u] = ld_ro_stream(px + (r + u * rstride) * C);
        static_assert(kBnUnroll == 4, "hold4 takes four vectors");
        hold4(raw[0], raw[1], raw[2], raw[3], zero);
#pragma unroll
        for (int u = 0; u < kBnUnroll; ++u) {
            const AccVec<T> v = unpack<T>(raw[u]);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                s[i] += v.a[i];
                q[i] = fmaf(v.a[i], v.a[i], q[i]);
            }
        }
    }
    for (; r < rows; r += rstride) {
        const AccVec<T> v = unpack<T>(ld_ro_stream(px + r * C));
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            s[i] += v.a[i];
            q[i] = fmaf(v.a[i], v.a[i], q[i]);
        }
    }
    bn_block_fold(s, q, cv, cvn, C, partial + (size_t)blockIdx.x * 2 * C, sm);
}

// Sum the per-block partials of 32 channels with 16 thread groups striding over the blocks (coalesced 128-byte rows of the
// partial matrix, 16-way parallel), then a shared-memory fold. A serial per-channel loop over ~600 partials costs ~80 us
// of pure load latency per BN layer; this shape brings it to a few microseconds.
__device__ __forceinline__ bool bn_fold_partials(const float* __restrict__ partial, int nblocks, int C, int c, float& s, float& q) {
    __shared__ float ss[16][33], sq[16][33];
    const int cl = threadIdx.x & 31, bg = threadIdx.x >> 5; // 32 channels x 16 partial groups
    float a0 = 0.f, a1 = 0.f;
    if (c < C) {
#pragma unroll 4
        for (int b = bg; b < nblocks; b += 16) {
            a0 += partial[(size_t)b * 2 * C + c];
            a1 += partial[(size_t)b * 2 * C + C + c];
        }
    }
    ss[bg][cl] = a0;
    sq[bg][cl] = a1;
    __syncthreads();
    if (bg != 0 || c >= C) return false;
    s = q = 0.f;
#pragma unroll
    for (int g = 0; g < 16; ++g) {
        s += ss[g][cl];
        q += sq[g][cl];
    }
    return true;
}

// reduce partials, produce mean / invstd / scale / shift, update the running statistics. grid = ceil(C / 32), block = 256
template <typename P>
__global__ void __launch_bounds__(512) bn_fwd_finalize_kernel(const float* __restrict__ partial, int nblocks, size_t rows, int C,
                                                              const P* __restrict__ gamma, const P* __restrict__ beta, float eps, float momentum,
                                                              float* __restrict__ mean_out, float* __restrict__ invstd_out,
                                                              float* __restrict__ scale_out, float* __restrict__ shift_out, P* running_mean,
                                                              P* running_var) {
    const int c = blockIdx.x * 32 + (threadIdx.x & 31);
    float s, q;
    if (!bn_fold_partials(partial, nblocks, C, c, s, q)) return;
    const float m = (float)rows;
    const float mean = s / m;
    float var = q / m - mean * mean;
    var = var < 0.f ? 0.f : var;
    const float invstd = rsqrtf(var + eps);
    const float g = gamma ? (float)Elem<P>::up(gamma[c]) : 1.f;
    const float b = beta ? (float)Elem<P>::up(beta[c]) : 0.f;
    mean_out[c] = mean;
    invstd_out[c] = invstd;
    scale_out[c] = g * invstd;
    shift_out[c] = b - mean * g * invstd;
    if (running_mean) {
        const float unbiased = m > 1.f ? var * m / (m - 1.f) : var;
        running_mean[c] = Elem<P>::down((1.f - momentum) * (float)Elem<P>::up(running_mean[c]) + momentum * mean);
        running_var[c] = Elem<P>::down((1.f - momentum) * (float)Elem<P>::up(running_var[c]) + momentum * unbiased);
    }
}

// y = relu(x * scale + shift (+ res)); mask byte per 8-channel vector (bit i = y_i > 0)
template <typename T, bool RES, bool RELU>
__global__ void __launch_bounds__(kBnThreads) bn_apply_kernel(const T* __restrict__ x, const T* __restrict__ res, T* __restrict__ y,
                                                              uint8_t* __restrict__ mask, size_t nvec, int C, const float* __restrict__ scale,
                                                              const float* __restrict__ shift) {
    const int cvn = C / 8;
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x; // multiple of cvn: the channel vector of a thread is fixed
    const int cv = (int)(tid % cvn);
    float sc[8], sh[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        sc[i] = __ldg(scale + cv * 8 + i);
        sh[i] = __ldg(shift + cv * 8 + i);
    }
    const uint32_t zero = (uint32_t)(nvec >> 62); // always 0, opaque to the compiler (see hold4)
    auto body = [&](const Vec16& rx, const Vec16& rr, size_t v) {
        AccVec<T> a = unpack<T>(rx);
        AccVec<T> r;
        if constexpr (RES) r = unpack<T>(rr);
        uint32_t m = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float z = fmaf(a.a[i], sc[i], sh[i]);
            if constexpr (RES) z += r.a[i];
            if constexpr (RELU) {
                m |= (z > 0.f ? 1u : 0u) << i;
                z = z > 0.f ? z : 0.f;
            }
            a.a[i] = z;
        }
        st_stream(y + v * 8, pack<T>(a));
        if constexpr (RELU) mask[v] = (uint8_t)m;
    };
    size_t v = tid;
    if constexpr (RES) {
        for (; v + stride < nvec; v += 2 * stride) { // two vectors (x + residual each) per trip: four loads before the first use
            Vec16 x0 = ld_ro_stream(x + v * 8), x1 = ld_ro_stream(x + (v + stride) * 8);
            Vec16 r0 = ld_ro_stream(res + v * 8), r1 = ld_ro_stream(res + (v + stride) * 8);
            hold4(x0, r0, x1, r1, zero);
            body(x0, r0, v);
            body(x1, r1, v + stride);
        }
    } else {
        for (; v + 3 * stride < nvec; v += 4 * stride) { // four vectors per trip
            Vec16 x0 = ld_ro_stream(x + v * 8), x1 = ld_ro_stream(x + (v + stride) * 8);
            Vec16 x2 = ld_ro_stream(x + (v + 2 * stride) * 8), x3 = ld_ro_stream(x + (v + 3 * stride) * 8);
            hold4(x0, x1, x2, x3, zero);
            const Vec16 none{};
            body(x0, none, v);
            body(x1, none, v + stride);
            body(x2, none, v + 2 * stride);
            body(x3, none, v + 3 * stride);
        }
    }
    for (; v < nvec; v += stride) {
        Vec16 r0{};
        if constexpr (RES) r0 = ld_ro_stream(res + v * 8);
        body(ld_ro_stream(x + v * 8), r0, v);
    }
}

// ---- backward ---------------------------------------------------------------------------------------------------
// partial[block][0][c] = sum dz, partial[block][1][c] = sum dz * xhat, dz = dy masked by the ReLU bit
template <typename T, bool RELU>
__global__ void __launch_bounds__(kBnThreads) bn_bwd_reduce_kernel(const T* __restrict__ dy, const T* __restrict__ x,
                                                                   const uint8_t* __restrict__ mask, size_t rows, int C,
                                                                   const float* __restrict__ mean, const float* __restrict__ invstd,
                                                                   float* __restrict__ partial) {
    extern __shared__ __align__(16) float sm[]; // [groups][2][C]
    const int cvn = C / 8;
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t nthreads = (size_t)gridDim.x * blockDim.x;
    const int cv = (int)(tid % cvn);
    const size_t rstride = nthreads / cvn;
    const uint32_t zero = (uint32_t)(rows >> 62); // always 0, opaque to the compiler (see hold4)
    size_t r = tid / cvn;
    float mu[8], is[8], s[8], q[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        mu[i] = __ldg(mean + cv * 8 + i);
        is[i] = __ldg(invstd + cv * 8 + i);
        s[i] = q[i] = 0.f;
    }
    auto accumulate = [&](const Vec16& rg, const Vec16& ra, uint32_t m) {
        const AccVec<T> g = unpack<T>(rg), a = unpack<T>(ra);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float dz = (m >> i) & 1u ? g.a[i] : 0.f;
            s[i] += dz;
            q[i] = fmaf(dz, (a.a[i] - mu[i]) * is[i], q[i]);
        }
    };
    constexpr int U = kBnUnroll / 2; // two tensors per row: 2 * U = kBnUnroll loads in flight
    for (; r + (U - 1) * rstride < rows; r += U * rstride) {
        Vec16 rg[U], ra[U];
        uint32_t m[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const size_t v = (r + u * rstride) * cvn + cv;
            rg[u] = ld_ro_stream(dy + v * 8);
            ra[u] = ld_ro_stream(x + v * 8);
            m[u] = 0xffu;
            if constexpr (RELU) m[u] = mask[v];
        }
        static_assert(U == 2, "hold4 takes four vectors");
        hold4(rg[0], ra[0], rg[1], ra[1], zero);
#pragma unroll
        for (int u = 0; u < U; ++u) accumulate(rg[u], ra[u], m[u]);
    }
    for (; r < rows; r += rstride) {
        const size_t v = r * cvn + cv;
        uint32_t m = 0xffu;
        if constexpr (RELU) m = mask[v];
        accumulate(ld_ro_stream(dy + v * 8), ld_ro_stream(x + v * 8), m);
    }
    bn_block_fold(s, q, cv, cvn, C, partial + (size_t)blockIdx.x * 2 * C, sm);
}

// dgamma / dbeta (parameter dtype) and the per-channel coefficients of the elementwise pass:
//   dx = a * (dz - b - xhat * c),  a = gamma * invstd, b = sum(dz) / M, c = sum(dz * xhat) / M
template <typename P>
__global__ void __launch_bounds__(512) bn_bwd_finalize_kernel(const float* __restrict__ partial, int nblocks, size_t rows, int C,
                                                              const P* __restrict__ gamma, const float* __restrict__ invstd,
                                                              P* __restrict__ dgamma, P* __restrict__ dbeta, float* __restrict__ coef /*[3][C]*/) {
    const int c = blockIdx.x * 32 + (threadIdx.x & 31);
    float s, q;
    if (!bn_fold_partials(partial, nblocks, C, c, s, q)) return;
    if (dgamma) dgamma[c] = Elem<P>::down(q);
    if (dbeta) dbeta[c] = Elem<P>::down(s);
    const float g = gamma ? (float)Elem<P>::up(gamma[c]) : 1.f;
    const float m = (float)rows;
    coef[c] = g * invstd[c];
    coef[C + c] = s / m;
    coef[2 * C + c] = q / m;
}

template <typename T, bool RES, bool RELU>
__global__ void __launch_bounds__(kBnThreads) bn_bwd_apply_kernel(const T* __restrict__ dy, const T* __restrict__ x,
                                                                  const uint8_t* __restrict__ mask, T* __restrict__ dx, T* __restrict__ dres,
                                                                  size_t nvec, int C, const float* __restrict__ mean,
                                                                  const float* __restrict__ invstd, const float* __restrict__ coef) {
    const int cvn = C / 8;
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const int cv = (int)(tid % cvn);
    // dx = ca * (dz - cb - (x - mu) * is * cc)  folded to  dx = ca * dz + k1 * x + k0  (24 coefficient registers instead of 40)
    float ca[8], k1[8], k0[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int c = cv * 8 + i;
        const float mu = __ldg(mean + c), is = __ldg(invstd + c), cb = __ldg(coef + C + c), cc = __ldg(coef + 2 * C + c);
        ca[i] = __ldg(coef + c);
        k1[i] = -ca[i] * cc * is;
        k0[i] = -ca[i] * cb - k1[i] * mu;
    }
    const uint32_t zero = (uint32_t)(nvec >> 62); // always 0, opaque to the compiler (see hold4)
    auto body = [&](const Vec16& rg, const Vec16& rx, uint32_t m, size_t v) {
        AccVec<T> g = unpack<T>(rg), a = unpack<T>(rx);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float z = (m >> i) & 1u ? g.a[i] : 0.f;
            g.a[i] = z;
            a.a[i] = fmaf(ca[i], z, fmaf(k1[i], a.a[i], k0[i]));
        }
        st_stream(dx + v * 8, pack<T>(a));
        if constexpr (RES) st_stream(dres + v * 8, pack<T>(g));
    };
    size_t v = tid;
    for (; v + stride < nvec; v += 2 * stride) {
        Vec16 g0 = ld_ro_stream(dy + v * 8), x0 = ld_ro_stream(x + v * 8);
        Vec16 g1 = ld_ro_stream(dy + (v + stride) * 8), x1 = ld_ro_stream(x + (v + stride) * 8);
        uint32_t m0 = 0xffu, m1 = 0xffu;
        if constexpr (RELU) m0 = mask[v], m1 = mask[v + stride];
        hold4(g0, x0, g1, x1, zero);
        body(g0, x0, m0, v);
        body(g1, x1, m1, v + stride);
    }
    for (; v < nvec; v += stride) {
        uint32_t m = 0xffu;
        if constexpr (RELU) m = mask[v];
        body(ld_ro_stream(dy + v * 8), ld_ro_stream(x + v * 8), m, v);
    }
}

// Element-wise passes want full occupancy (8 x 256 threads per SM keep enough loads in flight for HBM); the reducing
// passes pay a per-block epilogue (shared-memory fold + 2C partials), so they use fewer, longer-running blocks.
int bn_grid(size_t nvec, int C, bool reducing = false) {
    // every thread must keep a fixed channel vector: total threads is a multiple of C/8 (C/8 divides 256 or equals it)
    const size_t per_block = (size_t)kBnThreads * (reducing ? 16 : 4);
    size_t blocks = (nvec + per_block - 1) / per_block;
    if (blocks < 1) blocks = 1;
    const size_t cap = reducing ? 148 * 4 : 148 * 8;
    if (blocks > cap) blocks = cap;
    if (C / 8 > kBnThreads) {
        const size_t k = (C / 8) / kBnThreads; // blocks must come in groups of k
        blocks = (blocks + k - 1) / k * k;
    }
    return (int)blocks;
}

void bn_check(size_t rows, int C) {
    if (C <= 0 || C % 8 != 0) throw std::runtime_error("kungfu_b200 bn_act: C must be a positive multiple of 8");
    const int cvn = C / 8;
    if (!((cvn <= kBnThreads && kBnThreads % cvn == 0) || (cvn > kBnThreads && cvn % kBnThreads == 0)))
        throw std::runtime_error("kungfu_b200 bn_act: C/8 must divide or be a multiple of 256");
    if (2 * (size_t)C * sizeof(float) > 96 * 1024) throw std::runtime_error("kungfu_b200 bn_act: C too large");
    (void)rows;
}

} // namespace

int bn_num_blocks(size_t rows, int C) {
    bn_check(rows, C);
    return bn_grid(rows * (size_t)(C / 8), C, true);
}

#define KF_BN_T(dtype, CALL)                                                                   \
    do {                                                                                       \
        if (dtype == kBF16) { using T = __nv_bfloat16; CALL; }                                 \
        else if (dtype == kF16) { using T = __half; CALL; }                                    \
        else throw std::runtime_error("kungfu_b200 bn_act: activations must be bf16 or fp16"); \
    } while (0)

#define KF_BN_P(pdtype, CALL)                                                                           \
    do {                                                                                                \
        if (pdtype == kBF16) { using P = __nv_bfloat16; CALL; }                                         \
        else if (pdtype == kF16) { using P = __half; CALL; }                                            \
        else if (pdtype == kF32) { using P = float; CALL; }                                             \
        else throw std::runtime_error("kungfu_b200 bn_act: parameters must be bf16, fp16 or fp32");     \
    } while (0)

void launch_bn_fwd(const void* x, const void* res, void* y, uint8_t* mask, int dtype, size_t rows, int C, const void* gamma, const void* beta,
                   void* running_mean, void* running_var, int pdtype, float eps, float momentum, int relu, float* partial, float* mean,
                   float* invstd, float* scale, float* shift, cudaStream_t st) {
    bn_check(rows, C);
    const size_t nvec = rows * (size_t)(C / 8);
    const int blocks = bn_grid(nvec, C, true);
    const size_t smem = (size_t)bn_fold_groups(C / 8) * 2 * (size_t)C * sizeof(float);
    KF_BN_T(dtype, (bn_stats_kernel<T><<<blocks, kBnThreads, smem, st>>>((const T*)x, rows, C, partial)));
    KF_BN_P(pdtype, (bn_fwd_finalize_kernel<P><<<(C + 31) / 32, 512, 0, st>>>(partial, blocks, rows, C, (const P*)gamma, (const P*)beta, eps,
                                                                                momentum, mean, invstd, scale, shift, (P*)running_mean,
                                                                                (P*)running_var)));
    launch_bn_apply(x, res, y, mask, dtype, rows, C, scale, shift, relu, st);
    count_launch(2);
    KF_CUDA_CHECK(cudaGetLastError());
}

void launch_bn_apply(const void* x, const void* res, void* y, uint8_t* mask, int dtype, size_t rows, int C, const float* scale, const float* shift,
                     int relu, cudaStream_t st) {
    bn_check(rows, C);
    const size_t nvec = rows * (size_t)(C / 8);
    const int blocks = bn_grid(nvec, C);
    if (relu && !mask) throw std::runtime_error("kungfu_b200 bn_act: relu needs a mask buffer");
#define KF_BN_APPLY(RES, RELU) \
    KF_BN_T(dtype, (bn_apply_kernel<T, RES, RELU><<<blocks, kBnThreads, 0, st>>>((const T*)x, (const T*)res, (T*)y, mask, nvec, C, scale, shift)))
    if (res && relu) KF_BN_APPLY(true, true);
    else if (res) KF_BN_APPLY(true, false);
    else if (relu) KF_BN_APPLY(false, true);
    else KF_BN_APPLY(false, false);
#undef KF_BN_APPLY
    count_launch();
    KF_CUDA_CHECK(cudaGetLastError());
}

void launch_bn_bwd(const void* dy, const void* x, const uint8_t* mask, void* dx, void* dres, int dtype, size_t rows, int C, const void* gamma,
                   void* dgamma, void* dbeta, int pdtype, const float* mean, const float* invstd, int relu, float* partial, float* coef,
                   cudaStream_t st) {
    bn_check(rows, C);
    const size_t nvec = rows * (size_t)(C / 8);
    const int rblocks = bn_grid(nvec, C, true);
    const int blocks = bn_grid(nvec, C);
    const size_t smem = (size_t)bn_fold_groups(C / 8) * 2 * (size_t)C * sizeof(float);
    if (relu && !mask) throw std::runtime_error("kungfu_b200 bn_act: relu backward needs the mask");
    if (relu) KF_BN_T(dtype, (bn_bwd_reduce_kernel<T, true><<<rblocks, kBnThreads, smem, st>>>((const T*)dy, (const T*)x, mask, rows, C, mean, invstd, partial)));
    else KF_BN_T(dtype, (bn_bwd_reduce_kernel<T, false><<<rblocks, kBnThreads, smem, st>>>((const T*)dy, (const T*)x, mask, rows, C, mean, invstd, partial)));
    KF_BN_P(pdtype, (bn_bwd_finalize_kernel<P><<<(C + 31) / 32, 512, 0, st>>>(partial, rblocks, rows, C, (const P*)gamma, invstd, (P*)dgamma,
                                                                                (P*)dbeta, coef)));
#define KF_BN_BWD(RES, RELU)                                                                                                              \
    KF_BN_T(dtype, (bn_bwd_apply_kernel<T, RES, RELU><<<blocks, kBnThreads, 0, st>>>((const T*)dy, (const T*)x, mask, (T*)dx, (T*)dres, nvec, C, \
                                                                                       mean, invstd, coef)))
    if (dres && relu) KF_BN_BWD(true, true);
    else if (dres) KF_BN_BWD(true, false);
    else if (relu) KF_BN_BWD(false, true);
    else KF_BN_BWD(false, false);
#undef KF_BN_BWD
    count_launch(3);
    KF_CUDA_CHECK(cudaGetLastError());
}

} // namespace kf
