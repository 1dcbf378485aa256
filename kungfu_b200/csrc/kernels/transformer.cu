// kungfu_b200 — fused transformer element-wise kernels for the BERT-large PairAveraging configuration (BASELINE.json #4):
//   add_layernorm : z = x + residual, y = LayerNorm(z) * gamma + beta      (one pass; backward from (dy, z, mean, rstd))
//   bias_gelu     : y = gelu(x + bias)                                      (bias of the FFN up-projection folded in)
// Both backward kernels produce their parameter gradients (dgamma/dbeta, dbias) with a deterministic two-stage column
// reduction (per-block partials in registers -> shared memory -> a finalize kernel), no atomics, like the BatchNorm family.
//
// The reference ships no model code at all; eager PyTorch spends 8 separate memory-bound kernels per layer
// on these two patterns (add, layer_norm fwd/bwd, bias add, gelu fwd/bwd, two column sums).
#include "common.cuh"
#include "kernels.h"

namespace kf {

namespace {

template <typename T> __device__ __forceinline__ void unpack8(const Vec16& v, float* f);
template <> __device__ __forceinline__ void unpack8<__nv_bfloat16>(const Vec16& v, float* f) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162*>(&v.w[i]);
        f[2 * i] = __low2float(h);
        f[2 * i + 1] = __high2float(h);
    }
}
template <> __device__ __forceinline__ void unpack8<__half>(const Vec16& v, float* f) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const __half2 h = *reinterpret_cast<const __half2*>(&v.w[i]);
        f[2 * i] = __low2float(h);
        f[2 * i + 1] = __high2float(h);
    }
}
template <typename T> __device__ __forceinline__ Vec16 pack8(const float* f);
template <> __device__ __forceinline__ Vec16 pack8<__nv_bfloat16>(const float* f) {
    Vec16 v;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
        v.w[i] = *reinterpret_cast<const uint32_t*>(&h);
    }
    return v;
}
template <> __device__ __forceinline__ Vec16 pack8<__half>(const float* f) {
    Vec16 v;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const __half2 h = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
        v.w[i] = *reinterpret_cast<const uint32_t*>(&h);
    }
    return v;
}

// ------------------------------------------------------------------------------------------------------------------
// LayerNorm: one warp per row, VPL 16-byte vectors (8 elements) per lane -> H <= 256 * VPL
// ------------------------------------------------------------------------------------------------------------------
constexpr int kLnWarps = 4;

template <typename T, int VPL>
__global__ void __launch_bounds__(kLnWarps * 32) add_ln_fwd_kernel(const T* __restrict__ x, const T* __restrict__ res, const T* __restrict__ gamma,
                                                                    const T* __restrict__ beta, T* __restrict__ z, T* __restrict__ y,
                                                                    float* __restrict__ mean, float* __restrict__ rstd, int rows, int H, float eps) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nvec = H / 8;
    float g[VPL][8], b[VPL][8];
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
        const int v = lane + 32 * i;
        if (v < nvec) {
            unpack8<T>(ld_plain(gamma + (size_t)v * 8), g[i]);
            unpack8<T>(ld_plain(beta + (size_t)v * 8), b[i]);
        }
    }
    for (int row = blockIdx.x * kLnWarps + warp; row < rows; row += gridDim.x * kLnWarps) {
        const size_t base = (size_t)row * H;
        float f[VPL][8];
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
            const int v = lane + 32 * i;
            if (v < nvec) {
                unpack8<T>(ld_ro_stream(x + base + (size_t)v * 8), f[i]);
                if (res) {
                    float r[8];
                    unpack8<T>(ld_ro_stream(res + base + (size_t)v * 8), r);
#pragma unroll
                    for (int j = 0; j < 8; ++j) f[i][j] += r[j];
                }
                // the saved sum is what backward normalises again: round it to storage precision FIRST so that forward
                // and backward see the same z
                const Vec16 zv = pack8<T>(f[i]);
                if (z) st_plain(z + base + (size_t)v * 8, zv);
                unpack8<T>(zv, f[i]);
#pragma unroll
                for (int j = 0; j < 8; ++j) s += f[i][j];
            }
        }
        const float mu = warp_sum(s) / (float)H;
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
            if (lane + 32 * i < nvec) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float d = f[i][j] - mu;
                    q = fmaf(d, d, q);
                }
            }
        }
        const float rs = rsqrtf(warp_sum(q) / (float)H + eps);
        if (lane == 0) {
            mean[row] = mu;
            rstd[row] = rs;
        }
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
            const int v = lane + 32 * i;
            if (v < nvec) {
                float o[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) o[j] = fmaf((f[i][j] - mu) * rs, g[i][j], b[i][j]);
                st_plain(y + base + (size_t)v * 8, pack8<T>(o));
            }
        }
    }
}

// backward: dz = rstd * (g - mean(g) - xhat * mean(g * xhat)), g = dy * gamma; partial[block][2][H] <- column sums of
// dy * xhat (dgamma) and dy (dbeta) over the rows this block handled
template <typename T, int VPL>
__global__ void __launch_bounds__(kLnWarps * 32) add_ln_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ z, const float* __restrict__ mean,
                                                                    const float* __restrict__ rstd, const T* __restrict__ gamma, T* __restrict__ dz,
                                                                    float* __restrict__ partial, int rows, int H) {
    extern __shared__ float sm[]; // [kLnWarps][2][H]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nvec = H / 8;
    float g[VPL][8], dg[VPL][8], db[VPL][8];
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
        const int v = lane + 32 * i;
        if (v < nvec) unpack8<T>(ld_plain(gamma + (size_t)v * 8), g[i]);
#pragma unroll
        for (int j = 0; j < 8; ++j) dg[i][j] = db[i][j] = 0.f;
    }
    const float invH = 1.f / (float)H;
    for (int row = blockIdx.x * kLnWarps + warp; row < rows; row += gridDim.x * kLnWarps) {
        const size_t base = (size_t)row * H;
        const float mu = mean[row], rs = rstd[row];
        float d[VPL][8], xh[VPL][8];
        float c1 = 0.f, c2 = 0.f;
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
            const int v = lane + 32 * i;
            if (v < nvec) {
                unpack8<T>(ld_ro_stream(dy + base + (size_t)v * 8), d[i]);
                unpack8<T>(ld_ro_stream(z + base + (size_t)v * 8), xh[i]);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    xh[i][j] = (xh[i][j] - mu) * rs;
                    dg[i][j] = fmaf(d[i][j], xh[i][j], dg[i][j]);
                    db[i][j] += d[i][j];
                    d[i][j] *= g[i][j];
                    c1 += d[i][j];
                    c2 = fmaf(d[i][j], xh[i][j], c2);
                }
            }
        }
        c1 = warp_sum(c1) * invH;
        c2 = warp_sum(c2) * invH;
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
            const int v = lane + 32 * i;
            if (v < nvec) {
                float o[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) o[j] = rs * (d[i][j] - c1 - xh[i][j] * c2);
                st_plain(dz + base + (size_t)v * 8, pack8<T>(o));
            }
        }
    }
    // fold the block's warps through shared memory, one partial row per block
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
        const int v = lane + 32 * i;
        if (v < nvec) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                sm[(warp * 2 + 0) * H + v * 8 + j] = dg[i][j];
                sm[(warp * 2 + 1) * H + v * 8 + j] = db[i][j];
            }
        }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < 2 * H; c += blockDim.x) {
        const int which = c / H, col = c - which * H;
        float a = 0.f;
#pragma unroll
        for (int w = 0; w < kLnWarps; ++w) a += sm[(w * 2 + which) * H + col];
        partial[((size_t)blockIdx.x * 2 + which) * H + col] = a;
    }
}

// out[c] = sum_b partial[b][c]  (c over `width` columns), written in the parameter dtype
template <typename T> __global__ void colsum_finalize_kernel(const float* __restrict__ partial, int nblocks, int width, T* __restrict__ out) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= width) return;
    float a = 0.f;
    for (int b = 0; b < nblocks; ++b) a += partial[(size_t)b * width + c];
    out[c] = Elem<T>::down(a);
}

// ------------------------------------------------------------------------------------------------------------------
// bias + GELU (exact, erf form — what nn.GELU() / F.gelu default to)
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float gelu_f(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float gelu_grad_f(float x) {
    const float cdf = 0.5f * (1.f + erff(x * 0.70710678118654752f));
    const float pdf = 0.3989422804014327f * __expf(-0.5f * x * x);
    return fmaf(x, pdf, cdf);
}

template <typename T>
__global__ void __launch_bounds__(256) bias_gelu_fwd_kernel(const T* __restrict__ x, const T* __restrict__ bias, T* __restrict__ y, size_t rows, int I) {
    const int nvec = I / 8;
    const size_t total = rows * (size_t)nvec;
    for (size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v < total; v += (size_t)gridDim.x * blockDim.x) {
        const int cv = (int)(v % (size_t)nvec);
        float f[8], b[8];
        unpack8<T>(ld_ro_stream(x + v * 8), f);
        unpack8<T>(ld_plain(bias + (size_t)cv * 8), b);
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = gelu_f(f[j] + b[j]);
        st_plain(y + v * 8, pack8<T>(f));
    }
}

// dx = dy * gelu'(x + bias); partial[block][I] <- column sums of dx over the rows this block handled.
// Thread t owns column vectors t, t + 256, ... of every row it visits, so its partial sums stay in registers.
template <typename T, int VPT>
__global__ void __launch_bounds__(256) bias_gelu_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ x, const T* __restrict__ bias,
                                                            T* __restrict__ dx, float* __restrict__ partial, int rows, int I) {
    const int nvec = I / 8;
    float b[VPT][8], acc[VPT][8];
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
        const int cv = threadIdx.x + 256 * i;
        if (cv < nvec) unpack8<T>(ld_plain(bias + (size_t)cv * 8), b[i]);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    }
    for (int row = blockIdx.x; row < rows; row += gridDim.x) {
        const size_t base = (size_t)row * I;
#pragma unroll
        for (int i = 0; i < VPT; ++i) {
            const int cv = threadIdx.x + 256 * i;
            if (cv < nvec) {
                float d[8], f[8];
                unpack8<T>(ld_ro_stream(dy + base + (size_t)cv * 8), d);
                unpack8<T>(ld_ro_stream(x + base + (size_t)cv * 8), f);
#pragma unroll
                for (int j = 0; j < 8; ++j) d[j] *= gelu_grad_f(f[j] + b[i][j]);
                const Vec16 out = pack8<T>(d);
                st_plain(dx + base + (size_t)cv * 8, out);
                unpack8<T>(out, d); // the bias gradient sums what the GEMMs downstream will see
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] += d[j];
            }
        }
    }
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
        const int cv = threadIdx.x + 256 * i;
        if (cv < nvec) {
#pragma unroll
            for (int j = 0; j < 8; ++j) partial[(size_t)blockIdx.x * I + (size_t)cv * 8 + j] = acc[i][j];
        }
    }
}

int ln_blocks(int rows) {
    const int need = (rows + kLnWarps - 1) / kLnWarps;
    return need < 148 * 8 ? (need < 1 ? 1 : need) : 148 * 8;
}

} // namespace

int add_ln_num_blocks(int rows) { return ln_blocks(rows) < 148 * 2 ? ln_blocks(rows) : 148 * 2; }
int bias_gelu_num_blocks(int rows) { return rows < 148 * 2 ? (rows < 1 ? 1 : rows) : 148 * 2; }

template <typename T>
static void add_ln_fwd_t(const void* x, const void* res, const void* gamma, const void* beta, void* z, void* y, float* mean, float* rstd, int rows,
                         int H, float eps, cudaStream_t st) {
    const int vpl = (H / 8 + 31) / 32;
    const int blocks = ln_blocks(rows);
#define KF_LN_FWD(V)                                                                                                                     \
    add_ln_fwd_kernel<T, V><<<blocks, kLnWarps * 32, 0, st>>>((const T*)x, (const T*)res, (const T*)gamma, (const T*)beta, (T*)z, (T*)y, mean, \
                                                              rstd, rows, H, eps)
    if (vpl <= 1) KF_LN_FWD(1);
    else if (vpl == 2) KF_LN_FWD(2);
    else if (vpl == 3) KF_LN_FWD(3);
    else if (vpl == 4) KF_LN_FWD(4);
    else if (vpl <= 8) KF_LN_FWD(8);
    else throw std::runtime_error("kungfu_b200: add_layernorm supports hidden sizes up to 2048");
#undef KF_LN_FWD
}

void launch_add_ln_fwd(const void* x, const void* res, const void* gamma, const void* beta, void* z, void* y, float* mean, float* rstd, int dtype,
                       int rows, int H, float eps, cudaStream_t st) {
    if (H % 8 || H < 8) throw std::runtime_error("kungfu_b200: add_layernorm needs a hidden size that is a multiple of 8");
    if (dtype == kBF16) add_ln_fwd_t<__nv_bfloat16>(x, res, gamma, beta, z, y, mean, rstd, rows, H, eps, st);
    else if (dtype == kF16) add_ln_fwd_t<__half>(x, res, gamma, beta, z, y, mean, rstd, rows, H, eps, st);
    else throw std::runtime_error("kungfu_b200: add_layernorm supports bf16 / f16 activations");
    count_launch();
    KF_CUDA_CHECK(cudaGetLastError());
}

// partial rows alternate dgamma / dbeta: sum the even rows into dgamma and the odd rows into dbeta
template <typename T>
__global__ void ln_param_finalize_kernel(const float* __restrict__ partial, int nblocks, int H, T* __restrict__ dgamma, T* __restrict__ dbeta) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= 2 * H) return;
    const int which = c / H, col = c - which * H;
    float a = 0.f;
    for (int b = 0; b < nblocks; ++b) a += partial[((size_t)b * 2 + which) * H + col];
    (which ? dbeta : dgamma)[col] = Elem<T>::down(a);
}

void launch_add_ln_bwd(const void* dy, const void* z, const float* mean, const float* rstd, const void* gamma, void* dz, void* dgamma, void* dbeta,
                       float* partial, int dtype, int rows, int H, cudaStream_t st) {
    if (H % 8 || H < 8) throw std::runtime_error("kungfu_b200: add_layernorm needs a hidden size that is a multiple of 8");
    const int vpl = (H / 8 + 31) / 32;
    if (vpl > 8) throw std::runtime_error("kungfu_b200: add_layernorm supports hidden sizes up to 2048");
    const int blocks = add_ln_num_blocks(rows);
    const size_t smem = (size_t)kLnWarps * 2 * H * sizeof(float);
#define KF_LN_BWD2(T, V)                                                                                                                  \
    do {                                                                                                                                  \
        static bool cfg = false;                                                                                                          \
        if (!cfg && smem > 48 * 1024) {                                                                                                   \
            KF_CUDA_CHECK(cudaFuncSetAttribute(add_ln_bwd_kernel<T, V>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));         \
            cfg = true;                                                                                                                   \
        }                                                                                                                                 \
        add_ln_bwd_kernel<T, V><<<blocks, kLnWarps * 32, smem, st>>>((const T*)dy, (const T*)z, mean, rstd, (const T*)gamma, (T*)dz, partial, \
                                                                     rows, H);                                                            \
        ln_param_finalize_kernel<T><<<(2 * H + 127) / 128, 128, 0, st>>>(partial, blocks, H, (T*)dgamma, (T*)dbeta);                      \
    } while (0)
#define KF_LN_BWD_V(T)                                                                                                                    \
    do {                                                                                                                                  \
        if (vpl <= 1) KF_LN_BWD2(T, 1);                                                                                                   \
        else if (vpl == 2) KF_LN_BWD2(T, 2);                                                                                              \
        else if (vpl == 3) KF_LN_BWD2(T, 3);                                                                                              \
        else if (vpl == 4) KF_LN_BWD2(T, 4);                                                                                              \
        else KF_LN_BWD2(T, 8);                                                                                                            \
    } while (0)
    if (dtype == kBF16) KF_LN_BWD_V(__nv_bfloat16);
    else if (dtype == kF16) KF_LN_BWD_V(__half);
    else throw std::runtime_error("kungfu_b200: add_layernorm supports bf16 / f16 activations");
#undef KF_LN_BWD_V
#undef KF_LN_BWD2
    count_launch(2);
    KF_CUDA_CHECK(cudaGetLastError());
}

void launch_bias_gelu_fwd(const void* x, const void* bias, void* y, int dtype, size_t rows, int I, cudaStream_t st) {
    if (I % 8) throw std::runtime_error("kungfu_b200: bias_gelu needs a width that is a multiple of 8");
    const size_t total = rows * (size_t)(I / 8);
    size_t blocks = (total + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    if (blocks < 1) blocks = 1;
    if (dtype == kBF16) bias_gelu_fwd_kernel<__nv_bfloat16><<<(int)blocks, 256, 0, st>>>((const __nv_bfloat16*)x, (const __nv_bfloat16*)bias, (__nv_bfloat16*)y, rows, I);
    else if (dtype == kF16) bias_gelu_fwd_kernel<__half><<<(int)blocks, 256, 0, st>>>((const __half*)x, (const __half*)bias, (__half*)y, rows, I);
    else throw std::runtime_error("kungfu_b200: bias_gelu supports bf16 / f16 activations");
    count_launch();
    KF_CUDA_CHECK(cudaGetLastError());
}

void launch_bias_gelu_bwd(const void* dy, const void* x, const void* bias, void* dx, void* dbias, float* partial, int dtype, int rows, int I,
                          cudaStream_t st) {
    if (I % 8) throw std::runtime_error("kungfu_b200: bias_gelu needs a width that is a multiple of 8");
    const int vpt = (I / 8 + 255) / 256;
    if (vpt > 8) throw std::runtime_error("kungfu_b200: bias_gelu supports widths up to 16384");
    const int blocks = bias_gelu_num_blocks(rows);
#define KF_BG(T, V) bias_gelu_bwd_kernel<T, V><<<blocks, 256, 0, st>>>((const T*)dy, (const T*)x, (const T*)bias, (T*)dx, partial, rows, I)
#define KF_BG_V(T)                                                                                                                        \
    do {                                                                                                                                  \
        if (vpt <= 1) KF_BG(T, 1);                                                                                                        \
        else if (vpt == 2) KF_BG(T, 2);                                                                                                   \
        else if (vpt <= 4) KF_BG(T, 4);                                                                                                   \
        else KF_BG(T, 8);                                                                                                                 \
        colsum_finalize_kernel<T><<<(I + 127) / 128, 128, 0, st>>>(partial, blocks, I, (T*)dbias);                                        \
    } while (0)
    if (dtype == kBF16) KF_BG_V(__nv_bfloat16);
    else if (dtype == kF16) KF_BG_V(__half);
    else throw std::runtime_error("kungfu_b200: bias_gelu supports bf16 / f16 activations");
#undef KF_BG_V
#undef KF_BG
    count_launch(2);
    KF_CUDA_CHECK(cudaGetLastError());
}

} // namespace kf
