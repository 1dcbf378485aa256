// kungfu_b200 — fused transformer element-wise kernels for the BERT-large PairAveraging configuration (BASELINE.json #4):
//   add_layernorm : z = x (+ lin_bias) + residual, y = LayerNorm(z) * gamma + beta   (one pass; backward from (dy, z, mean, rstd);
//                   lin_bias = bias of the GEMM that produced x, folded in: its gradient is one more column sum of dz)
//   bias_gelu     : y = gelu(x + bias)                                      (bias of the FFN up-projection folded in)
//   qkv_pack_bwd  : (dq, dk, dv) as attention backward leaves them (any strides) -> ONE contiguous [tokens, 3H] gradient of the fused QKV
//                   projection + its bias gradient (replaces autograd's stack/cat, a layout copy and a column-sum kernel)
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

// Hands back `v` through an empty asm statement: the compiler cannot prove the value loop-invariant any more, so the fp32 expansion of a
// packed parameter vector is recomputed where it is used (two ALU ops per element pair) instead of living in 8 registers for the whole
// kernel.
__device__ __forceinline__ Vec16 opaque(Vec16 v) {
    asm volatile("" : "+r"(v.w[0]), "+r"(v.w[1]), "+r"(v.w[2]), "+r"(v.w[3]));
    return v;
}

// ------------------------------------------------------------------------------------------------------------------
// LayerNorm: one warp per row, VPL 16-byte vectors (8 elements) per lane -> H <= 256 * VPL
// ------------------------------------------------------------------------------------------------------------------
constexpr int kLnWarps = 4;

template <typename T, int VPL>
__global__ void __launch_bounds__(kLnWarps * 32, (VPL <= 4 ? 4 : 1)) add_ln_fwd_kernel(const T* __restrict__ x, const T* __restrict__ res, const T* __restrict__ lin_bias,
                                                                    const T* __restrict__ gamma, const T* __restrict__ beta, T* __restrict__ z,
                                                                    T* __restrict__ y, float* __restrict__ mean, float* __restrict__ rstd, int rows,
                                                                    int H, float eps) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nvec = H / 8;
    // parameters stay PACKED in registers (4 per 8 elements, unpacked where they are used): with fp32 copies of gamma / beta / bias the
    // kernel needed 157 registers -> 3 CTAs per SM -> 15 % occupancy, and it is a latency-bound streaming kernel (ncu: long-scoreboard)
    Vec16 gp[VPL], bp[VPL], lp[VPL];
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
        const int v = lane + 32 * i;
        gp[i] = bp[i] = lp[i] = Vec16{{0u, 0u, 0u, 0u}};
        if (v < nvec) {
            gp[i] = ld_plain(gamma + (size_t)v * 8);
            bp[i] = ld_plain(beta + (size_t)v * 8);
            if (lin_bias) lp[i] = ld_plain(lin_bias + (size_t)v * 8);
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
                if (lin_bias) {
                    float lb[8];
                    unpack8<T>(opaque(lp[i]), lb);
#pragma unroll
                    for (int j = 0; j < 8; ++j) f[i][j] += lb[j];
                }
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
                float o[8], g[8], b[8];
                unpack8<T>(opaque(gp[i]), g);
                unpack8<T>(opaque(bp[i]), b);
#pragma unroll
                for (int j = 0; j < 8; ++j) o[j] = fmaf((f[i][j] - mu) * rs, g[j], b[j]);
                st_plain(y + base + (size_t)v * 8, pack8<T>(o));
            }
        }
    }
}

// backward: dz = rstd * (g - mean(g) - xhat * mean(g * xhat)), g = dy * gamma; partial[block][NP][H] <- column sums of
// dy * xhat (dgamma), dy (dbeta) and, with LB (a linear bias was folded into the forward), dz as stored (dlin_bias)
template <typename T, int VPL, bool LB>
__global__ void __launch_bounds__(kLnWarps * 32) add_ln_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ z, const float* __restrict__ mean,
                                                                    const float* __restrict__ rstd, const T* __restrict__ gamma, T* __restrict__ dz,
                                                                    float* __restrict__ partial, int rows, int H) {
    extern __shared__ float sm[]; // [kLnWarps][NP][H]
    constexpr int NP = LB ? 3 : 2;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nvec = H / 8;
    // Unlike the forward kernel this one keeps gamma expanded: three fp32 accumulator sets (96 registers) cannot be avoided, and the
    // variant with packed gamma + __launch_bounds__(128, 3) (168 registers, 128 bytes of spills, 3 CTAs per SM) measured SLOWER:
    // 23.5 us vs 18.2 us per call at 4096 x 1024.
    float g[VPL][8], dg[VPL][8], db[VPL][8], dl[LB ? VPL : 1][8];
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
        const int v = lane + 32 * i;
        if (v < nvec) unpack8<T>(ld_plain(gamma + (size_t)v * 8), g[i]);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            dg[i][j] = db[i][j] = 0.f;
            if (LB) dl[i][j] = 0.f;
        }
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
                const Vec16 ov = pack8<T>(o);
                st_plain(dz + base + (size_t)v * 8, ov);
                if (LB) { // the bias gradient sums what the GEMMs downstream will see
                    unpack8<T>(ov, o);
#pragma unroll
                    for (int j = 0; j < 8; ++j) dl[i][j] += o[j];
                }
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
                sm[(warp * NP + 0) * H + v * 8 + j] = dg[i][j];
                sm[(warp * NP + 1) * H + v * 8 + j] = db[i][j];
                if (LB) sm[(warp * NP + 2) * H + v * 8 + j] = dl[i][j];
            }
        }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < NP * H; c += blockDim.x) {
        const int which = c / H, col = c - which * H;
        float a = 0.f;
#pragma unroll
        for (int w = 0; w < kLnWarps; ++w) a += sm[(w * NP + which) * H + col];
        partial[((size_t)blockIdx.x * NP + which) * H + col] = a;
    }
}

// out_k[c] = sum_b partial[b][k * H + c] for the W = NP * H columns of `partial` ([nblocks][W]), written in the parameter dtype.
// 32 columns per block, 32 warps stride over the partial rows (the first version walked the rows serially per column: 11-13 us for
// 296 x 4096 floats, pure latency; 8 warps: 5.6 us), folded through shared memory in a fixed order -> deterministic.
template <typename T>
__global__ void __launch_bounds__(1024) colsum_finalize_kernel(const float* __restrict__ partial, int nblocks, int W, int H, T* __restrict__ o0,
                                                               T* __restrict__ o1, T* __restrict__ o2) {
    __shared__ float sm[32][33];
    const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + cx;
    float a = 0.f;
    if (c < W) {
#pragma unroll 4
        for (int b = ry; b < nblocks; b += 32) a += partial[(size_t)b * W + c];
    }
    sm[ry][cx] = a;
    __syncthreads();
    if (ry == 0 && c < W) {
        float t = 0.f;
#pragma unroll
        for (int r = 0; r < 32; ++r) t += sm[r][cx];
        const int which = c / H, col = c - which * H;
        T* out = which == 0 ? o0 : which == 1 ? o1 : o2;
        out[col] = Elem<T>::down(t);
    }
}

// ------------------------------------------------------------------------------------------------------------------
// bias + GELU (exact, erf form — what nn.GELU() / F.gelu default to)
// ------------------------------------------------------------------------------------------------------------------
// Phi(x) = 0.5 erfc(-x / sqrt 2) through the Abramowitz-Stegun 7.1.26 rational form (|error| < 1.5e-7, far below the bf16 / fp16 output
// step): ONE exp and ONE reciprocal on the MUFU plus five FMAs, and the same exp(-x^2/2) is the density the gradient needs. erff() +
// expf() cost ~60 instructions per element and made both kernels compute-bound (bwd 47 us for 3 x 32 MiB = 2.0 TB/s).
__device__ __forceinline__ void gelu_parts(float x, float& cdf, float& pdf_over_c) {
    const float u = fabsf(x) * 0.70710678118654752f;
    const float e = __expf(-u * u);
    const float t = __fdividef(1.f, fmaf(0.3275911f, u, 1.f));
    float p = fmaf(1.061405429f, t, -1.453152027f);
    p = fmaf(p, t, 1.421413741f);
    p = fmaf(p, t, -0.284496736f);
    p = fmaf(p, t, 0.254829592f);
    const float q = 0.5f * p * t * e; // 0.5 erfc(|x| / sqrt 2)
    cdf = x >= 0.f ? 1.f - q : q;
    pdf_over_c = e;
}
__device__ __forceinline__ float gelu_f(float x) {
    float cdf, e;
    gelu_parts(x, cdf, e);
    return x * cdf;
}
__device__ __forceinline__ float gelu_grad_f(float x) {
    float cdf, e;
    gelu_parts(x, cdf, e);
    return fmaf(x * 0.3989422804014327f, e, cdf);
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

// ------------------------------------------------------------------------------------------------------------------
// gradient of the fused QKV projection: dq / dk / dv are [B, heads, S, d] tensors with whatever strides the attention backward chose
// (innermost stride 1); out is the contiguous [B * S, 3 * heads * d] matrix the weight / input GEMMs consume, laid out like the forward
// output (token-major, then q|k|v, then head, then d). Thread t owns column vectors t, t + 256, ... of every row it visits (bias_gelu_bwd's
// scheme): per-column partial sums stay in registers -> partial[block][3H] -> colsum_finalize = the bias gradient.
// ------------------------------------------------------------------------------------------------------------------
struct QkvSrc {
    const void* p[3];
    long long sb[3], sh[3], ss[3]; // element strides of the batch, head and sequence dimensions
};

template <typename T, int VPT>
__global__ void __launch_bounds__(256) qkv_pack_bwd_kernel(QkvSrc src, T* __restrict__ out, float* __restrict__ partial, int B, int S, int heads, int d) {
    const int H = heads * d;
    const int W = 3 * H, nvec = W / 8;
    const T* base[VPT];
    long long sb[VPT], ss[VPT];
    float acc[VPT][8];
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
        const int cv = threadIdx.x + 256 * i;
        base[i] = nullptr;
        sb[i] = ss[i] = 0;
        if (cv < nvec) {
            const int col = cv * 8;
            const int which = col / H, within = col - which * H;
            const int head = within / d, e = within - head * d;
            // selects, not indexing: a dynamically indexed kernel parameter is copied to local memory first
            const void* p = which == 0 ? src.p[0] : which == 1 ? src.p[1] : src.p[2];
            const long long sh = which == 0 ? src.sh[0] : which == 1 ? src.sh[1] : src.sh[2];
            sb[i] = which == 0 ? src.sb[0] : which == 1 ? src.sb[1] : src.sb[2];
            ss[i] = which == 0 ? src.ss[0] : which == 1 ? src.ss[1] : src.ss[2];
            base[i] = (const T*)p + (long long)head * sh + e;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    }
    const int rows = B * S;
    for (int row = blockIdx.x; row < rows; row += gridDim.x) {
        const int b = row / S, t = row - b * S;
#pragma unroll
        for (int i = 0; i < VPT; ++i) {
            const int cv = threadIdx.x + 256 * i;
            if (cv < nvec) {
                const Vec16 v = ld_ro_stream(base[i] + (long long)b * sb[i] + (long long)t * ss[i]);
                st_plain(out + (size_t)row * W + (size_t)cv * 8, v);
                float f[8];
                unpack8<T>(v, f);
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] += f[j];
            }
        }
    }
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
        const int cv = threadIdx.x + 256 * i;
        if (cv < nvec) {
#pragma unroll
            for (int j = 0; j < 8; ++j) partial[(size_t)blockIdx.x * W + (size_t)cv * 8 + j] = acc[i][j];
        }
    }
}

} // namespace

// backward: 2 CTAs of 4 warps are resident per SM (225 registers)
int add_ln_num_blocks(int rows) { return ln_blocks(rows) < 148 * 2 ? ln_blocks(rows) : 148 * 2; }
// 4 CTAs of 256 threads per SM: the kernel is latency-bound (ncu: 2 CTAs/SM = 24 % occupancy, long-scoreboard stalls, 2.2 TB/s)
int bias_gelu_num_blocks(int rows) { return rows < 148 * 4 ? (rows < 1 ? 1 : rows) : 148 * 4; }
int qkv_pack_num_blocks(int rows) { return rows < 148 * 4 ? (rows < 1 ? 1 : rows) : 148 * 4; }

template <typename T>
static void add_ln_fwd_t(const void* x, const void* res, const void* lin_bias, const void* gamma, const void* beta, void* z, void* y, float* mean,
                         float* rstd, int rows, int H, float eps, cudaStream_t st) {
    const int vpl = (H / 8 + 31) / 32;
    const int blocks = ln_blocks(rows);
#define KF_LN_FWD(V)                                                                                                                       \
    add_ln_fwd_kernel<T, V><<<blocks, kLnWarps * 32, 0, st>>>((const T*)x, (const T*)res, (const T*)lin_bias, (const T*)gamma, (const T*)beta, \
                                                              (T*)z, (T*)y, mean, rstd, rows, H, eps)
    if (vpl <= 1) KF_LN_FWD(1);
    else if (vpl == 2) KF_LN_FWD(2);
    else if (vpl == 3) KF_LN_FWD(3);
    else if (vpl == 4) KF_LN_FWD(4);
    else if (vpl <= 8) KF_LN_FWD(8);
    else throw std::runtime_error("kungfu_b200: add_layernorm supports hidden sizes up to 2048");
#undef KF_LN_FWD
}

void launch_add_ln_fwd(const void* x, const void* res, const void* lin_bias, const void* gamma, const void* beta, void* z, void* y, float* mean,
                       float* rstd, int dtype, int rows, int H, float eps, cudaStream_t st) {
    if (H % 8 || H < 8) throw std::runtime_error("kungfu_b200: add_layernorm needs a hidden size that is a multiple of 8");
    if (dtype == kBF16) add_ln_fwd_t<__nv_bfloat16>(x, res, lin_bias, gamma, beta, z, y, mean, rstd, rows, H, eps, st);
    else if (dtype == kF16) add_ln_fwd_t<__half>(x, res, lin_bias, gamma, beta, z, y, mean, rstd, rows, H, eps, st);
    else throw std::runtime_error("kungfu_b200: add_layernorm supports bf16 / f16 activations");
    count_launch();
    KF_CUDA_CHECK(cudaGetLastError());
}

template <typename T, int V, bool LB>
static void add_ln_bwd_launch(const void* dy, const void* z, const float* mean, const float* rstd, const void* gamma, void* dz, float* partial, int rows,
                              int H, int blocks, cudaStream_t st) {
    const size_t smem = (size_t)kLnWarps * (LB ? 3 : 2) * H * sizeof(float);
    static bool cfg = false;
    if (!cfg && smem > 48 * 1024) {
        KF_CUDA_CHECK(cudaFuncSetAttribute(add_ln_bwd_kernel<T, V, LB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        cfg = true;
    }
    add_ln_bwd_kernel<T, V, LB><<<blocks, kLnWarps * 32, smem, st>>>((const T*)dy, (const T*)z, mean, rstd, (const T*)gamma, (T*)dz, partial, rows, H);
}

template <typename T>
static void add_ln_bwd_t(const void* dy, const void* z, const float* mean, const float* rstd, const void* gamma, void* dz, void* dgamma, void* dbeta,
                         void* dlin_bias, float* partial, int rows, int H, cudaStream_t st) {
    const int vpl = (H / 8 + 31) / 32;
    const int blocks = add_ln_num_blocks(rows);
    const bool lb = dlin_bias != nullptr;
#define KF_LN_BWD(V)                                                                                                                       \
    do {                                                                                                                                   \
        if (lb) add_ln_bwd_launch<T, V, true>(dy, z, mean, rstd, gamma, dz, partial, rows, H, blocks, st);                                 \
        else add_ln_bwd_launch<T, V, false>(dy, z, mean, rstd, gamma, dz, partial, rows, H, blocks, st);                                   \
    } while (0)
    if (vpl <= 1) KF_LN_BWD(1);
    else if (vpl == 2) KF_LN_BWD(2);
    else if (vpl == 3) KF_LN_BWD(3);
    else if (vpl == 4) KF_LN_BWD(4);
    else KF_LN_BWD(8);
#undef KF_LN_BWD
    const int W = (lb ? 3 : 2) * H;
    colsum_finalize_kernel<T><<<(W + 31) / 32, 1024, 0, st>>>(partial, blocks, W, H, (T*)dgamma, (T*)dbeta, (T*)dlin_bias);
}

void launch_add_ln_bwd(const void* dy, const void* z, const float* mean, const float* rstd, const void* gamma, void* dz, void* dgamma, void* dbeta,
                       void* dlin_bias, float* partial, int dtype, int rows, int H, cudaStream_t st) {
    if (H % 8 || H < 8) throw std::runtime_error("kungfu_b200: add_layernorm needs a hidden size that is a multiple of 8");
    if ((H / 8 + 31) / 32 > 8) throw std::runtime_error("kungfu_b200: add_layernorm supports hidden sizes up to 2048");
    if (dtype == kBF16) add_ln_bwd_t<__nv_bfloat16>(dy, z, mean, rstd, gamma, dz, dgamma, dbeta, dlin_bias, partial, rows, H, st);
    else if (dtype == kF16) add_ln_bwd_t<__half>(dy, z, mean, rstd, gamma, dz, dgamma, dbeta, dlin_bias, partial, rows, H, st);
    else throw std::runtime_error("kungfu_b200: add_layernorm supports bf16 / f16 activations");
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
        colsum_finalize_kernel<T><<<(I + 31) / 32, 1024, 0, st>>>(partial, blocks, I, I, (T*)dbias, (T*)nullptr, (T*)nullptr);             \
    } while (0)
    if (dtype == kBF16) KF_BG_V(__nv_bfloat16);
    else if (dtype == kF16) KF_BG_V(__half);
    else throw std::runtime_error("kungfu_b200: bias_gelu supports bf16 / f16 activations");
#undef KF_BG_V
#undef KF_BG
    count_launch(2);
    KF_CUDA_CHECK(cudaGetLastError());
}

void launch_qkv_pack_bwd(const void* dq, const void* dk, const void* dv, const long long* strides, void* out, void* dbias, float* partial, int dtype,
                         int B, int S, int heads, int d, cudaStream_t st) {
    const int W = 3 * heads * d;
    if (d % 8) throw std::runtime_error("kungfu_b200: qkv_pack_bwd needs a head size that is a multiple of 8");
    const int vpt = (W / 8 + 255) / 256;
    if (vpt > 4) throw std::runtime_error("kungfu_b200: qkv_pack_bwd supports hidden sizes up to 2730");
    QkvSrc src;
    src.p[0] = dq;
    src.p[1] = dk;
    src.p[2] = dv;
    for (int i = 0; i < 3; ++i) {
        src.sb[i] = strides[3 * i + 0];
        src.sh[i] = strides[3 * i + 1];
        src.ss[i] = strides[3 * i + 2];
        if ((src.sb[i] | src.sh[i] | src.ss[i]) % 8 || ((uintptr_t)src.p[i] & 15))
            throw std::runtime_error("kungfu_b200: qkv_pack_bwd needs 16-byte aligned rows (strides that are multiples of 8 elements)");
    }
    const int blocks = qkv_pack_num_blocks(B * S);
#define KF_QP(T, V) qkv_pack_bwd_kernel<T, V><<<blocks, 256, 0, st>>>(src, (T*)out, partial, B, S, heads, d)
#define KF_QP_V(T)                                                                                                                        \
    do {                                                                                                                                  \
        if (vpt <= 1) KF_QP(T, 1);                                                                                                        \
        else if (vpt == 2) KF_QP(T, 2);                                                                                                   \
        else if (vpt == 3) KF_QP(T, 3);                                                                                                   \
        else KF_QP(T, 4);                                                                                                                 \
        colsum_finalize_kernel<T><<<(W + 31) / 32, 1024, 0, st>>>(partial, blocks, W, W, (T*)dbias, (T*)nullptr, (T*)nullptr);             \
    } while (0)
    if (dtype == kBF16) KF_QP_V(__nv_bfloat16);
    else if (dtype == kF16) KF_QP_V(__half);
    else throw std::runtime_error("kungfu_b200: qkv_pack_bwd supports bf16 / f16 activations");
#undef KF_QP_V
#undef KF_QP
    count_launch(2);
    KF_CUDA_CHECK(cudaGetLastError());
}

} // namespace kf
