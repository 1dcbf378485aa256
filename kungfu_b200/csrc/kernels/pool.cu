// kungfu_b200 — 3x3 / stride-2 / pad-1 max pooling for channels_last bf16/fp16 activations (the ResNet stem), forward
// and backward. PyTorch's `max_pool_{forward,backward}_nhwc` take 0.75 ms + 1.41 ms per step at batch 256 on a B200
// (profiles/resnet50_step_kernels.md) for an op whose HBM roofline is ~0.1 ms each: they keep int64 indices and run a
// scalar window loop. Here one thread owns an 8-channel vector of one output (forward) or input (backward) pixel; the
// arg-max position inside the 3x3 window is stored as one byte per element and backward is a GATHER over the <= 4
// windows that cover an input pixel — no atomics, every byte of dx written exactly once.
#include "common.cuh"
#include "kernels.h"

namespace kf {

namespace {

// Two channels at a time on the packed half-precision ALU: the window maximum is 8 HMNMX2 per channel pair instead of
// 9 x 2 x (convert, compare, NaN test, two selects), which had made these kernels issue-bound (ncu: sm throughput 83 %, DRAM
// 30 %) rather than memory-bound.
template <typename T> struct Pair2;
template <> struct Pair2<__nv_bfloat16> {
    using P = __nv_bfloat162;
    static constexpr uint32_t kNegInf = 0xFF80FF80u;
};
template <> struct Pair2<__half> {
    using P = __half2;
    static constexpr uint32_t kNegInf = 0xFC00FC00u;
};
template <typename T> __device__ __forceinline__ typename Pair2<T>::P as_pair(uint32_t w) {
    typename Pair2<T>::P p;
    *reinterpret_cast<uint32_t*>(&p) = w;
    return p;
}
template <typename T> __device__ __forceinline__ uint32_t pair_max_nan(uint32_t a, uint32_t b) {
    const auto m = __hmax2_nan(as_pair<T>(a), as_pair<T>(b));
    return *reinterpret_cast<const uint32_t*>(&m);
}
template <typename T> __device__ __forceinline__ uint32_t pair_eq_mask(uint32_t a, uint32_t b) { return __heq2_mask(as_pair<T>(a), as_pair<T>(b)); }
template <typename T> __device__ __forceinline__ uint32_t pair_nan_mask(uint32_t a) { return __hneu2_mask(as_pair<T>(a), as_pair<T>(a)); }

template <typename T>
__global__ void __launch_bounds__(256) maxpool3x3s2_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, uint8_t* __restrict__ idx, int N, int H,
                                                               int W, int C, int OH, int OW) {
    // grid = (ceil(OH*OW*C/8 / 256), N): one thread per 8-channel vector of one output pixel, 32-bit index math only
    // (the flat 64-bit version spent ~300 instructions per thread in three 64-bit div/mod pairs)
    const uint32_t cvn = (uint32_t)C / 8;
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (uint32_t)OH * OW * cvn) return;
    const uint32_t zero = (uint32_t)N >> 31; // always 0, opaque to the compiler
    {
        const int n = blockIdx.y;
        const int cv = (int)(t % cvn);
        const uint32_t p = t / cvn;
        const int ow = (int)(p % (uint32_t)OW);
        const int oh = (int)(p / (uint32_t)OW);
        // all nine window loads are issued before the first compare (out-of-range taps read a clamped, valid address and are
        // ignored), see common.cuh::hold4 for why the dependency below is needed
        Vec16 tap[9];
        bool ok[9];
#pragma unroll
        for (int kh = 0; kh < 3; ++kh) {
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) {
                const int ih = oh * 2 - 1 + kh, iw = ow * 2 - 1 + kw;
                ok[kh * 3 + kw] = ih >= 0 && ih < H && iw >= 0 && iw < W;
                const int ch = min(max(ih, 0), H - 1), cw = min(max(iw, 0), W - 1);
                tap[kh * 3 + kw] = ld_ro_stream(x + (((size_t)n * H + ch) * W + cw) * C + (size_t)cv * 8);
            }
        }
        {
            uint32_t m = 0;
#pragma unroll
            for (int k = 1; k < 9; ++k) m ^= tap[k].w[0];
            m &= zero;
            tap[0].w[0] ^= m, tap[0].w[1] ^= m, tap[0].w[2] ^= m, tap[0].w[3] ^= m;
        }
        uint32_t best[4], arg[4], found[4], nan = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) best[j] = Pair2<T>::kNegInf, arg[j] = 0, found[j] = 0;
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            if (!ok[k]) continue;
#pragma unroll
            for (int j = 0; j < 4; ++j) best[j] = pair_max_nan<T>(best[j], tap[k].w[j]);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) nan |= pair_nan_mask<T>(best[j]);
        // arg-max = first tap equal to the maximum (PyTorch: first maximum wins); 16-bit lanes hold the tap number
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            if (!ok[k]) continue;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint32_t eq = pair_eq_mask<T>(tap[k].w[j], best[j]);
                arg[j] |= eq & ~found[j] & ((uint32_t)k * 0x00010001u);
                found[j] |= eq;
            }
        }
        if (nan) { // rare: NaN compares unequal to itself; PyTorch moves the arg-max to every NaN it meets, i.e. the last one
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint32_t want = pair_nan_mask<T>(best[j]);
                for (int k = 0; k < 9; ++k) {
                    if (!ok[k]) continue;
                    const uint32_t isn = pair_nan_mask<T>(tap[k].w[j]) & want;
                    arg[j] = (arg[j] & ~isn) | (isn & ((uint32_t)k * 0x00010001u));
                }
            }
        }
        Vec16 o;
#pragma unroll
        for (int j = 0; j < 4; ++j) o.w[j] = best[j];
        // lanes -> bytes: channel i of the vector owns byte i of the packed index word
        const uint64_t packed = (uint64_t)__byte_perm(arg[0], arg[1], 0x6420) | ((uint64_t)__byte_perm(arg[2], arg[3], 0x6420) << 32);
        const size_t ov = (((size_t)n * OH + oh) * OW + ow) * cvn + cv;
        st_stream(y + ov * 8, o);
        *reinterpret_cast<uint64_t*>(idx + ov * 8) = packed;
    }
}

template <typename T>
__global__ void __launch_bounds__(256) maxpool3x3s2_bwd_kernel(const T* __restrict__ dy, const uint8_t* __restrict__ idx, T* __restrict__ dx, int N,
                                                               int H, int W, int C, int OH, int OW) {
    const uint32_t cvn = (uint32_t)C / 8;
    const uint32_t tt = blockIdx.x * blockDim.x + threadIdx.x; // grid = (ceil(H*W*C/8 / 256), N)
    if (tt >= (uint32_t)H * W * cvn) return;
    const uint32_t zero = (uint32_t)N >> 31; // always 0, opaque to the compiler
    {
        const int n = blockIdx.y;
        const int cv = (int)(tt % cvn);
        const uint32_t p = tt / cvn;
        const int iw = (int)(p % (uint32_t)W);
        const int ih = (int)(p / (uint32_t)W);
        const size_t t = (size_t)n * H * W * cvn + tt;
        // output windows covering (ih, iw): oh in {floor((ih+1)/2) - 1 .. floor((ih+1)/2)} with 0 <= ih - (2 oh - 1) <= 2.
        // The four candidate (dy, arg-max) pairs are loaded together (clamped addresses for the invalid ones), then gathered.
        const int oh1 = (ih + 1) >> 1, ow1 = (iw + 1) >> 1;
        Vec16 g4[4];
        uint64_t a4[4];
        uint32_t me[4];
#pragma unroll
        for (int a = 0; a < 2; ++a) {
#pragma unroll
            for (int b = 0; b < 2; ++b) {
                const int oh = oh1 - a, ow = ow1 - b;
                const int kh = ih - (oh * 2 - 1), kw = iw - (ow * 2 - 1);
                const int k = a * 2 + b;
                const bool ok = oh >= 0 && oh < OH && kh >= 0 && kh <= 2 && ow >= 0 && ow < OW && kw >= 0 && kw <= 2;
                me[k] = ok ? (uint32_t)(kh * 3 + kw) * 0x01010101u : 0xffffffffu; // 0xff never matches a stored tap number
                const int ch = min(max(oh, 0), OH - 1), cw = min(max(ow, 0), OW - 1);
                const size_t ov = (((size_t)n * OH + ch) * OW + cw) * cvn + cv;
                a4[k] = __ldg(reinterpret_cast<const unsigned long long*>(idx + ov * 8));
                g4[k] = ld_ro_stream(dy + ov * 8);
            }
        }
        hold4(g4[0], g4[1], g4[2], g4[3], zero);
        float acc[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = 0.f;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            // byte-wise compare of the eight stored tap numbers, widened to 16-bit lane masks, selects this pixel's share of dy
            const uint32_t mlo = __vcmpeq4((uint32_t)a4[k], me[k]), mhi = __vcmpeq4((uint32_t)(a4[k] >> 32), me[k]);
            Vec16 sel;
            sel.w[0] = g4[k].w[0] & __byte_perm(mlo, 0, 0x1100);
            sel.w[1] = g4[k].w[1] & __byte_perm(mlo, 0, 0x3322);
            sel.w[2] = g4[k].w[2] & __byte_perm(mhi, 0, 0x1100);
            sel.w[3] = g4[k].w[3] & __byte_perm(mhi, 0, 0x3322);
            const AccVec<T> g = unpack<T>(sel);
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] += (float)g.a[i];
        }
        AccVec<T> o;
#pragma unroll
        for (int i = 0; i < 8; ++i) o.a[i] = acc[i];
        st_stream(dx + t * 8, pack<T>(o));
    }
}

dim3 pool_grid(size_t per_image, int N) {
    if (N > 65535 || per_image > 0x7fffffffull) throw std::runtime_error("kungfu_b200 maxpool: tensor too large for the 2-D grid");
    return dim3((unsigned)((per_image + 255) / 256), (unsigned)N, 1);
}

} // namespace

void launch_maxpool3x3s2_fwd(const void* x, void* y, uint8_t* idx, int dtype, int N, int H, int W, int C, cudaStream_t st) {
    if (C % 8) throw std::runtime_error("kungfu_b200 maxpool: C must be a multiple of 8");
    const int OH = (H + 2 - 3) / 2 + 1, OW = (W + 2 - 3) / 2 + 1;
    const dim3 grid = pool_grid((size_t)OH * OW * (C / 8), N);
    if (dtype == kBF16) maxpool3x3s2_fwd_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)y, idx, N, H, W, C, OH, OW);
    else if (dtype == kF16) maxpool3x3s2_fwd_kernel<__half><<<grid, 256, 0, st>>>((const __half*)x, (__half*)y, idx, N, H, W, C, OH, OW);
    else throw std::runtime_error("kungfu_b200 maxpool: bf16/fp16 only");
    count_launch();
    KF_CUDA_CHECK(cudaGetLastError());
}

void launch_maxpool3x3s2_bwd(const void* dy, const uint8_t* idx, void* dx, int dtype, int N, int H, int W, int C, cudaStream_t st) {
    if (C % 8) throw std::runtime_error("kungfu_b200 maxpool: C must be a multiple of 8");
    const int OH = (H + 2 - 3) / 2 + 1, OW = (W + 2 - 3) / 2 + 1;
    const dim3 grid = pool_grid((size_t)H * W * (C / 8), N);
    if (dtype == kBF16) maxpool3x3s2_bwd_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)dy, idx, (__nv_bfloat16*)dx, N, H, W, C, OH, OW);
    else if (dtype == kF16) maxpool3x3s2_bwd_kernel<__half><<<grid, 256, 0, st>>>((const __half*)dy, idx, (__half*)dx, N, H, W, C, OH, OW);
    else throw std::runtime_error("kungfu_b200 maxpool: bf16/fp16 only");
    count_launch();
    KF_CUDA_CHECK(cudaGetLastError());
}

} // namespace kf
