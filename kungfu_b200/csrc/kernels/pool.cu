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

template <typename T>
__global__ void __launch_bounds__(256) maxpool3x3s2_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, uint8_t* __restrict__ idx, int N, int H,
                                                               int W, int C, int OH, int OW) {
    const int cvn = C / 8;
    const size_t total = (size_t)N * OH * OW * cvn;
    const uint32_t zero = (uint32_t)(total >> 62); // always 0, opaque to the compiler
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
        const int cv = (int)(t % cvn);
        size_t p = t / cvn;
        const int ow = (int)(p % OW);
        p /= OW;
        const int oh = (int)(p % OH);
        const int n = (int)(p / OH);
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
        float best[8];
        uint32_t arg[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            best[i] = -INFINITY;
            arg[i] = 0;
        }
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            if (!ok[k]) continue;
            const AccVec<T> v = unpack<T>(tap[k]);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float f = v.a[i];
                if (f > best[i] || f != f) { // NaN propagates, first maximum wins (PyTorch semantics)
                    best[i] = f;
                    arg[i] = (uint32_t)k;
                }
            }
        }
        AccVec<T> o;
        uint64_t packed = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            o.a[i] = best[i];
            packed |= (uint64_t)arg[i] << (8 * i);
        }
        const size_t ov = (((size_t)n * OH + oh) * OW + ow) * cvn + cv;
        st_stream(y + ov * 8, pack<T>(o));
        *reinterpret_cast<uint64_t*>(idx + ov * 8) = packed;
    }
}

template <typename T>
__global__ void __launch_bounds__(256) maxpool3x3s2_bwd_kernel(const T* __restrict__ dy, const uint8_t* __restrict__ idx, T* __restrict__ dx, int N,
                                                               int H, int W, int C, int OH, int OW) {
    const int cvn = C / 8;
    const size_t total = (size_t)N * H * W * cvn;
    const uint32_t zero = (uint32_t)(total >> 62); // always 0, opaque to the compiler
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
        const int cv = (int)(t % cvn);
        size_t p = t / cvn;
        const int iw = (int)(p % W);
        p /= W;
        const int ih = (int)(p % H);
        const int n = (int)(p / H);
        // output windows covering (ih, iw): oh in {floor((ih+1)/2) - 1 .. floor((ih+1)/2)} with 0 <= ih - (2 oh - 1) <= 2.
        // The four candidate (dy, arg-max) pairs are loaded together (clamped addresses for the invalid ones), then gathered.
        const int oh1 = (ih + 1) >> 1, ow1 = (iw + 1) >> 1;
        Vec16 g4[4];
        uint64_t a4[4];
        uint32_t me[4];
        bool ok[4];
#pragma unroll
        for (int a = 0; a < 2; ++a) {
#pragma unroll
            for (int b = 0; b < 2; ++b) {
                const int oh = oh1 - a, ow = ow1 - b;
                const int kh = ih - (oh * 2 - 1), kw = iw - (ow * 2 - 1);
                const int k = a * 2 + b;
                ok[k] = oh >= 0 && oh < OH && kh >= 0 && kh <= 2 && ow >= 0 && ow < OW && kw >= 0 && kw <= 2;
                me[k] = (uint32_t)(kh * 3 + kw);
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
            if (!ok[k]) continue;
            const AccVec<T> g = unpack<T>(g4[k]);
#pragma unroll
            for (int i = 0; i < 8; ++i)
                if (((a4[k] >> (8 * i)) & 0xff) == me[k]) acc[i] += (float)g.a[i];
        }
        AccVec<T> o;
#pragma unroll
        for (int i = 0; i < 8; ++i) o.a[i] = acc[i];
        st_stream(dx + t * 8, pack<T>(o));
    }
}

int pool_grid(size_t work) {
    size_t b = (work + 255) / 256;
    if (b > 148 * 16) b = 148 * 16;
    return (int)(b < 1 ? 1 : b);
}

} // namespace

void launch_maxpool3x3s2_fwd(const void* x, void* y, uint8_t* idx, int dtype, int N, int H, int W, int C, cudaStream_t st) {
    if (C % 8) throw std::runtime_error("kungfu_b200 maxpool: C must be a multiple of 8");
    const int OH = (H + 2 - 3) / 2 + 1, OW = (W + 2 - 3) / 2 + 1;
    const size_t work = (size_t)N * OH * OW * (C / 8);
    if (dtype == kBF16) maxpool3x3s2_fwd_kernel<__nv_bfloat16><<<pool_grid(work), 256, 0, st>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)y, idx, N, H, W, C, OH, OW);
    else if (dtype == kF16) maxpool3x3s2_fwd_kernel<__half><<<pool_grid(work), 256, 0, st>>>((const __half*)x, (__half*)y, idx, N, H, W, C, OH, OW);
    else throw std::runtime_error("kungfu_b200 maxpool: bf16/fp16 only");
    count_launch();
    KF_CUDA_CHECK(cudaGetLastError());
}

void launch_maxpool3x3s2_bwd(const void* dy, const uint8_t* idx, void* dx, int dtype, int N, int H, int W, int C, cudaStream_t st) {
    if (C % 8) throw std::runtime_error("kungfu_b200 maxpool: C must be a multiple of 8");
    const int OH = (H + 2 - 3) / 2 + 1, OW = (W + 2 - 3) / 2 + 1;
    const size_t work = (size_t)N * H * W * (C / 8);
    if (dtype == kBF16) maxpool3x3s2_bwd_kernel<__nv_bfloat16><<<pool_grid(work), 256, 0, st>>>((const __nv_bfloat16*)dy, idx, (__nv_bfloat16*)dx, N, H, W, C, OH, OW);
    else if (dtype == kF16) maxpool3x3s2_bwd_kernel<__half><<<pool_grid(work), 256, 0, st>>>((const __half*)dy, idx, (__half*)dx, N, H, W, C, OH, OW);
    else throw std::runtime_error("kungfu_b200 maxpool: bf16/fp16 only");
    count_launch();
    KF_CUDA_CHECK(cudaGetLastError());
}

} // namespace kf
