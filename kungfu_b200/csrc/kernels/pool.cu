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
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
        const int cv = (int)(t % cvn);
        size_t p = t / cvn;
        const int ow = (int)(p % OW);
        p /= OW;
        const int oh = (int)(p % OH);
        const int n = (int)(p / OH);
        float best[8];
        uint32_t arg[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            best[i] = -INFINITY;
            arg[i] = 0;
        }
#pragma unroll
        for (int kh = 0; kh < 3; ++kh) {
            const int ih = oh * 2 - 1 + kh;
            if (ih < 0 || ih >= H) continue;
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) {
                const int iw = ow * 2 - 1 + kw;
                if (iw < 0 || iw >= W) continue;
                const AccVec<T> v = unpack<T>(ld_stream(x + (((size_t)n * H + ih) * W + iw) * C + (size_t)cv * 8));
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float f = v.a[i];
                    if (f > best[i] || f != f) { // NaN propagates, first maximum wins (PyTorch semantics)
                        best[i] = f;
                        arg[i] = (uint32_t)(kh * 3 + kw);
                    }
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
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
        const int cv = (int)(t % cvn);
        size_t p = t / cvn;
        const int iw = (int)(p % W);
        p /= W;
        const int ih = (int)(p % H);
        const int n = (int)(p / H);
        float acc[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = 0.f;
        // output windows covering (ih, iw): oh in {floor((ih+1)/2) - 1 .. floor((ih+1)/2)} with 0 <= ih - (2 oh - 1) <= 2
        const int oh1 = (ih + 1) >> 1, ow1 = (iw + 1) >> 1;
#pragma unroll
        for (int a = 0; a < 2; ++a) {
            const int oh = oh1 - a;
            const int kh = ih - (oh * 2 - 1);
            if (oh < 0 || oh >= OH || kh < 0 || kh > 2) continue;
#pragma unroll
            for (int b = 0; b < 2; ++b) {
                const int ow = ow1 - b;
                const int kw = iw - (ow * 2 - 1);
                if (ow < 0 || ow >= OW || kw < 0 || kw > 2) continue;
                const size_t ov = (((size_t)n * OH + oh) * OW + ow) * cvn + cv;
                const uint64_t packed = *reinterpret_cast<const uint64_t*>(idx + ov * 8);
                const AccVec<T> g = unpack<T>(ld_stream(dy + ov * 8));
                const uint32_t me = (uint32_t)(kh * 3 + kw);
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    if (((packed >> (8 * i)) & 0xff) == me) acc[i] += (float)g.a[i];
            }
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
