// kungfu_b200 — the optimizer update rule shared by the fused collective kernels and the local
// flat-arena optimizer. Semantics follow torch.optim.SGD / Adam / AdamW exactly so that wrapping a
// torch optimizer (reference API: kungfu.torch.optimizers.SynchronousSGDOptimizer, srcs/python/
// kungfu/torch/optimizers/sync_sgd.py:6-32) and taking the fused path are interchangeable.
#pragma once
#include "common.cuh"
#include "kernels.h"

namespace kf {

struct Hyper {
    float v[kHpCount];
};

__device__ __forceinline__ Hyper load_hyper(const float* hp) {
    Hyper h;
    const float4* p = reinterpret_cast<const float4*>(hp);
#pragma unroll
    for (int i = 0; i < kHpCount / 4; ++i) {
        const float4 q = __ldg(p + i);
        h.v[4 * i + 0] = q.x;
        h.v[4 * i + 1] = q.y;
        h.v[4 * i + 2] = q.z;
        h.v[4 * i + 3] = q.w;
    }
    return h;
}

// One scalar optimizer update. `g` must already be averaged/scaled. s1 = momentum buffer (SGD) or
// exp_avg (Adam); s2 = exp_avg_sq (Adam only).
template <int KIND> __device__ __forceinline__ void opt_update(float& w, float g, float& s1, float& s2, const Hyper& h) {
    const float lr = h.v[kHpLr];
    const float wd = h.v[kHpWeightDecay];
    if constexpr (KIND == kOptSgd) {
        const float mom = h.v[kHpMomentum];
        g = fmaf(wd, w, g);
        if (mom != 0.f) {
            s1 = (h.v[kHpFirstStep] != 0.f) ? g : fmaf(mom, s1, (1.f - h.v[kHpDampening]) * g);
            g = (h.v[kHpNesterov] != 0.f) ? fmaf(mom, s1, g) : s1;
        }
        w = fmaf(-lr, g, w);
    } else {
        const float b1 = h.v[kHpBeta1], b2 = h.v[kHpBeta2];
        if (h.v[kHpDecoupledWd] != 0.f) w *= (1.f - lr * wd);
        else g = fmaf(wd, w, g);
        s1 = fmaf(b1, s1, (1.f - b1) * g);
        s2 = fmaf(b2, s2, (1.f - b2) * g * g);
        const float denom = sqrtf(s2) / sqrtf(h.v[kHpBiasC2]) + h.v[kHpEps];
        w = fmaf(-(lr / h.v[kHpBiasC1]), s1 / denom, w);
    }
}

// Lane-wise float view of one 16-byte storage vector worth of fp32 state (4 lanes -> 1 vector,
// 8 lanes -> 2 vectors).
template <int LANES> struct F32Lanes {
    float f[LANES];
    __device__ __forceinline__ void load(const float* p) {
#pragma unroll
        for (int i = 0; i < LANES / 4; ++i) {
            const float4 q = *reinterpret_cast<const float4*>(p + 4 * i);
            f[4 * i + 0] = q.x;
            f[4 * i + 1] = q.y;
            f[4 * i + 2] = q.z;
            f[4 * i + 3] = q.w;
        }
    }
    __device__ __forceinline__ void store(float* p) const {
#pragma unroll
        for (int i = 0; i < LANES / 4; ++i)
            *reinterpret_cast<float4*>(p + 4 * i) = make_float4(f[4 * i + 0], f[4 * i + 1], f[4 * i + 2], f[4 * i + 3]);
    }
    __device__ __forceinline__ void zero() {
#pragma unroll
        for (int i = 0; i < LANES; ++i) f[i] = 0.f;
    }
};

} // namespace kf
