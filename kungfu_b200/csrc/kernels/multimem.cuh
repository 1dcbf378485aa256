// kungfu_b200 — PTX wrappers for NVLS (NVSwitch in-fabric reduction / multicast).
//
// `multimem.ld_reduce` on a multicast address returns the reduction of that address across every
// GPU bound to the multicast object (the switch does the fan-in); `multimem.st` writes the value to
// every bound GPU (fan-out). This is the large-message path of K1 in SURVEY.md §2.3 — there is no
// counterpart in the reference (its GPU path is `ncclAllReduce`, srcs/cpp/src/nccl/
// gpu_collective.cpp:146-166).
#pragma once
#include "common.cuh"

namespace kf {

template <typename T, int OP> struct Multimem {
    static constexpr bool kSupported = false;
    __device__ static __forceinline__ Vec16 ld_reduce(const void*) { return Vec16{}; }
};

// fp32 sum: exact fp32 adds in the switch.
template <> struct Multimem<float, kSum> {
    static constexpr bool kSupported = true;
    __device__ static __forceinline__ Vec16 ld_reduce(const void* mc) {
        Vec16 v;
        asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
                     : "=r"(v.w[0]), "=r"(v.w[1]), "=r"(v.w[2]), "=r"(v.w[3])
                     : "l"(mc)
                     : "memory");
        return v;
    }
};
// bf16: fp32 accumulation inside the switch (.acc::f32), bf16x2-packed result.
template <> struct Multimem<__nv_bfloat16, kSum> {
    static constexpr bool kSupported = true;
    __device__ static __forceinline__ Vec16 ld_reduce(const void* mc) {
        Vec16 v;
        asm volatile("multimem.ld_reduce.relaxed.sys.global.add.acc::f32.v4.bf16x2 {%0,%1,%2,%3}, [%4];"
                     : "=r"(v.w[0]), "=r"(v.w[1]), "=r"(v.w[2]), "=r"(v.w[3])
                     : "l"(mc)
                     : "memory");
        return v;
    }
};
template <> struct Multimem<__nv_bfloat16, kMin> {
    static constexpr bool kSupported = true;
    __device__ static __forceinline__ Vec16 ld_reduce(const void* mc) {
        Vec16 v;
        asm volatile("multimem.ld_reduce.relaxed.sys.global.min.v4.bf16x2 {%0,%1,%2,%3}, [%4];"
                     : "=r"(v.w[0]), "=r"(v.w[1]), "=r"(v.w[2]), "=r"(v.w[3])
                     : "l"(mc)
                     : "memory");
        return v;
    }
};
template <> struct Multimem<__nv_bfloat16, kMax> {
    static constexpr bool kSupported = true;
    __device__ static __forceinline__ Vec16 ld_reduce(const void* mc) {
        Vec16 v;
        asm volatile("multimem.ld_reduce.relaxed.sys.global.max.v4.bf16x2 {%0,%1,%2,%3}, [%4];"
                     : "=r"(v.w[0]), "=r"(v.w[1]), "=r"(v.w[2]), "=r"(v.w[3])
                     : "l"(mc)
                     : "memory");
        return v;
    }
};
template <> struct Multimem<__half, kSum> {
    static constexpr bool kSupported = true;
    __device__ static __forceinline__ Vec16 ld_reduce(const void* mc) {
        Vec16 v;
        asm volatile("multimem.ld_reduce.relaxed.sys.global.add.acc::f32.v4.f16x2 {%0,%1,%2,%3}, [%4];"
                     : "=r"(v.w[0]), "=r"(v.w[1]), "=r"(v.w[2]), "=r"(v.w[3])
                     : "l"(mc)
                     : "memory");
        return v;
    }
};
template <> struct Multimem<__half, kMin> {
    static constexpr bool kSupported = true;
    __device__ static __forceinline__ Vec16 ld_reduce(const void* mc) {
        Vec16 v;
        asm volatile("multimem.ld_reduce.relaxed.sys.global.min.v4.f16x2 {%0,%1,%2,%3}, [%4];"
                     : "=r"(v.w[0]), "=r"(v.w[1]), "=r"(v.w[2]), "=r"(v.w[3])
                     : "l"(mc)
                     : "memory");
        return v;
    }
};
template <> struct Multimem<__half, kMax> {
    static constexpr bool kSupported = true;
    __device__ static __forceinline__ Vec16 ld_reduce(const void* mc) {
        Vec16 v;
        asm volatile("multimem.ld_reduce.relaxed.sys.global.max.v4.f16x2 {%0,%1,%2,%3}, [%4];"
                     : "=r"(v.w[0]), "=r"(v.w[1]), "=r"(v.w[2]), "=r"(v.w[3])
                     : "l"(mc)
                     : "memory");
        return v;
    }
};

// Fan-out store of one 16-byte vector to every GPU of the multicast group. The payload is opaque
// bits, so a single f32x4 form serves every dtype.
__device__ __forceinline__ void multimem_st(void* mc, const Vec16& v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(mc), "r"(v.w[0]), "r"(v.w[1]),
                 "r"(v.w[2]), "r"(v.w[3])
                 : "memory");
}

} // namespace kf
