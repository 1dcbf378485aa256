#include "reduce.hpp"

#include <cstring>
#include <type_traits>

#include "base.hpp"

#if defined(__x86_64__)
#include <immintrin.h>
#endif

namespace kfrt {

size_t dtype_size(int dt) {
    switch (dt) {
    case kF32: case kI32: case kU32: return 4;
    case kBF16: case kF16: case kI16: case kU16: return 2;
    case kI64: case kF64: case kU64: return 8;
    case kU8: case kI8: return 1;
    default: fail("unknown dtype code " + std::to_string(dt));
    }
}
const char* dtype_name(int dt) {
    static const char* n[] = {"f32", "bf16", "f16", "i32", "i64", "u8", "f64", "i8", "u32", "u64", "i16", "u16"};
    return dt >= 0 && dt < 12 ? n[dt] : "?";
}
const char* op_name(int op) {
    static const char* n[] = {"sum", "min", "max", "prod"};
    return op >= 0 && op < 4 ? n[op] : "?";
}

// ---- 16-bit float conversions -----------------------------------------------------------------------
float bf16_to_float(uint16_t h) {
    uint32_t u = (uint32_t)h << 16;
    float f;
    memcpy(&f, &u, 4);
    return f;
}
uint16_t float_to_bf16(float f) {
    uint32_t u;
    memcpy(&u, &f, 4);
    if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40); // NaN
    u += 0x7fffu + ((u >> 16) & 1); // round to nearest even
    return (uint16_t)(u >> 16);
}
float half_to_float(uint16_t h) {
    const uint32_t sign = (uint32_t)(h & 0x8000) << 16;
    uint32_t exp = (h >> 10) & 0x1f, man = h & 0x3ff, u;
    if (exp == 0) {
        if (man == 0) {
            u = sign;
        } else { // subnormal
            int e = -1;
            do {
                ++e;
                man <<= 1;
            } while (!(man & 0x400));
            u = sign | ((uint32_t)(127 - 15 - e) << 23) | ((man & 0x3ff) << 13);
        }
    } else if (exp == 31) {
        u = sign | 0x7f800000u | (man << 13);
    } else {
        u = sign | ((exp + 112) << 23) | (man << 13);
    }
    float f;
    memcpy(&f, &u, 4);
    return f;
}
uint16_t float_to_half(float f) {
    uint32_t u;
    memcpy(&u, &f, 4);
    const uint16_t sign = (uint16_t)((u >> 16) & 0x8000);
    const uint32_t absu = u & 0x7fffffffu;
    if (absu >= 0x7f800000u) return (uint16_t)(sign | 0x7c00 | (absu > 0x7f800000u ? 0x200 : 0)); // inf / nan
    if (absu >= 0x477ff000u) return (uint16_t)(sign | 0x7c00); // overflow -> inf (after rounding)
    if (absu < 0x33000001u) return sign;                         // underflow -> 0
    int32_t exp = (int32_t)(absu >> 23) - 127;
    uint32_t man = (absu & 0x7fffffu) | 0x800000u;
    if (exp < -14) { // subnormal half
        const int shift = -14 - exp + 13;
        uint32_t r = man >> shift;
        const uint32_t rem = man & ((1u << shift) - 1), half = 1u << (shift - 1);
        if (rem > half || (rem == half && (r & 1))) ++r;
        return (uint16_t)(sign | r);
    }
    uint32_t r = ((uint32_t)(exp + 15) << 10) | ((man >> 13) & 0x3ff);
    const uint32_t rem = man & 0x1fff;
    if (rem > 0x1000 || (rem == 0x1000 && (r & 1))) ++r;
    return (uint16_t)(sign | r);
}

// ---- generic element-wise kernels -------------------------------------------------------------------
template <typename T, int OP> static inline T apply(T a, T b) {
    if constexpr (OP == kSum) return (T)(a + b);
    else if constexpr (OP == kMin) return a < b ? a : b;
    else if constexpr (OP == kMax) return a > b ? a : b;
    else return (T)(a * b);
}

template <typename T, int OP> static void kernel(T* __restrict__ d, const T* x, const T* y, size_t n) {
    // plain loop: gcc -O2/-O3 vectorises it for every arithmetic type
    for (size_t i = 0; i < n; ++i) d[i] = apply<T, OP>(x[i], y[i]);
}

template <int OP, float (*UP)(uint16_t), uint16_t (*DOWN)(float)>
static void kernel16(uint16_t* d, const uint16_t* x, const uint16_t* y, size_t n) {
    for (size_t i = 0; i < n; ++i) d[i] = DOWN(apply<float, OP>(UP(x[i]), UP(y[i])));
}

#if defined(__x86_64__)
// fp16 sum, 8 lanes at a time through F16C (what the reference hand-writes in f16.c)
__attribute__((target("avx,f16c"))) static void f16_sum_avx(uint16_t* d, const uint16_t* x, const uint16_t* y, size_t n) {
    size_t i = 0;
    for (; i + 8 <= n; i += 8) {
        const __m256 a = _mm256_cvtph_ps(_mm_loadu_si128(reinterpret_cast<const __m128i*>(x + i)));
        const __m256 b = _mm256_cvtph_ps(_mm_loadu_si128(reinterpret_cast<const __m128i*>(y + i)));
        _mm_storeu_si128(reinterpret_cast<__m128i*>(d + i), _mm256_cvtps_ph(_mm256_add_ps(a, b), _MM_FROUND_TO_NEAREST_INT | _MM_FROUND_NO_EXC));
    }
    for (; i < n; ++i) d[i] = float_to_half(half_to_float(x[i]) + half_to_float(y[i]));
}
// bf16 sum: widen by shifting, add in fp32, round-to-nearest-even back
__attribute__((target("avx2"))) static void bf16_sum_avx2(uint16_t* d, const uint16_t* x, const uint16_t* y, size_t n) {
    size_t i = 0;
    const __m256i bias = _mm256_set1_epi32(0x7fff), one = _mm256_set1_epi32(1);
    for (; i + 8 <= n; i += 8) {
        const __m256i xa = _mm256_slli_epi32(_mm256_cvtepu16_epi32(_mm_loadu_si128(reinterpret_cast<const __m128i*>(x + i))), 16);
        const __m256i ya = _mm256_slli_epi32(_mm256_cvtepu16_epi32(_mm_loadu_si128(reinterpret_cast<const __m128i*>(y + i))), 16);
        const __m256i s = _mm256_castps_si256(_mm256_add_ps(_mm256_castsi256_ps(xa), _mm256_castsi256_ps(ya)));
        const __m256i lsb = _mm256_and_si256(_mm256_srli_epi32(s, 16), one);
        const __m256i r = _mm256_srli_epi32(_mm256_add_epi32(s, _mm256_add_epi32(bias, lsb)), 16);
        const __m256i packed = _mm256_permute4x64_epi64(_mm256_packus_epi32(r, r), 0xD8);
        _mm_storeu_si128(reinterpret_cast<__m128i*>(d + i), _mm256_castsi256_si128(packed));
    }
    for (; i < n; ++i) d[i] = float_to_bf16(bf16_to_float(x[i]) + bf16_to_float(y[i]));
}
static bool cpu_has(const char* f) {
    __builtin_cpu_init();
    if (!strcmp(f, "f16c")) return __builtin_cpu_supports("avx") && __builtin_cpu_supports("f16c");
    if (!strcmp(f, "avx2")) return __builtin_cpu_supports("avx2");
    return false;
}
#endif

template <typename T> static void dispatch_op(void* d, const void* x, const void* y, size_t n, int op) {
    T* dd = static_cast<T*>(d);
    const T* xx = static_cast<const T*>(x);
    const T* yy = static_cast<const T*>(y);
    switch (op) {
    case kSum: kernel<T, kSum>(dd, xx, yy, n); break;
    case kMin: kernel<T, kMin>(dd, xx, yy, n); break;
    case kMax: kernel<T, kMax>(dd, xx, yy, n); break;
    case kProd: kernel<T, kProd>(dd, xx, yy, n); break;
    default: fail("unknown reduce op " + std::to_string(op));
    }
}

template <float (*UP)(uint16_t), uint16_t (*DOWN)(float)>
static void dispatch16(void* d, const void* x, const void* y, size_t n, int op) {
    uint16_t* dd = static_cast<uint16_t*>(d);
    const uint16_t* xx = static_cast<const uint16_t*>(x);
    const uint16_t* yy = static_cast<const uint16_t*>(y);
    switch (op) {
    case kSum: kernel16<kSum, UP, DOWN>(dd, xx, yy, n); break;
    case kMin: kernel16<kMin, UP, DOWN>(dd, xx, yy, n); break;
    case kMax: kernel16<kMax, UP, DOWN>(dd, xx, yy, n); break;
    case kProd: kernel16<kProd, UP, DOWN>(dd, xx, yy, n); break;
    default: fail("unknown reduce op " + std::to_string(op));
    }
}

void transform2(void* d, const void* x, const void* y, size_t n, int dtype, int op) {
    if (n == 0) return;
    switch (dtype) {
    case kF32: dispatch_op<float>(d, x, y, n, op); break;
    case kF64: dispatch_op<double>(d, x, y, n, op); break;
    case kI32: dispatch_op<int32_t>(d, x, y, n, op); break;
    case kI64: dispatch_op<int64_t>(d, x, y, n, op); break;
    case kU8: dispatch_op<uint8_t>(d, x, y, n, op); break;
    case kI8: dispatch_op<int8_t>(d, x, y, n, op); break;
    case kU32: dispatch_op<uint32_t>(d, x, y, n, op); break;
    case kU64: dispatch_op<uint64_t>(d, x, y, n, op); break;
    case kI16: dispatch_op<int16_t>(d, x, y, n, op); break;
    case kU16: dispatch_op<uint16_t>(d, x, y, n, op); break;
    case kF16:
#if defined(__x86_64__)
        if (op == kSum && cpu_has("f16c")) {
            f16_sum_avx(static_cast<uint16_t*>(d), static_cast<const uint16_t*>(x), static_cast<const uint16_t*>(y), n);
            break;
        }
#endif
        dispatch16<half_to_float, float_to_half>(d, x, y, n, op);
        break;
    case kBF16:
#if defined(__x86_64__)
        if (op == kSum && cpu_has("avx2")) {
            bf16_sum_avx2(static_cast<uint16_t*>(d), static_cast<const uint16_t*>(x), static_cast<const uint16_t*>(y), n);
            break;
        }
#endif
        dispatch16<bf16_to_float, float_to_bf16>(d, x, y, n, op);
        break;
    default: fail("unknown dtype code " + std::to_string(dtype));
    }
}

} // namespace kfrt
