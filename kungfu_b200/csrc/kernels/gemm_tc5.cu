// kungfu_b200 — hand-written tcgen05 GEMM for sm_100a:  Y[M,N] = act(X[M,K] · W[N,K]^T + bias[N]),  bf16 in / fp32 accumulate / bf16 out.
//
// No CUTLASS / CuTe: the mainloop, the descriptors and the epilogue below are written against the PTX ISA directly.
//
//   * persistent kernel, one CTA per SM, 192 threads = 6 warps with fixed roles:
//       warp 0  TMA producer   one elected lane issues `cp.async.bulk.tensor.2d` (128B-swizzled boxes) for the X and W tiles of
//                              every k-block into a 4-deep shared-memory ring, completion on `full[stage]` via complete_tx
//       warp 1  MMA issuer     allocates TMEM (512 columns = two 128x256 fp32 accumulators); one elected lane issues
//                              `tcgen05.mma.cta_group::1.kind::f16` 128x256x16 straight from shared memory (K-major SW128 smem
//                              descriptors, +32 B per k-step), `tcgen05.commit` hands the ring slot back (`empty[stage]`) and, after
//                              the last k-block, publishes the accumulator (`tmem_full[acc]`)
//       warps 2-5  epilogue    `tcgen05.ld.32x32b.x32` (each warp reads its own 32 TMEM lanes = 32 rows of the tile), bias +
//                              GELU in registers (erf by a 5-term rational form: one MUFU.RCP + one MUFU.EX2 per element), bf16
//                              pack, 16-byte row-contiguous global stores; then `tmem_empty[acc]` lets the MMA warp reuse the
//                              accumulator — the epilogue of tile i overlaps the mainloop of tile i+1 (two accumulators in TMEM)
//   * every mbarrier wait is bounded (trap instead of hanging the GPU if a descriptor is ever wrong)
//
// Shapes: K % 64 == 0 (one 128-byte swizzle atom per k-block row), N % 8 == 0 (16-byte stores); M, N edges are handled by TMA
// zero fill on loads and predicated stores. SASS: UTCHMMA / UTMALDG / LDTM / UTCBAR (profiles/sass/gemm_tc5.sass).
//
// Reference: none — the reference has no GEMM or any device code; this serves the BERT feed-forward
// up-projection (bias + GELU fused into the GEMM epilogue) of the PairAveraging configuration.
#include <cuda.h>

#include <map>
#include <mutex>
#include <tuple>

#include "common.cuh"
#include "kernels.h"

namespace kf {

namespace {

constexpr int kBM = 128, kBN = 256, kBK = 64, kStages = 4, kUmmaK = 16;
constexpr int kThreads = 192;
constexpr uint32_t kABytes = kBM * kBK * 2;          // 16 KiB
constexpr uint32_t kBBytes = kBN * kBK * 2;          // 32 KiB
constexpr uint32_t kStageBytes = kABytes + kBBytes;  // 48 KiB
constexpr uint32_t kBiasBytes = 2 * kBN * 4;          // the bias of the current / next tile as fp32, staged once per tile
constexpr uint32_t kSmemBytes = kStages * kStageBytes + 1024 /* alignment slack */ + 256 /* barriers */ + kBiasBytes;
constexpr uint32_t kTmemCols = 512;                  // 2 accumulators x 256 fp32 columns

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// bounded wait: a wrong descriptor / byte count must end in a trap (an error the host sees), never in a hung GPU
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    unsigned long long t0 = 0;
    for (uint32_t it = 0;; ++it) {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (done) return;
        if ((it & 0xfff) == 0xfff) {
            unsigned long long now;
            asm volatile("mov.u64 %0, %globaltimer;" : "=l"(now));
            if (t0 == 0) t0 = now;
            else if (now - t0 > 4000000000ull) __trap(); // 4 s: something is structurally wrong
        }
    }
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "elect.sync _|P1, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, P1;\n\t"
        "}"
        : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tma_load_2d(uint32_t smem_dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(smem_dst),
                 "l"(map), "r"(c0), "r"(c1), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
// K-major operand tile in shared memory, 128-byte swizzle: rows of 128 B, 8-row groups of 1024 B (SBO), LBO unused (1)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);  // start address, 16-byte units, bits [0,14)
    d |= (uint64_t)1 << 16;                        // leading byte offset (ignored for swizzled K-major), bits [16,30)
    d |= (uint64_t)(1024 >> 4) << 32;              // stride byte offset between 8-row groups, bits [32,46)
    d |= (uint64_t)1 << 46;                        // descriptor version 1 (Blackwell), bits [46,48)
    d |= (uint64_t)2 << 61;                        // layout type SWIZZLE_128B, bits [61,64)
    return d;
}
// kind::f16 instruction descriptor: D = f32, A = B = bf16, both K-major, N >> 3 at [17,23), M >> 4 at [24,29)
constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kBN >> 3) << 17) | ((uint32_t)(kBM >> 4) << 24);

__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
          "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
          "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]),
          "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// erf(x) by Abramowitz-Stegun 7.1.26 (|error| <= 1.5e-7): one reciprocal, one exponential, five FMAs
__device__ __forceinline__ float erf_fast(float x) {
    const float ax = fabsf(x);
    const float t = __fdividef(1.f, fmaf(0.3275911f, ax, 1.f)); // MUFU.RCP, no IEEE slow path: |error| stays far below bf16 resolution
    float p = fmaf(1.061405429f, t, -1.453152027f);
    p = fmaf(p, t, 1.421413741f);
    p = fmaf(p, t, -0.284496736f);
    p = fmaf(p, t, 0.254829592f);
    const float e = 1.f - p * t * __expf(-ax * ax);
    return copysignf(e, x);
}
__device__ __forceinline__ float gelu_fast(float x) { return 0.5f * x * (1.f + erf_fast(x * 0.70710678118654752f)); }

struct GemmArgs {
    __nv_bfloat16* y;
    const __nv_bfloat16* bias; // may be null
    int M, N, K;
    int act; // 0 identity, 1 GELU
    int tiles_m, tiles_n;
};

__global__ void __launch_bounds__(kThreads, 1) gemm_tc5_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w,
                                                                const __grid_constant__ GemmArgs a) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t tiles = (raw + 1023u) & ~1023u; // SWIZZLE_128B tiles need 1024-byte alignment
    const uint32_t bars = tiles + kStages * kStageBytes;
    // barrier words (8 bytes each): full[kStages], empty[kStages], tmem_full[2], tmem_empty[2]; then the TMEM base address
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 8u * (kStages + s); };
    auto tfull_bar = [&](int s) { return bars + 8u * (2 * kStages + s); };
    auto tempty_bar = [&](int s) { return bars + 8u * (2 * kStages + 2 + s); };
    const uint32_t tmem_slot = bars + 8u * (2 * kStages + 4);
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - raw));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_kb = a.K / kBK;
    const int num_tiles = a.tiles_m * a.tiles_n;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&map_x);
        prefetch_tmap(&map_w);
        for (int s = 0; s < kStages; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(tfull_bar(s), 1);
            mbar_init(tempty_bar(s), 4); // one arrival per epilogue warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) { // TMEM allocation is a warp-wide instruction; the same warp frees it at the end
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    if (warp == 0) {
        // ================= TMA producer =================
        if (elect_one()) {
            uint32_t stage = 0, phase = 0;
            for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
                const int m0 = (t % a.tiles_m) * kBM, n0 = (t / a.tiles_m) * kBN;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(empty_bar(stage), phase ^ 1);
                    const uint32_t sa = tiles + stage * kStageBytes, sb = sa + kABytes;
                    mbar_expect_tx(full_bar(stage), kStageBytes);
                    tma_load_2d(sa, &map_x, kb * kBK, m0, full_bar(stage));
                    tma_load_2d(sb, &map_w, kb * kBK, n0, full_bar(stage));
                    if (++stage == kStages) stage = 0, phase ^= 1;
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        if (elect_one()) {
            uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
            for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
                mbar_wait(tempty_bar(acc), acc_phase ^ 1); // the epilogue has drained this accumulator
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * kBN;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(full_bar(stage), phase); // the TMA bytes of this stage have landed
                    tc_fence_after();
                    const uint32_t sa = tiles + stage * kStageBytes, sb = sa + kABytes;
                    const uint64_t da = make_smem_desc(sa), db = make_smem_desc(sb);
#pragma unroll
                    for (int k = 0; k < kBK / kUmmaK; ++k) {
                        // +32 bytes along K inside the 128-byte swizzle atom = +2 in the descriptor's 16-byte address units
                        umma_f16(d_tmem, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), kIdesc, (kb | k) != 0 ? 1u : 0u);
                    }
                    umma_commit(empty_bar(stage)); // frees the smem slot once the MMAs that read it are done
                    if (kb == num_kb - 1) umma_commit(tfull_bar(acc)); // accumulator complete
                    if (++stage == kStages) stage = 0, phase ^= 1;
                }
                if (++acc == 2) acc = 0, acc_phase ^= 1;
            }
        }
    } else {
        // ================= epilogue: TMEM -> registers -> bias/GELU -> bf16 -> global =================
        const int q = warp & 3;            // the TMEM lane quarter this warp may read: lanes [32 q, 32 q + 32)
        const int et = threadIdx.x - 64;   // 0..127 among the epilogue threads
        float* sbias = reinterpret_cast<float*>(smem_raw + (bars + 256 - raw));
        uint32_t acc = 0, acc_phase = 0;
        // one 32-column chunk: bias (shared memory, staged per tile) + activation + bf16 pack + 16-byte row-contiguous stores
        auto emit = [&](const uint32_t* r, int c, int row, int n0, const float* sb, __nv_bfloat16* yrow) {
            if (row >= a.M || n0 + c >= a.N) return;
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
                if (n0 + c + j < a.N) { // N % 8 == 0: a 16-byte chunk is entirely inside or outside
                    float v[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[j + i]) + sb[c + j + i];
                    if (a.act == 1) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) v[i] = gelu_fast(v[i]);
                    }
                    uint4 o;
                    uint32_t* ow = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
                        ow[i] = *reinterpret_cast<const uint32_t*>(&h);
                    }
                    *reinterpret_cast<uint4*>(yrow + c + j) = o;
                }
            }
        };
        for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
            const int m0 = (t % a.tiles_m) * kBM, n0 = (t / a.tiles_m) * kBN;
            // the tile's bias, once, as fp32 in shared memory (a global load per 8 columns inside the loop was the epilogue's
            // critical path: 32 dependent L2 round trips per tile, ncu: long-scoreboard stalls on the bias unpack)
            float* sb = sbias + acc * kBN;
            for (int i = et; i < kBN; i += 128) sb[i] = (a.bias && n0 + i < a.N) ? __bfloat162float(a.bias[n0 + i]) : 0.f;
            asm volatile("bar.sync 1, 128;" ::: "memory"); // the four epilogue warps only
            mbar_wait(tfull_bar(acc), acc_phase);
            tc_fence_after();
            const int row = m0 + q * 32 + lane;
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * kBN;
            __nv_bfloat16* yrow = a.y + (size_t)row * a.N + n0;
            // software pipeline over the 8 chunks: the TMEM load of chunk c+1 is in flight while chunk c is computed
            uint32_t ra[32], rb[32];
            tmem_ld32(taddr, ra);
#pragma unroll 1
            for (int c = 0; c < kBN; c += 64) {
                tmem_ld_wait();
                tmem_ld32(taddr + c + 32, rb);
                emit(ra, c, row, n0, sb, yrow);
                tmem_ld_wait();
                if (c + 64 < kBN) tmem_ld32(taddr + c + 64, ra);
                emit(rb, c + 32, row, n0, sb, yrow);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar(acc));
            if (++acc == 2) acc = 0, acc_phase ^= 1;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
    }
}

// ---- host side: tensor maps -------------------------------------------------------------------------------------------
using EncodeFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                              const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeFn encode_fn() {
    static EncodeFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
        return reinterpret_cast<EncodeFn>(p);
    }();
    return fn;
}

// row-major [rows, K] bf16 matrix, boxes of [box_rows, 64] elements, 128-byte swizzle, zero fill out of bounds
CUtensorMap make_map(const void* base, int rows, int K, int box_rows) {
    CUtensorMap m;
    const cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)K * 2};
    const cuuint32_t box[2] = {(cuuint32_t)kBK, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = encode_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) throw std::runtime_error("kungfu_b200: cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
    return m;
}

} // namespace

bool gemm_sm100_available() { return encode_fn() != nullptr; }

size_t gemm_bias_act_workspace(int, int, int, int) { return 0; }

void launch_gemm_bias_act(const void* x, const void* w, const void* bias, void* y, int tokens, int in_f, int out_f, int act, void*, size_t,
                          cudaStream_t st) {
    if (!gemm_sm100_available()) throw std::runtime_error("kungfu_b200: cuTensorMapEncodeTiled is not available in this driver");
    if (in_f % kBK != 0 || in_f < kBK) throw std::runtime_error("kungfu_b200: tcgen05 GEMM needs in_features % 64 == 0");
    if (out_f % 8 != 0) throw std::runtime_error("kungfu_b200: tcgen05 GEMM needs out_features % 8 == 0");
    if (((uintptr_t)x | (uintptr_t)w | (uintptr_t)y | (uintptr_t)bias) & 15) throw std::runtime_error("kungfu_b200: tcgen05 GEMM needs 16-byte aligned tensors");
    static std::mutex mu;
    static std::map<std::tuple<const void*, int, int, int>, CUtensorMap> cache; // tensor maps are pure functions of (ptr, shape, box)
    auto get = [&](const void* p, int rows, int box) {
        std::lock_guard<std::mutex> g(mu);
        const auto key = std::make_tuple(p, rows, in_f, box);
        auto it = cache.find(key);
        if (it == cache.end()) {
            if (cache.size() > 4096) cache.clear();
            it = cache.emplace(key, make_map(p, rows, in_f, box)).first;
        }
        return it->second;
    };
    const CUtensorMap mx = get(x, tokens, kBM), mw = get(w, out_f, kBN);
    GemmArgs a;
    a.y = reinterpret_cast<__nv_bfloat16*>(y);
    a.bias = reinterpret_cast<const __nv_bfloat16*>(bias);
    a.M = tokens;
    a.N = out_f;
    a.K = in_f;
    a.act = act;
    a.tiles_m = (tokens + kBM - 1) / kBM;
    a.tiles_n = (out_f + kBN - 1) / kBN;
    static int sms = 0;
    if (sms == 0) {
        int dev = 0;
        KF_CUDA_CHECK(cudaGetDevice(&dev));
        KF_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        KF_CUDA_CHECK(cudaFuncSetAttribute(gemm_tc5_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
    }
    const int tiles = a.tiles_m * a.tiles_n;
    const int grid = tiles < sms ? tiles : sms;
    gemm_tc5_kernel<<<grid, kThreads, kSmemBytes, st>>>(mx, mw, a);
    count_launch();
    KF_CUDA_CHECK(cudaGetLastError());
}

} // namespace kf
