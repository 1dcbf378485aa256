// kungfu_b200 — tcgen05 GEMM with a fused bias (+ GELU) epilogue for the transformer MLP of the BERT family.
//
//     Y[tokens, out] = act(X[tokens, in] · W[out, in]^T + b[out])        bf16 in / bf16 out, fp32 accumulate
//
// Built from CUTLASS 4.x / CuTe sm100 collectives inside our own translation unit (the vendored header tree under
// flashinfer/data/cutlass/include): TMA (`cp.async.bulk.tensor`) stages A/B tiles in shared memory, one elected thread
// issues `tcgen05.mma` with the accumulator in TMEM, `cta_group::2` pairs two SMs on a 256x128 tile, the epilogue reads
// TMEM back with `tcgen05.ld`, applies bias + GELU in registers and TMA-stores bf16 — the activation never makes an
// extra HBM round trip. SASS evidence: UTCHMMA / UTMALDG / LDTM (see profiles/).
//
// The problem is evaluated transposed (D^T = W · X^T) so that the bias is a per-ROW vector of the kernel's output tile,
// which is the broadcast the epilogue fusion supports natively; D^T stored column-major is exactly Y row-major.
//
// No counterpart in the reference (it ships no model or GEMM code); this belongs to the model-family layer.
#include "kernels.h"

#if defined(KF_WITH_CUTLASS)

#include "cute/tensor.hpp"
#include "cutlass/cutlass.h"
#include "cutlass/epilogue/collective/collective_builder.hpp"
#include "cutlass/epilogue/fusion/operations.hpp"
#include "cutlass/epilogue/thread/activation.h"
#include "cutlass/gemm/collective/collective_builder.hpp"
#include "cutlass/gemm/device/gemm_universal_adapter.h"
#include "cutlass/gemm/kernel/gemm_universal.hpp"
#include "cutlass/util/packed_stride.hpp"

namespace kf {

namespace {

using namespace cute;

using ElementA = cutlass::bfloat16_t; // W   [out, in]   row-major
using ElementB = cutlass::bfloat16_t; // X^T [in, tokens] column-major  (= X row-major)
using ElementC = cutlass::bfloat16_t;
using ElementD = cutlass::bfloat16_t; // Y^T [out, tokens] column-major (= Y row-major)
using ElementAcc = float;
using ElementCompute = float;
using ElementBias = cutlass::bfloat16_t;
using LayoutA = cutlass::layout::RowMajor;
using LayoutB = cutlass::layout::ColumnMajor;
using LayoutC = cutlass::layout::ColumnMajor;
using LayoutD = cutlass::layout::ColumnMajor;
constexpr int kAlign = 8; // 16 bytes of bf16

using MmaTileShape = Shape<_256, _128, _64>; // per CTA pair (cta_group::2)
using ClusterShape = Shape<_2, _1, _1>;

template <template <class> class Act> struct GemmFor {
    using FusionOp = cutlass::epilogue::fusion::LinCombPerRowBiasEltAct<Act, ElementD, ElementCompute, ElementBias, ElementC, ElementCompute, kAlign>;
    using CollectiveEpilogue = typename cutlass::epilogue::collective::CollectiveBuilder<
        cutlass::arch::Sm100, cutlass::arch::OpClassTensorOp, MmaTileShape, ClusterShape, cutlass::epilogue::collective::EpilogueTileAuto,
        ElementAcc, ElementCompute, ElementC, LayoutC, kAlign, ElementD, LayoutD, kAlign, cutlass::epilogue::collective::EpilogueScheduleAuto,
        FusionOp>::CollectiveOp;
    using CollectiveMainloop = typename cutlass::gemm::collective::CollectiveBuilder<
        cutlass::arch::Sm100, cutlass::arch::OpClassTensorOp, ElementA, LayoutA, kAlign, ElementB, LayoutB, kAlign, ElementAcc, MmaTileShape,
        ClusterShape, cutlass::gemm::collective::StageCountAutoCarveout<static_cast<int>(sizeof(typename CollectiveEpilogue::SharedStorage))>,
        cutlass::gemm::collective::KernelScheduleAuto>::CollectiveOp;
    using GemmKernel = cutlass::gemm::kernel::GemmUniversal<Shape<int, int, int, int>, CollectiveMainloop, CollectiveEpilogue, void>;
    using Gemm = cutlass::gemm::device::GemmUniversalAdapter<GemmKernel>;
};

template <template <class> class Act>
size_t run_gemm(const void* x, const void* w, const void* bias, void* y, int tokens, int in_f, int out_f, void* workspace, size_t workspace_bytes,
                cudaStream_t st, bool query_only) {
    using G = GemmFor<Act>;
    using Gemm = typename G::Gemm;
    using StrideA = typename Gemm::GemmKernel::StrideA;
    using StrideB = typename Gemm::GemmKernel::StrideB;
    using StrideC = typename Gemm::GemmKernel::StrideC;
    using StrideD = typename Gemm::GemmKernel::StrideD;
    const int M = out_f, N = tokens, K = in_f;
    StrideA sa = cutlass::make_cute_packed_stride(StrideA{}, make_shape(M, K, 1));
    StrideB sb = cutlass::make_cute_packed_stride(StrideB{}, make_shape(N, K, 1));
    StrideC sc = cutlass::make_cute_packed_stride(StrideC{}, make_shape(M, N, 1));
    StrideD sd = cutlass::make_cute_packed_stride(StrideD{}, make_shape(M, N, 1));
    typename Gemm::Arguments args{cutlass::gemm::GemmUniversalMode::kGemm,
                                  {M, N, K, 1},
                                  {static_cast<const ElementA*>(w), sa, static_cast<const ElementB*>(x), sb},
                                  {{}, static_cast<const ElementC*>(y), sc, static_cast<ElementD*>(y), sd}};
    args.epilogue.thread.alpha = 1.0f;
    args.epilogue.thread.beta = 0.0f;
    args.epilogue.thread.bias_ptr = static_cast<const ElementBias*>(bias);
    Gemm gemm;
    const size_t need = Gemm::get_workspace_size(args);
    if (query_only) return need;
    if (need > workspace_bytes) throw std::runtime_error("kungfu_b200 gemm: workspace too small");
    if (gemm.can_implement(args) != cutlass::Status::kSuccess) throw std::runtime_error("kungfu_b200 gemm: unsupported problem shape (need in/out/tokens % 8 == 0)");
    if (gemm.initialize(args, workspace, st) != cutlass::Status::kSuccess) throw std::runtime_error("kungfu_b200 gemm: initialize failed");
    if (gemm.run(st) != cutlass::Status::kSuccess) throw std::runtime_error("kungfu_b200 gemm: launch failed");
    count_launch();
    return need;
}

} // namespace

bool gemm_sm100_available() { return true; }

size_t gemm_bias_act_workspace(int tokens, int in_f, int out_f, int act) {
    if (act == 1) return run_gemm<cutlass::epilogue::thread::GELU>(nullptr, nullptr, nullptr, nullptr, tokens, in_f, out_f, nullptr, 0, nullptr, true);
    return run_gemm<cutlass::epilogue::thread::Identity>(nullptr, nullptr, nullptr, nullptr, tokens, in_f, out_f, nullptr, 0, nullptr, true);
}

void launch_gemm_bias_act(const void* x, const void* w, const void* bias, void* y, int tokens, int in_f, int out_f, int act, void* workspace,
                          size_t workspace_bytes, cudaStream_t st) {
    if (tokens % 8 || in_f % 8 || out_f % 8) throw std::runtime_error("kungfu_b200 gemm: tokens, in and out features must be multiples of 8");
    if (act == 1) run_gemm<cutlass::epilogue::thread::GELU>(x, w, bias, y, tokens, in_f, out_f, workspace, workspace_bytes, st, false);
    else run_gemm<cutlass::epilogue::thread::Identity>(x, w, bias, y, tokens, in_f, out_f, workspace, workspace_bytes, st, false);
}

} // namespace kf

#else // built without the CUTLASS header tree

namespace kf {
bool gemm_sm100_available() { return false; }
size_t gemm_bias_act_workspace(int, int, int, int) { return 0; }
void launch_gemm_bias_act(const void*, const void*, const void*, void*, int, int, int, int, void*, size_t, cudaStream_t) {
    throw std::runtime_error("kungfu_b200: built without CUTLASS headers; tcgen05 GEMM unavailable");
}
} // namespace kf

#endif
