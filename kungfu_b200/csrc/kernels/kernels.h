// kungfu_b200 — host-callable launch API of the sm_100a kernel library (torch-free: raw pointers
// and cudaStream_t only, so the library builds in seconds and has no ABI coupling to libtorch).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include <stdexcept>
#include <string>

#include "mesh_args.h"

namespace kf {

enum Algo : int { kAlgoAuto = 0, kAlgoOneShot = 1, kAlgoTwoShot = 2, kAlgoNvls = 3 };
// kArGated: split-phase launch (a gate kernel precedes, a wait kernel follows): no flag barriers inside the kernel,
// its blocks signal completion on done word 0 and exit
// kArSumSqPost: with kArSumSq, only [1] (|result|^2) is accumulated — the caller already has sum_r |x_r|^2 (the gradient
// hand-over copy computes it on the way into the arena)
enum ArFlags : int { kArScale = 1, kArSumSq = 2, kArGated = 4, kArSumSqPost = 8 };
enum OptKind : int { kOptSgd = 0, kOptAdam = 1 };

// Device-resident optimizer hyper-parameters (16 floats). Kept in device memory so that a captured
// CUDA graph picks up learning-rate changes and the step counter without re-capture.
enum HyperSlot : int {
    kHpLr = 0, kHpMomentum = 1, kHpDampening = 2, kHpWeightDecay = 3, kHpNesterov = 4, kHpGradScale = 5,
    kHpFirstStep = 6, kHpBeta1 = 7, kHpBeta2 = 8, kHpEps = 9, kHpBiasC1 = 10, kHpBiasC2 = 11, kHpStep = 12,
    kHpAlpha = 13, kHpDecoupledWd = 14, kHpReserved = 15, kHpCount = 16
};

// ---- allreduce.cu
// `out_off`: arena offset the sliced algorithms deliver the result to on every rank (== off: in place)
void launch_allreduce(int algo, const MeshArgs& m, size_t off, void* out, size_t nbytes, int dtype, int op, float scale,
                      float* sumsq, int flags, int blocks, int threads, cudaStream_t st, size_t out_off = (size_t)-1);
bool nvls_supported(int dtype, int op);
// LL one-shot (flag-carrying pushes, no barrier): reads `in` / writes `out` (local tensors) directly, slot area in the arena
size_t allreduce_ll_slot_bytes(size_t max_payload_bytes);
void launch_allreduce_ll(const MeshArgs& m, const void* in, void* out, size_t nbytes, size_t slot_off, size_t max_payload_bytes, int dtype,
                         int op, float scale, int flags, cudaStream_t st);

// ---- fused_step.cu
// S-SGD: reduce-scatter(grads) -> optimizer on the owned shard (fp32 master/state) -> all-gather(params).
void launch_ssgd_fused(int algo, const MeshArgs& m, size_t grad_off, size_t param_off, int dtype, size_t nbytes,
                       float* master_shard, float* s1_shard, float* s2_shard, const float* hyper, int opt_kind,
                       float* sumsq, int flags, int blocks, int threads, cudaStream_t st);
// SMA: all-reduce(params)/np -> alpha blend -> local optimizer step with LOCAL gradients.
void launch_sma_fused(int algo, const MeshArgs& m, size_t param_off, size_t avg_off, int dtype, size_t nbytes,
                      const void* grads, int grad_dtype, float* s1, float* s2, void* model_out, int model_dtype,
                      const float* hyper, int opt_kind, int blocks, int threads, cudaStream_t st);
// SMA, overlapped form: the model average is produced by `launch_allreduce(param_off -> avg_off, scale 1/np)` while
// forward/backward run; this LOCAL kernel then does blend + optimizer step in one pass (phase 3 of sma_fused).
void launch_sma_blend(void* params, const void* avg, int dtype, size_t nbytes, const void* grads, int grad_dtype, float* s1, float* s2,
                      const float* hyper, int opt_kind, int blocks, int threads, cudaStream_t st);
// The published-model store: `nbuf` (2 or 3) buffers at pub_off[], version words at ver_off ([0] published, [1] being
// written). Version v lives in buffer v % nbuf; a reader of version v is torn iff the owner started writing v + nbuf.
struct PubStore {
    size_t pub_off[3];
    size_t ver_off;
    int nbuf;
};
// PairAveraging: v <- 0.5 (v + other), local optimizer step, publish the new model. `other` is the target's published
// model read over NVLink inside the kernel (pulled == nullptr), or a private copy pulled earlier on a side stream
// (pulled_ok[0] = version it holds, 0: none; pulled_ok[2] != 0: that copy is torn, skip the averaging). `target_dev`
// (pull, pull_tma, pair_avg): a device word holding the target rank, read by the kernel — one captured CUDA graph serves
// every target.
void launch_pair_avg(const MeshArgs& m, int target, const PubStore& ps, void* params, int dtype, size_t nbytes, const void* grads,
                     int grad_dtype, float* s1, float* s2, const float* hyper, int opt_kind, int do_step, uint32_t* done_counter,
                     const void* pulled, const uint32_t* pulled_ok, int blocks, int threads, cudaStream_t st,
                     const int* target_dev = nullptr, float* master = nullptr);
// publish(params) without averaging (step 0 of PairAveraging; `save_variable` on device)
void launch_publish(const MeshArgs& m, const PubStore& ps, const void* src, size_t nbytes, uint32_t* done_counter, int blocks,
                    int threads, cudaStream_t st);
// request(peer): copy the peer's latest published blob into dst. ok[0] = version read (0 = nothing published),
// ok[1] += torn reads, ok[2] = 1 when this copy is torn (ok[3] is scratch; `ok` is 4 words, zeroed by the caller before
// the first attempt). retry != 0: run only if ok[2] is set — a second attempt after a torn first one, followed by
// `launch_pull_settle`, which leaves ok[2] = "still torn".
void launch_pull(const MeshArgs& m, int target, const PubStore& ps, void* dst, size_t nbytes, uint32_t* ok, int retry, int blocks,
                 int threads, cudaStream_t st, const int* target_dev = nullptr);

void launch_pull_settle(uint32_t* ok, cudaStream_t st);

// ---- tma_copy.cu: the same pull driven by the TMA engine (cp.async.bulk, 4-stage smem pipeline, one thread per CTA)
void launch_pull_tma(const MeshArgs& m, int target, const PubStore& ps, void* dst, size_t nbytes, uint32_t* ok, int retry, int blocks,
                     cudaStream_t st, const int* target_dev = nullptr);

// ---- misc.cu
void launch_broadcast(int algo, const MeshArgs& m, size_t off, size_t nbytes, int root, int blocks, int threads,
                      cudaStream_t st);
void launch_allgather(int algo, const MeshArgs& m, size_t in_off, size_t out_off, size_t nbytes_per_rank, int blocks,
                      int threads, cudaStream_t st);
void launch_barrier(const MeshArgs& m, cudaStream_t st);
// split-phase collectives: one warp waits until `add` more block completions have arrived on done word `word`
void launch_wait_done(const MeshArgs& m, int word, uint32_t add, cudaStream_t st);

// ---- optim.cu (single-GPU building blocks)
void launch_hyper_tick(float* hyper, cudaStream_t st);
void launch_flat_opt(void* params, int dtype, const void* grads, int grad_dtype, float* master, float* s1, float* s2,
                     const float* hyper, int opt_kind, size_t nelem, cudaStream_t st);
void launch_cast(const void* src, int src_dtype, void* dst, int dst_dtype, size_t nelem, float scale, cudaStream_t st);
void launch_sumsq(const void* src, int dtype, size_t nelem, float* out, cudaStream_t st); // out[0] += |x|^2
void launch_sumsq_diff(const void* a, const void* b, int dtype, size_t nelem, float* out, cudaStream_t st);
void launch_axpby(void* y, const void* x, int dtype, size_t nelem, float a, float b, cudaStream_t st); // y = a*x + b*y
// gradient-noise-scale finalisation on device (reference: cpu/collective.cpp:212-258, ops/monitor.py:6-17)
//   in : sums[0] = sum_r |g_r|^2 (over ranks), sums[1] = |g_avg|^2
//   state: [0]=G_ema [1]=S_ema [2]=initialised flag; out[0]=noise scale, out[1]=G_biased, out[2]=S_biased
void launch_gns_finalize(const float* sums, float* state, float* out, float batch_small, float batch_big, float alpha,
                         float nranks, cudaStream_t st);
// uint8 NHWC image -> normalised bf16/fp32 NHWC (channels_last) tensor, fused cast+normalise
void launch_u8_normalize(const uint8_t* src, void* dst, int dst_dtype, size_t npix, const float* mean3,
                         const float* invstd3, cudaStream_t st);
void launch_fill_u32(uint32_t* p, uint32_t v, size_t n, cudaStream_t st);
// n (src -> dst, nbytes) copies in one launch per 48 tensors (gradient hand-over into the flat arena)
constexpr int kMaxCopyBatch = 48;
// sumsq != nullptr: sumsq[0] += sum of squares of everything copied, elements of `dtype` (f32 / bf16 / f16)
void launch_multi_copy(const void* const* src, void* const* dst, const size_t* nbytes, int n, cudaStream_t st, float* sumsq = nullptr, int dtype = 0);

// ---- bn_act.cu: fused BatchNorm (+residual) + ReLU on channels_last activations
int bn_num_blocks(size_t rows, int C);
// forward (training): partial is [bn_num_blocks][2][C] scratch; mean/invstd/scale/shift are [C] outputs; mask rows*C/8 bytes
void launch_bn_fwd(const void* x, const void* res, void* y, uint8_t* mask, int dtype, size_t rows, int C, const void* gamma, const void* beta,
                   void* running_mean, void* running_var, int pdtype, float eps, float momentum, int relu, float* partial, float* mean,
                   float* invstd, float* scale, float* shift, cudaStream_t st);
// forward with given scale/shift (eval mode, or the second half of training forward)
void launch_bn_apply(const void* x, const void* res, void* y, uint8_t* mask, int dtype, size_t rows, int C, const float* scale, const float* shift,
                     int relu, cudaStream_t st);
// backward: coef is [3][C] scratch; dres may be null
void launch_bn_bwd(const void* dy, const void* x, const uint8_t* mask, void* dx, void* dres, int dtype, size_t rows, int C, const void* gamma,
                   void* dgamma, void* dbeta, int pdtype, const float* mean, const float* invstd, int relu, float* partial, float* coef,
                   cudaStream_t st);

// ---- pool.cu: 3x3 stride-2 pad-1 max pooling, channels_last, 1-byte arg-max indices, gather backward
void launch_maxpool3x3s2_fwd(const void* x, void* y, uint8_t* idx, int dtype, int N, int H, int W, int C, cudaStream_t st);
void launch_maxpool3x3s2_bwd(const void* dy, const uint8_t* idx, void* dx, int dtype, int N, int H, int W, int C, cudaStream_t st);

// ---- transformer.cu: fused residual + LayerNorm and bias + GELU (forward / backward), bf16 / f16 activations
int add_ln_num_blocks(int rows);    // rows of the [blocks][2 or 3][H] fp32 scratch the backward needs
int bias_gelu_num_blocks(int rows); // rows of the [blocks][I] fp32 scratch the backward needs
int qkv_pack_num_blocks(int rows);  // rows of the [blocks][3H] fp32 scratch of qkv_pack_bwd
// z = x (+ lin_bias) + res (res / lin_bias may be null), y = LN(z) * gamma + beta; saves z (may be null: inference), mean, rstd
void launch_add_ln_fwd(const void* x, const void* res, const void* lin_bias, const void* gamma, const void* beta, void* z, void* y, float* mean,
                       float* rstd, int dtype, int rows, int H, float eps, cudaStream_t st);
// dz (= dx = dres), dgamma, dbeta and (dlin_bias != null) the gradient of the folded linear bias from (dy, z, mean, rstd, gamma)
void launch_add_ln_bwd(const void* dy, const void* z, const float* mean, const float* rstd, const void* gamma, void* dz, void* dgamma, void* dbeta,
                       void* dlin_bias, float* partial, int dtype, int rows, int H, cudaStream_t st);
void launch_bias_gelu_fwd(const void* x, const void* bias, void* y, int dtype, size_t rows, int I, cudaStream_t st);
void launch_bias_gelu_bwd(const void* dy, const void* x, const void* bias, void* dx, void* dbias, float* partial, int dtype, int rows, int I,
                          cudaStream_t st);
// (dq, dk, dv) [B, heads, S, d] with element strides {batch, head, seq} x 3 in `strides` -> out [B * S, 3 * heads * d] contiguous + dbias [3 * heads * d]
void launch_qkv_pack_bwd(const void* dq, const void* dk, const void* dv, const long long* strides, void* out, void* dbias, float* partial, int dtype,
                         int B, int S, int heads, int d, cudaStream_t st);

// ---- gemm_tc5.cu: hand-written tcgen05 GEMM + bias (+GELU) epilogue (TMA ring, TMEM accumulators; no CUTLASS); act: 0 = identity, 1 = GELU.
// `gemm_sm100_available` keeps its round-1 name: true when the driver exports cuTensorMapEncodeTiled
bool gemm_sm100_available();
size_t gemm_bias_act_workspace(int tokens, int in_f, int out_f, int act);
void launch_gemm_bias_act(const void* x, const void* w, const void* bias, void* y, int tokens, int in_f, int out_f, int act, void* workspace,
                          size_t workspace_bytes, cudaStream_t st);

void count_launch(int n = 1);
uint64_t launch_count();

} // namespace kf
