// kungfu_b200 — Python bindings of the CUDA kernel library + PeerMesh (`kungfu_b200._kfcuda`).
//
// Counterpart of the reference's pybind module `kungfu_torch_ops` (srcs/cpp/src/torch/
// module_cuda.cpp:28-40). Tensors are passed as raw device pointers and streams as integers
// (`tensor.data_ptr()`, `torch.cuda.current_stream().cuda_stream`), which keeps this extension
// independent of libtorch's ABI and lets it build in seconds.
#include <pybind11/functional.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include "../engine/grad_engine.h"
#include "../kernels/kernels.h"
#include "../mesh/peer_mesh.h"

namespace py = pybind11;
using namespace kf;

static inline cudaStream_t S(uintptr_t s) { return reinterpret_cast<cudaStream_t>(s); }
template <typename T> static inline T* P(uintptr_t p) { return reinterpret_cast<T*>(p); }
static PubStore make_pub(const std::vector<size_t>& pubs, size_t ver_off) {
    if (pubs.size() < 2 || pubs.size() > 3) throw std::runtime_error("kungfu_b200: a published-model store has 2 or 3 buffers");
    PubStore ps{};
    for (size_t i = 0; i < pubs.size(); ++i) ps.pub_off[i] = pubs[i];
    ps.ver_off = ver_off;
    ps.nbuf = (int)pubs.size();
    return ps;
}

PYBIND11_MODULE(_kfcuda, m) {
    m.doc() = "kungfu_b200 sm_100a kernels and peer mesh";
    m.attr("MAX_RANKS") = kMaxRanks;
    m.attr("MAX_BLOCKS") = kMaxBlocks;
    m.attr("NUM_CHANNELS") = kNumChannels;
    m.attr("ALGO_AUTO") = (int)kAlgoAuto;
    m.attr("ALGO_ONESHOT") = (int)kAlgoOneShot;
    m.attr("ALGO_TWOSHOT") = (int)kAlgoTwoShot;
    m.attr("ALGO_NVLS") = (int)kAlgoNvls;
    m.attr("FLAG_SCALE") = (int)kArScale;
    m.attr("FLAG_SUMSQ") = (int)kArSumSq;
    m.attr("FLAG_GATED") = (int)kArGated;
    m.attr("OPT_SGD") = (int)kOptSgd;
    m.attr("OPT_ADAM") = (int)kOptAdam;
    m.attr("HYPER_COUNT") = (int)kHpCount;

    py::class_<PeerMesh>(m, "PeerMesh")
        .def(py::init<int, size_t, const std::string&>(), py::arg("device"), py::arg("arena_bytes"), py::arg("mode") = "auto")
        .def("connect",
             [](PeerMesh& self, int rank, int nranks, py::object oob, bool want_multicast) {
                 OobAllGather fn = [oob](const std::string& mine) {
                     py::object res = oob(py::bytes(mine));
                     std::vector<std::string> out;
                     for (auto item : res) out.push_back(py::cast<std::string>(py::bytes(py::reinterpret_borrow<py::object>(item))));
                     return out;
                 };
                 self.connect(rank, nranks, fn, want_multicast);
             },
             py::arg("rank"), py::arg("nranks"), py::arg("oob"), py::arg("want_multicast") = true)
        .def("disconnect", &PeerMesh::disconnect)
        .def_property_readonly("rank", &PeerMesh::rank)
        .def_property_readonly("nranks", &PeerMesh::nranks)
        .def_property_readonly("device", &PeerMesh::device)
        .def_property_readonly("epoch", &PeerMesh::epoch)
        .def_property_readonly("arena_bytes", &PeerMesh::arena_bytes)
        .def_property_readonly("alloc_bytes", &PeerMesh::alloc_bytes)
        .def_property_readonly("local_ptr", &PeerMesh::local_ptr)
        .def_property_readonly("mc_ptr", &PeerMesh::mc_ptr)
        .def_property_readonly("has_multicast", &PeerMesh::has_multicast)
        .def_property_readonly("mode", &PeerMesh::mode)
        .def_property_readonly("multicast_note", &PeerMesh::multicast_note)
        .def("peer_ptr", &PeerMesh::peer_ptr)
        .def("status", &PeerMesh::status)
        .def("clear_status", &PeerMesh::clear_status)
        .def("abort", &PeerMesh::abort)
        .def("set_timeout_ms", &PeerMesh::set_timeout_ms)
        .def("reset_signals", [](PeerMesh& self, uintptr_t st) { self.reset_signals(S(st)); })
        .def("topology", [](const PeerMesh& self) {
            MeshTopology t = self.topology();
            py::dict d;
            d["device"] = t.device;
            d["device_count"] = t.device_count;
            d["can_access"] = t.can_access;
            d["perf_rank"] = t.perf_rank;
            d["native_atomic"] = t.native_atomic;
            d["multicast_supported"] = t.multicast_supported;
            d["sm_count"] = t.sm_count;
            d["total_mem"] = t.total_mem;
            d["name"] = t.name;
            return d;
        });

    py::class_<GradEngine>(m, "GradEngine")
        .def(py::init<PeerMesh*, int, bool, bool, int, int, int>(), py::arg("mesh"), py::arg("mode"), py::arg("average"), py::arg("gated"),
             py::arg("max_blocks"), py::arg("threads"), py::arg("algo"), py::keep_alive<1, 2>())
        .def("add_set", &GradEngine::add_set)
        .def("add_bucket", &GradEngine::add_bucket)
        .def("add_entry", &GradEngine::add_entry)
        .def("set_bucket_state", &GradEngine::set_bucket_state)
        .def("set_monitor", &GradEngine::set_monitor)
        .def("set_comm_stream", &GradEngine::set_comm_stream)
        .def("set_deferred", &GradEngine::set_deferred)
        .def("open_step", [](GradEngine& e, uintptr_t st) { e.open_step(S(st)); })
        .def("grad_ready", [](GradEngine& e, int set, int entry, uintptr_t grad, uintptr_t st) { return e.grad_ready(set, entry, P<const void>(grad), S(st)); })
        .def("flush", [](GradEngine& e, uintptr_t st, bool zero_missing) { e.flush(S(st), zero_missing); })
        .def("finish", [](GradEngine& e, uintptr_t st) { return e.finish(S(st)); })
        .def_property_readonly("launches", &GradEngine::launches)
        .def_property_readonly("nranks", &GradEngine::nranks);
    m.attr("MODE_SSGD") = (int)kModeSsgd;
    m.attr("MODE_GRADS") = (int)kModeGrads;
    m.attr("MODE_COPY") = (int)kModeCopy;
    m.def("multi_copy", [](const std::vector<uintptr_t>& src, const std::vector<uintptr_t>& dst, const std::vector<size_t>& nbytes, uintptr_t st) {
        if (src.size() != dst.size() || src.size() != nbytes.size()) throw std::runtime_error("kungfu_b200: multi_copy lists differ in length");
        std::vector<const void*> s(src.size());
        std::vector<void*> d(dst.size());
        for (size_t i = 0; i < src.size(); ++i) s[i] = P<const void>(src[i]), d[i] = P<void>(dst[i]);
        launch_multi_copy(s.data(), d.data(), nbytes.data(), (int)src.size(), S(st));
    });

    m.def("launch_count", &launch_count);
    m.def("nvls_supported", &nvls_supported);

    m.def("allreduce",
          [](const PeerMesh& mesh, int channel, int algo, size_t off, uintptr_t out, size_t nbytes, int dtype, int op,
             float scale, uintptr_t sumsq, int flags, int blocks, int threads, uintptr_t st) {
              launch_allreduce(algo, mesh.args(channel), off, P<void>(out), nbytes, dtype, op, scale, P<float>(sumsq), flags,
                               blocks, threads, S(st));
          });
    // out-of-place between two regions of the symmetric arena (sliced algorithms only)
    m.def("allreduce_to",
          [](const PeerMesh& mesh, int channel, int algo, size_t off, size_t out_off, size_t nbytes, int dtype, int op,
             float scale, uintptr_t sumsq, int flags, int blocks, int threads, uintptr_t st) {
              launch_allreduce(algo, mesh.args(channel), off, nullptr, nbytes, dtype, op, scale, P<float>(sumsq), flags, blocks,
                               threads, S(st), out_off);
          });
    m.def("allreduce_ll_slot_bytes", &allreduce_ll_slot_bytes);
    m.def("allreduce_ll",
          [](const PeerMesh& mesh, int channel, uintptr_t in, uintptr_t out, size_t nbytes, size_t slot_off, size_t max_payload, int dtype, int op,
             float scale, int flags, uintptr_t st) {
              launch_allreduce_ll(mesh.args(channel), P<const void>(in), P<void>(out), nbytes, slot_off, max_payload, dtype, op, scale, flags, S(st));
          });
    m.def("wait_done", [](const PeerMesh& mesh, int channel, int word, uint32_t add, uintptr_t st) {
        launch_wait_done(mesh.args(channel), word, add, S(st));
    });
    m.def("sma_blend",
          [](uintptr_t params, uintptr_t avg, int dtype, size_t nbytes, uintptr_t grads, int grad_dtype, uintptr_t s1, uintptr_t s2,
             uintptr_t hyper, int opt_kind, int blocks, int threads, uintptr_t st) {
              launch_sma_blend(P<void>(params), P<const void>(avg), dtype, nbytes, P<const void>(grads), grad_dtype, P<float>(s1),
                               P<float>(s2), P<const float>(hyper), opt_kind, blocks, threads, S(st));
          });
    m.def("ssgd_fused",
          [](const PeerMesh& mesh, int channel, int algo, size_t grad_off, size_t param_off, int dtype, size_t nbytes,
             uintptr_t master, uintptr_t s1, uintptr_t s2, uintptr_t hyper, int opt_kind, uintptr_t sumsq, int flags,
             int blocks, int threads, uintptr_t st) {
              launch_ssgd_fused(algo, mesh.args(channel), grad_off, param_off, dtype, nbytes, P<float>(master), P<float>(s1),
                                P<float>(s2), P<const float>(hyper), opt_kind, P<float>(sumsq), flags, blocks, threads, S(st));
          });
    m.def("sma_fused",
          [](const PeerMesh& mesh, int channel, int algo, size_t param_off, size_t avg_off, int dtype, size_t nbytes,
             uintptr_t grads, int grad_dtype, uintptr_t s1, uintptr_t s2, uintptr_t model_out, int model_dtype,
             uintptr_t hyper, int opt_kind, int blocks, int threads, uintptr_t st) {
              launch_sma_fused(algo, mesh.args(channel), param_off, avg_off, dtype, nbytes, P<const void>(grads), grad_dtype,
                               P<float>(s1), P<float>(s2), P<void>(model_out), model_dtype, P<const float>(hyper), opt_kind,
                               blocks, threads, S(st));
          });
    m.def("pair_avg",
          [](const PeerMesh& mesh, int channel, int target, const std::vector<size_t>& pubs, size_t ver_off, uintptr_t params,
             int dtype, size_t nbytes, uintptr_t grads, int grad_dtype, uintptr_t s1, uintptr_t s2, uintptr_t hyper,
             int opt_kind, int do_step, uintptr_t done_counter, uintptr_t pulled, uintptr_t pulled_ok, int blocks, int threads,
             uintptr_t st, uintptr_t target_dev, uintptr_t master) {
              launch_pair_avg(mesh.args(channel), target, make_pub(pubs, ver_off), P<void>(params), dtype, nbytes,
                              P<const void>(grads), grad_dtype, P<float>(s1), P<float>(s2), P<const float>(hyper), opt_kind,
                              do_step, P<uint32_t>(done_counter), P<const void>(pulled), P<const uint32_t>(pulled_ok), blocks,
                              threads, S(st), P<const int>(target_dev), P<float>(master));
          },
          py::arg("mesh"), py::arg("channel"), py::arg("target"), py::arg("pubs"), py::arg("ver_off"), py::arg("params"), py::arg("dtype"),
          py::arg("nbytes"), py::arg("grads"), py::arg("grad_dtype"), py::arg("s1"), py::arg("s2"), py::arg("hyper"), py::arg("opt_kind"),
          py::arg("do_step"), py::arg("done_counter"), py::arg("pulled"), py::arg("pulled_ok"), py::arg("blocks"), py::arg("threads"),
          py::arg("stream"), py::arg("target_dev") = (uintptr_t)0, py::arg("master") = (uintptr_t)0);
    m.def("publish",
          [](const PeerMesh& mesh, int channel, const std::vector<size_t>& pubs, size_t ver_off, uintptr_t src, size_t nbytes,
             uintptr_t done_counter, int blocks, int threads, uintptr_t st) {
              launch_publish(mesh.args(channel), make_pub(pubs, ver_off), P<const void>(src), nbytes, P<uint32_t>(done_counter),
                             blocks, threads, S(st));
          });
    m.def("pull",
          [](const PeerMesh& mesh, int channel, int target, const std::vector<size_t>& pubs, size_t ver_off, uintptr_t dst,
             size_t nbytes, uintptr_t ok, int retry, int blocks, int threads, uintptr_t st, uintptr_t target_dev) {
              launch_pull(mesh.args(channel), target, make_pub(pubs, ver_off), P<void>(dst), nbytes, P<uint32_t>(ok), retry, blocks,
                          threads, S(st), P<const int>(target_dev));
          },
          py::arg("mesh"), py::arg("channel"), py::arg("target"), py::arg("pubs"), py::arg("ver_off"), py::arg("dst"), py::arg("nbytes"),
          py::arg("ok"), py::arg("retry"), py::arg("blocks"), py::arg("threads"), py::arg("stream"), py::arg("target_dev") = (uintptr_t)0);
    m.def("pull_tma",
          [](const PeerMesh& mesh, int channel, int target, const std::vector<size_t>& pubs, size_t ver_off, uintptr_t dst,
             size_t nbytes, uintptr_t ok, int retry, int blocks, uintptr_t st, uintptr_t target_dev) {
              launch_pull_tma(mesh.args(channel), target, make_pub(pubs, ver_off), P<void>(dst), nbytes, P<uint32_t>(ok), retry,
                              blocks, S(st), P<const int>(target_dev));
          },
          py::arg("mesh"), py::arg("channel"), py::arg("target"), py::arg("pubs"), py::arg("ver_off"), py::arg("dst"), py::arg("nbytes"),
          py::arg("ok"), py::arg("retry"), py::arg("blocks"), py::arg("stream"), py::arg("target_dev") = (uintptr_t)0);
    m.def("broadcast",
          [](const PeerMesh& mesh, int channel, int algo, size_t off, size_t nbytes, int root, int blocks, int threads,
             uintptr_t st) { launch_broadcast(algo, mesh.args(channel), off, nbytes, root, blocks, threads, S(st)); });
    m.def("allgather",
          [](const PeerMesh& mesh, int channel, int algo, size_t in_off, size_t out_off, size_t nbytes_per_rank, int blocks,
             int threads, uintptr_t st) {
              launch_allgather(algo, mesh.args(channel), in_off, out_off, nbytes_per_rank, blocks, threads, S(st));
          });
    m.def("barrier", [](const PeerMesh& mesh, int channel, uintptr_t st) { launch_barrier(mesh.args(channel), S(st)); });

    m.def("hyper_tick", [](uintptr_t hyper, uintptr_t st) { launch_hyper_tick(P<float>(hyper), S(st)); });
    m.def("flat_opt",
          [](uintptr_t params, int dtype, uintptr_t grads, int grad_dtype, uintptr_t master, uintptr_t s1, uintptr_t s2,
             uintptr_t hyper, int opt_kind, size_t nelem, uintptr_t st) {
              launch_flat_opt(P<void>(params), dtype, P<const void>(grads), grad_dtype, P<float>(master), P<float>(s1),
                              P<float>(s2), P<const float>(hyper), opt_kind, nelem, S(st));
          });
    m.def("cast", [](uintptr_t src, int sdt, uintptr_t dst, int ddt, size_t nelem, float scale, uintptr_t st) {
        launch_cast(P<const void>(src), sdt, P<void>(dst), ddt, nelem, scale, S(st));
    });
    m.def("sumsq", [](uintptr_t src, int dt, size_t nelem, uintptr_t out, uintptr_t st) {
        launch_sumsq(P<const void>(src), dt, nelem, P<float>(out), S(st));
    });
    m.def("sumsq_diff", [](uintptr_t a, uintptr_t b, int dt, size_t nelem, uintptr_t out, uintptr_t st) {
        launch_sumsq_diff(P<const void>(a), P<const void>(b), dt, nelem, P<float>(out), S(st));
    });
    m.def("axpby", [](uintptr_t y, uintptr_t x, int dt, size_t nelem, float a, float b, uintptr_t st) {
        launch_axpby(P<void>(y), P<const void>(x), dt, nelem, a, b, S(st));
    });
    m.def("gns_finalize",
          [](uintptr_t sums, uintptr_t state, uintptr_t out, float b_small, float b_big, float alpha, float nranks,
             uintptr_t st) {
              launch_gns_finalize(P<const float>(sums), P<float>(state), P<float>(out), b_small, b_big, alpha, nranks, S(st));
          });
    m.def("u8_normalize", [](uintptr_t src, uintptr_t dst, int ddt, size_t npix, uintptr_t mean3, uintptr_t invstd3, uintptr_t st) {
        launch_u8_normalize(P<const uint8_t>(src), P<void>(dst), ddt, npix, P<const float>(mean3), P<const float>(invstd3), S(st));
    });
    m.def("maxpool3x3s2_fwd", [](uintptr_t x, uintptr_t y, uintptr_t idx, int dtype, int N, int H, int W, int C, uintptr_t st) {
        launch_maxpool3x3s2_fwd(P<const void>(x), P<void>(y), P<uint8_t>(idx), dtype, N, H, W, C, S(st));
    });
    m.def("maxpool3x3s2_bwd", [](uintptr_t dy, uintptr_t idx, uintptr_t dx, int dtype, int N, int H, int W, int C, uintptr_t st) {
        launch_maxpool3x3s2_bwd(P<const void>(dy), P<const uint8_t>(idx), P<void>(dx), dtype, N, H, W, C, S(st));
    });
    m.def("gemm_sm100_available", &gemm_sm100_available);
    m.def("gemm_bias_act_workspace", &gemm_bias_act_workspace);
    m.def("gemm_bias_act", [](uintptr_t x, uintptr_t w, uintptr_t bias, uintptr_t y, int tokens, int in_f, int out_f, int act, uintptr_t ws,
                              size_t ws_bytes, uintptr_t st) {
        launch_gemm_bias_act(P<const void>(x), P<const void>(w), P<const void>(bias), P<void>(y), tokens, in_f, out_f, act, P<void>(ws), ws_bytes, S(st));
    });
    m.def("add_ln_num_blocks", &add_ln_num_blocks);
    m.def("bias_gelu_num_blocks", &bias_gelu_num_blocks);
    m.def("add_ln_fwd", [](uintptr_t x, uintptr_t res, uintptr_t lin_bias, uintptr_t gamma, uintptr_t beta, uintptr_t z, uintptr_t y, uintptr_t mean,
                           uintptr_t rstd, int dtype, int rows, int H, float eps, uintptr_t st) {
        launch_add_ln_fwd(P<const void>(x), P<const void>(res), P<const void>(lin_bias), P<const void>(gamma), P<const void>(beta), P<void>(z),
                          P<void>(y), P<float>(mean), P<float>(rstd), dtype, rows, H, eps, S(st));
    });
    m.def("add_ln_bwd", [](uintptr_t dy, uintptr_t z, uintptr_t mean, uintptr_t rstd, uintptr_t gamma, uintptr_t dz, uintptr_t dgamma, uintptr_t dbeta,
                           uintptr_t dlin_bias, uintptr_t partial, int dtype, int rows, int H, uintptr_t st) {
        launch_add_ln_bwd(P<const void>(dy), P<const void>(z), P<const float>(mean), P<const float>(rstd), P<const void>(gamma), P<void>(dz),
                          P<void>(dgamma), P<void>(dbeta), P<void>(dlin_bias), P<float>(partial), dtype, rows, H, S(st));
    });
    m.def("qkv_pack_num_blocks", &qkv_pack_num_blocks);
    m.def("qkv_pack_bwd", [](uintptr_t dq, uintptr_t dk, uintptr_t dv, std::vector<long long> strides, uintptr_t out, uintptr_t dbias,
                             uintptr_t partial, int dtype, int B, int S_, int heads, int d, uintptr_t st) {
        if (strides.size() != 9) throw std::runtime_error("qkv_pack_bwd: 9 strides expected ({batch, head, seq} of dq, dk, dv)");
        launch_qkv_pack_bwd(P<const void>(dq), P<const void>(dk), P<const void>(dv), strides.data(), P<void>(out), P<void>(dbias), P<float>(partial),
                            dtype, B, S_, heads, d, S(st));
    });
    m.def("bias_gelu_fwd", [](uintptr_t x, uintptr_t bias, uintptr_t y, int dtype, size_t rows, int I, uintptr_t st) {
        launch_bias_gelu_fwd(P<const void>(x), P<const void>(bias), P<void>(y), dtype, rows, I, S(st));
    });
    m.def("bias_gelu_bwd", [](uintptr_t dy, uintptr_t x, uintptr_t bias, uintptr_t dx, uintptr_t dbias, uintptr_t partial, int dtype, int rows, int I,
                              uintptr_t st) {
        launch_bias_gelu_bwd(P<const void>(dy), P<const void>(x), P<const void>(bias), P<void>(dx), P<void>(dbias), P<float>(partial), dtype, rows,
                             I, S(st));
    });
    m.def("bn_num_blocks", &bn_num_blocks);
    m.def("bn_fwd", [](uintptr_t x, uintptr_t res, uintptr_t y, uintptr_t mask, int dtype, size_t rows, int C, uintptr_t gamma, uintptr_t beta,
                       uintptr_t rmean, uintptr_t rvar, int pdtype, float eps, float momentum, int relu, uintptr_t partial, uintptr_t mean,
                       uintptr_t invstd, uintptr_t scale, uintptr_t shift, uintptr_t st) {
        launch_bn_fwd(P<const void>(x), P<const void>(res), P<void>(y), P<uint8_t>(mask), dtype, rows, C, P<const void>(gamma), P<const void>(beta),
                      P<void>(rmean), P<void>(rvar), pdtype, eps, momentum, relu, P<float>(partial), P<float>(mean), P<float>(invstd), P<float>(scale),
                      P<float>(shift), S(st));
    });
    m.def("bn_apply", [](uintptr_t x, uintptr_t res, uintptr_t y, uintptr_t mask, int dtype, size_t rows, int C, uintptr_t scale, uintptr_t shift,
                         int relu, uintptr_t st) {
        launch_bn_apply(P<const void>(x), P<const void>(res), P<void>(y), P<uint8_t>(mask), dtype, rows, C, P<const float>(scale), P<const float>(shift),
                        relu, S(st));
    });
    m.def("bn_bwd", [](uintptr_t dy, uintptr_t x, uintptr_t mask, uintptr_t dx, uintptr_t dres, int dtype, size_t rows, int C, uintptr_t gamma,
                       uintptr_t dgamma, uintptr_t dbeta, int pdtype, uintptr_t mean, uintptr_t invstd, int relu, uintptr_t partial, uintptr_t coef,
                       uintptr_t st) {
        launch_bn_bwd(P<const void>(dy), P<const void>(x), P<const uint8_t>(mask), P<void>(dx), P<void>(dres), dtype, rows, C, P<const void>(gamma),
                      P<void>(dgamma), P<void>(dbeta), pdtype, P<const float>(mean), P<const float>(invstd), relu, P<float>(partial), P<float>(coef), S(st));
    });
    m.def("fill_u32", [](uintptr_t p, uint32_t v, size_t n, uintptr_t st) { launch_fill_u32(P<uint32_t>(p), v, n, S(st)); });
}
