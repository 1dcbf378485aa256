"""Numerics of the sm_100a kernels on ONE GPU against plain PyTorch fp32 references.

Reference test model: C++ unit tests of `std_transform_2` / dtype sizes (tests/cpp/unit/test_operations.cpp) and the
NCCL `simple_test` sweep (tests/cpp/integration/fake_nccl_trainer.cpp:34-63).
"""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def plane():
    from kungfu_b200.runtime import get_runtime

    rt = get_runtime()
    return rt.gpu


def _ptr(t):
    return t.data_ptr()


def _st():
    return torch.cuda.current_stream().cuda_stream


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16, torch.int32, torch.int64, torch.uint8, torch.float64])
@pytest.mark.parametrize("algo", ["oneshot", "twoshot"])
def test_allreduce_single_rank_identity(plane, dtype, algo):
    """np=1: all-reduce must be the identity for every dtype/algorithm (≙ test_kungfu.cpp AllReduce with np=1)."""
    from kungfu_b200.utils import dtypes

    ext = plane.ext
    arena, off, cap = plane.staging()
    n = 4096 + 8
    if dtype.is_floating_point:
        x = torch.randn(n, device="cuda").to(dtype)
    else:
        x = torch.randint(0, 100, (n,), device="cuda").to(dtype)
    nb = x.numel() * x.element_size()
    nb_pad = (nb + 15) // 16 * 16
    stage = arena.bytes_view(off, nb_pad)
    stage.zero_()
    stage[:nb].copy_(x.view(torch.uint8))
    out = torch.zeros(nb_pad, dtype=torch.uint8, device="cuda")
    code = {"oneshot": ext.ALGO_ONESHOT, "twoshot": ext.ALGO_TWOSHOT}[algo]
    ext.allreduce(arena.mesh, 0, code, off, _ptr(out), nb_pad, dtypes.dtype_code(dtype), 0, 1.0, 0, 0, 4, 256, _st())
    torch.cuda.synchronize()
    arena.check()
    got = (out if algo == "oneshot" else stage)[:nb].view(dtype)
    assert torch.equal(got, x)


def test_allreduce_scale_and_sumsq(plane):
    from kungfu_b200.utils import dtypes

    ext = plane.ext
    arena, off, cap = plane.staging()
    n = 1 << 16
    x = torch.randn(n, device="cuda", dtype=torch.bfloat16)
    stage = arena.view(off, n, torch.bfloat16)
    stage.copy_(x)
    acc = torch.zeros(2, device="cuda")
    ext.allreduce(arena.mesh, 0, ext.ALGO_TWOSHOT, off, 0, n * 2, dtypes.BF16, 0, 0.5, _ptr(acc), ext.FLAG_SCALE | ext.FLAG_SUMSQ,
                  8, 512, _st())
    torch.cuda.synchronize()
    ref = (x.float() * 0.5).to(torch.bfloat16)
    assert torch.equal(stage, ref)
    pre = (x.float() ** 2).sum().item()
    post = ((x.float() * 0.5) ** 2).sum().item()
    assert math.isclose(acc[0].item(), pre, rel_tol=1e-3)
    assert math.isclose(acc[1].item(), post, rel_tol=1e-3)


@pytest.mark.parametrize("pdtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("kind", ["sgd", "sgd_nesterov", "adam", "adamw"])
def test_flat_opt_matches_torch(plane, pdtype, kind):
    """The device update rule == torch.optim on an fp32 reference (3 steps)."""
    from kungfu_b200.utils import dtypes

    ext = plane.ext
    torch.manual_seed(0)
    n = 10007
    w0 = torch.randn(n, device="cuda")
    ref = torch.nn.Parameter(w0.clone())
    if kind == "sgd":
        opt = torch.optim.SGD([ref], lr=0.1, momentum=0.9, weight_decay=1e-2, dampening=0.1)
    elif kind == "sgd_nesterov":
        opt = torch.optim.SGD([ref], lr=0.1, momentum=0.9, nesterov=True)
    elif kind == "adam":
        opt = torch.optim.Adam([ref], lr=1e-2, weight_decay=1e-2)
    else:
        opt = torch.optim.AdamW([ref], lr=1e-2, weight_decay=1e-1)
    g = opt.param_groups[0]
    hp = torch.zeros(16)
    hp[0] = g["lr"]
    hp[3] = g["weight_decay"]
    hp[5] = 1.0
    if kind.startswith("sgd"):
        hp[1], hp[2], hp[4] = g["momentum"], g["dampening"], float(g["nesterov"])
        k = ext.OPT_SGD
    else:
        hp[7], hp[8], hp[9] = g["betas"][0], g["betas"][1], g["eps"]
        hp[14] = 1.0 if kind == "adamw" else 0.0
        k = ext.OPT_ADAM
    hyper = hp.cuda()
    params = w0.to(pdtype).clone()
    master = w0.clone() if pdtype != torch.float32 else None
    s1 = torch.zeros(n, device="cuda")
    s2 = torch.zeros(n, device="cuda")
    for step in range(3):
        grad = torch.randn(n, device="cuda")
        gd = grad.to(pdtype)
        ref.grad = gd.float().clone()  # same (rounded) gradient; clone: foreach-nesterov updates .grad in place
        opt.step()
        ext.hyper_tick(_ptr(hyper), _st())
        ext.flat_opt(_ptr(params), dtypes.dtype_code(pdtype), _ptr(gd), dtypes.dtype_code(pdtype), _ptr(master) if master is not None else 0,
                     _ptr(s1), _ptr(s2), _ptr(hyper), k, n, _st())
    torch.cuda.synchronize()
    got = master if master is not None else params
    tol = 2e-5
    assert torch.allclose(got, ref.data, atol=tol, rtol=tol), (got - ref.data).abs().max()
    if pdtype != torch.float32:
        assert torch.equal(params, master.to(pdtype))


def test_cast_sumsq_axpby(plane):
    from kungfu_b200.utils import dtypes

    ext = plane.ext
    n = 123457
    x = torch.randn(n, device="cuda")
    y = torch.empty(n, device="cuda", dtype=torch.bfloat16)
    ext.cast(_ptr(x), dtypes.F32, _ptr(y), dtypes.BF16, n, 2.0, _st())
    assert torch.equal(y, (x * 2).to(torch.bfloat16))
    acc = torch.zeros(1, device="cuda")
    ext.sumsq(_ptr(x), dtypes.F32, n, _ptr(acc), _st())
    assert math.isclose(acc.item(), (x.double() ** 2).sum().item(), rel_tol=1e-4)
    z = torch.randn(n, device="cuda")
    acc.zero_()
    ext.sumsq_diff(_ptr(x), _ptr(z), dtypes.F32, n, _ptr(acc), _st())
    assert math.isclose(acc.item(), ((x - z).double() ** 2).sum().item(), rel_tol=1e-4)
    z0 = z.clone()
    ext.axpby(_ptr(z), _ptr(x), dtypes.F32, n, 0.25, 0.75, _st())
    assert torch.allclose(z, 0.25 * x + 0.75 * z0, atol=1e-6)


def test_u8_normalize(plane):
    from kungfu_b200.utils import dtypes

    ext = plane.ext
    B, H, W = 3, 17, 19
    img = torch.randint(0, 256, (B, H, W, 3), dtype=torch.uint8, device="cuda")
    mean = torch.tensor([120.0, 110.0, 100.0], device="cuda")
    inv = torch.tensor([1 / 58.0, 1 / 57.0, 1 / 56.0], device="cuda")
    out = torch.empty((B, 3, H, W), dtype=torch.bfloat16, device="cuda").contiguous(memory_format=torch.channels_last)
    ext.u8_normalize(_ptr(img), _ptr(out), dtypes.BF16, B * H * W, _ptr(mean), _ptr(inv), _st())
    ref = ((img.float() - mean) * inv).permute(0, 3, 1, 2)
    assert torch.allclose(out.float(), ref, atol=2e-2, rtol=1e-2)


def test_gns_finalize(plane):
    """Device finalisation == the reference formulas (ops/monitor.py:6-17, cpu/collective.cpp:245-257)."""
    ext = plane.ext
    state = torch.zeros(4, device="cuda")
    out = torch.zeros(4, device="cuda")
    bs, bb, alpha, nr = 32.0, 128.0, 0.9, 4.0
    ge = se = None
    for gs_sum, gb in [(40.0, 6.0), (36.0, 5.0), (44.0, 7.5)]:
        sums = torch.tensor([gs_sum, gb], device="cuda")
        ext.gns_finalize(_ptr(sums), _ptr(state), _ptr(out), bs, bb, alpha, nr, _st())
        g_small = gs_sum / nr
        G = (bb * gb - bs * g_small) / (bb - bs)
        S = (g_small - gb) / (1 / bs - 1 / bb)
        ge = G if ge is None else alpha * ge + (1 - alpha) * G
        se = S if se is None else alpha * se + (1 - alpha) * S
        torch.cuda.synchronize()
        assert math.isclose(out[0].item(), se / ge, rel_tol=1e-4)
        assert math.isclose(out[1].item(), G, rel_tol=1e-4)
        assert math.isclose(out[2].item(), S, rel_tol=1e-4)


def _mlp(dtype):
    torch.manual_seed(1)
    return torch.nn.Sequential(torch.nn.Linear(64, 128), torch.nn.ReLU(), torch.nn.Linear(128, 10)).cuda().to(dtype)


@pytest.mark.parametrize("optname", ["sgd", "adam"])
def test_ssgd_optimizer_fp32_matches_torch(plane, optname):
    """np=1 fused path (flat arena + our optimizer kernel) == torch.optim, fp32, several steps."""
    import kungfu_b200.torch as kf

    m1, m2 = _mlp(torch.float32), _mlp(torch.float32)
    m2.load_state_dict(m1.state_dict())
    mk = (lambda p: torch.optim.SGD(p, lr=0.05, momentum=0.9, weight_decay=1e-3)) if optname == "sgd" else (
        lambda p: torch.optim.Adam(p, lr=1e-2))
    o1 = mk(m1.parameters())
    o2 = kf.optimizers.SynchronousSGDOptimizer(mk(m2.parameters()), named_parameters=m2.named_parameters())
    assert o2.engine is not None and o2.engine.mode == "ssgd"
    x = torch.randn(32, 64, device="cuda")
    y = torch.randint(0, 10, (32,), device="cuda")
    for _ in range(5):
        for m, o in ((m1, o1), (m2, o2)):
            o.zero_grad()
            torch.nn.functional.cross_entropy(m(x), y).backward()
            o.step()
    torch.cuda.synchronize()
    o2.engine.check()
    for a, b in zip(m1.parameters(), m2.parameters()):
        assert torch.allclose(a, b, atol=1e-5, rtol=1e-4), (a - b).abs().max()


def test_ssgd_optimizer_generic_path(plane):
    """An optimizer without a fused kernel (RMSprop) goes through the bucketed grads mode."""
    import kungfu_b200.torch as kf

    m1, m2 = _mlp(torch.float32), _mlp(torch.float32)
    m2.load_state_dict(m1.state_dict())
    o1 = torch.optim.RMSprop(m1.parameters(), lr=1e-3)
    o2 = kf.optimizers.SynchronousSGDOptimizer(torch.optim.RMSprop(m2.parameters(), lr=1e-3), named_parameters=m2.named_parameters())
    assert o2.engine.mode == "grads"
    x = torch.randn(16, 64, device="cuda")
    y = torch.randint(0, 10, (16,), device="cuda")
    for _ in range(3):
        for m, o in ((m1, o1), (m2, o2)):
            o.zero_grad()
            torch.nn.functional.cross_entropy(m(x), y).backward()
            o.step()
    for a, b in zip(m1.parameters(), m2.parameters()):
        assert torch.allclose(a, b, atol=1e-6)


def test_ssgd_bf16_master_weights(plane):
    """bf16 parameters + fp32 master shard: tracks an fp32 torch reference within bf16 resolution."""
    import kungfu_b200.torch as kf

    ref = _mlp(torch.float32)
    m = _mlp(torch.bfloat16)
    with torch.no_grad():
        for a, b in zip(ref.parameters(), m.parameters()):
            a.copy_(b.float())
    o1 = torch.optim.SGD(ref.parameters(), lr=0.05, momentum=0.9)
    o2 = kf.optimizers.SynchronousSGDOptimizer(torch.optim.SGD(m.parameters(), lr=0.05, momentum=0.9), named_parameters=m.named_parameters())
    x = torch.randn(32, 64, device="cuda")
    y = torch.randint(0, 10, (32,), device="cuda")
    l0 = None
    for i in range(10):
        o1.zero_grad()
        torch.nn.functional.cross_entropy(ref(x), y).backward()
        o1.step()
        o2.zero_grad()
        loss = torch.nn.functional.cross_entropy(m(x.bfloat16()).float(), y)
        loss.backward()
        o2.step()
        l0 = l0 or loss.item()
    assert loss.item() < l0
    for a, b in zip(ref.parameters(), m.parameters()):
        assert torch.allclose(a, b.float(), atol=3e-2, rtol=3e-2)


def test_image_trainer_cuda_graph(plane):
    """Whole-step CUDA graph capture (fwd + bwd + hooks + fused optimizer on the comm stream) replays correctly."""
    import kungfu_b200.torch as kf
    from kungfu_b200.models import resnet
    from kungfu_b200.trainer import ImageTrainer

    torch.manual_seed(0)
    model = resnet.to_b200(resnet.resnet18(num_classes=10), torch.bfloat16, "cuda")
    opt = kf.optimizers.SynchronousSGDOptimizer(torch.optim.SGD(model.parameters(), lr=0.02, momentum=0.9),
                                                named_parameters=model.named_parameters())
    tr = ImageTrainer(model, opt, 16, 64, torch.bfloat16, use_graph=True)
    assert tr.graph is not None, tr.graph_error
    assert tr.launches_per_step > 0
    imgs = torch.randint(0, 256, (16, 64, 64, 3), dtype=torch.uint8).pin_memory()
    labels = torch.randint(0, 10, (16,), dtype=torch.int64).pin_memory()
    losses = [tr.step(imgs, labels) for _ in range(12)]
    opt.engine.check()
    assert all(math.isfinite(l) for l in losses)
    assert losses[-1] < losses[0]


@pytest.mark.parametrize("C,HW,residual,relu", [(64, 56, False, True), (256, 14, True, True), (2048, 7, True, True), (512, 28, False, False), (128, 9, True, False)])
def test_fused_bn_act_matches_torch(plane, C, HW, residual, relu):
    """Fused BN(+add)+ReLU forward/backward == PyTorch batch_norm + add + relu in fp32 (bf16 activations)."""
    from kungfu_b200.ops.bn_act import bn_act, supported

    torch.manual_seed(C + HW)
    N = 8
    x = (torch.randn(N, C, HW, HW, device="cuda") * 2 + 0.5).bfloat16().contiguous(memory_format=torch.channels_last).requires_grad_()
    res = torch.randn(N, C, HW, HW, device="cuda").bfloat16().contiguous(memory_format=torch.channels_last).requires_grad_() if residual else None
    w = (torch.rand(C, device="cuda") + 0.5).bfloat16().requires_grad_()
    b = torch.randn(C, device="cuda").bfloat16().requires_grad_()
    rm, rv = torch.zeros(C, device="cuda").bfloat16(), torch.ones(C, device="cuda").bfloat16()
    assert supported(x, w)
    y = bn_act(x, w, b, rm, rv, res, relu, True, 0.1, 1e-5)
    dy = torch.randn_like(y)
    y.backward(dy)
    # fp32 reference
    xf, wf, bf = x.detach().float().requires_grad_(), w.detach().float().requires_grad_(), b.detach().float().requires_grad_()
    rf = res.detach().float().requires_grad_() if residual else None
    rmf, rvf = torch.zeros(C, device="cuda"), torch.ones(C, device="cuda")
    z = torch.nn.functional.batch_norm(xf, rmf, rvf, wf, bf, True, 0.1, 1e-5)
    if residual:
        z = z + rf
    yf = torch.relu(z) if relu else z
    yf.backward(dy.float())
    assert torch.allclose(y.float(), yf, atol=3e-2, rtol=3e-2), (y.float() - yf).abs().max()
    assert torch.allclose(x.grad.float(), xf.grad, atol=3e-2, rtol=5e-2), (x.grad.float() - xf.grad).abs().max()
    if residual:
        assert torch.allclose(res.grad.float(), rf.grad, atol=1e-2, rtol=1e-2)
    scale = max(1.0, float(wf.grad.abs().max()))
    assert torch.allclose(w.grad.float(), wf.grad, atol=2e-2 * scale, rtol=3e-2), (w.grad.float() - wf.grad).abs().max()
    assert torch.allclose(b.grad.float(), bf.grad, atol=2e-2 * max(1.0, float(bf.grad.abs().max())), rtol=3e-2)
    assert torch.allclose(rm.float(), rmf, atol=1e-2) and torch.allclose(rv.float(), rvf, atol=2e-2, rtol=2e-2)
    # eval mode uses the running statistics
    ye = bn_act(x.detach(), w.detach(), b.detach(), rm, rv, res.detach() if residual else None, relu, False, 0.1, 1e-5)
    ze = torch.nn.functional.batch_norm(x.detach().float(), rm.float(), rv.float(), w.detach().float(), b.detach().float(), False, 0.1, 1e-5)
    if residual:
        ze = ze + res.detach().float()
    assert torch.allclose(ye.float(), torch.relu(ze) if relu else ze, atol=5e-2, rtol=3e-2)


def test_resnet_fused_bn_trains(plane):
    """ResNet-18 with the fused BN kernels trains as well as with the PyTorch ops (same seed, loss curves agree)."""
    from kungfu_b200.models import resnet

    def run(fused):
        resnet.BNAct.fused = fused
        torch.manual_seed(0)
        m = resnet.to_b200(resnet.resnet18(num_classes=10), torch.bfloat16, "cuda")
        opt = torch.optim.SGD(m.parameters(), lr=0.05, momentum=0.9)
        g = torch.Generator(device="cuda").manual_seed(1)
        x = torch.randn(16, 3, 64, 64, device="cuda", generator=g).bfloat16().contiguous(memory_format=torch.channels_last)
        y = torch.randint(0, 10, (16,), device="cuda", generator=g)
        losses = []
        for _ in range(8):
            opt.zero_grad()
            loss = torch.nn.functional.cross_entropy(m(x).float(), y)
            loss.backward()
            opt.step()
            losses.append(loss.item())
        return losses

    try:
        a, b = run(True), run(False)
    finally:
        resnet.BNAct.fused = True
    assert a[-1] < a[0] and b[-1] < b[0]
    assert abs(a[0] - b[0]) < 0.05 and abs(a[-1] - b[-1]) < 0.35, (a, b)


@pytest.mark.parametrize("tokens,in_f,out_f", [(256, 1024, 4096), (4096, 1024, 4096), (136, 64, 128)])
@pytest.mark.parametrize("act", [None, "gelu"])
def test_tcgen05_linear_act_matches_torch(plane, tokens, in_f, out_f, act):
    """tcgen05 GEMM + fused bias/GELU epilogue == fp32 reference of the same op (forward and backward)."""
    from kungfu_b200.ops.linear_act import linear_act, supported

    torch.manual_seed(tokens + out_f)
    x = (torch.randn(tokens, in_f, device="cuda") * 0.5).bfloat16().requires_grad_()
    w = (torch.randn(out_f, in_f, device="cuda") * in_f ** -0.5).bfloat16().requires_grad_()
    b = torch.randn(out_f, device="cuda").bfloat16().requires_grad_()
    assert supported(x, w, b), "the tcgen05 GEMM must be built and usable on this GPU"
    c0 = plane.ext.launch_count()
    y = linear_act(x, w, b, act)
    assert plane.ext.launch_count() == c0 + 1
    dy = torch.randn_like(y)
    y.backward(dy)
    xf, wf, bf = x.detach().float().requires_grad_(), w.detach().float().requires_grad_(), b.detach().float().requires_grad_()
    z = torch.nn.functional.linear(xf, wf, bf)
    yf = torch.nn.functional.gelu(z) if act == "gelu" else z
    yf.backward(dy.float())
    assert torch.allclose(y.float(), yf, atol=3e-2, rtol=3e-2), (y.float() - yf).abs().max()
    assert torch.allclose(x.grad.float(), xf.grad, atol=5e-2, rtol=5e-2)
    assert torch.allclose(w.grad.float(), wf.grad, atol=2e-1 * (tokens / 256) ** 0.5, rtol=5e-2)
    assert torch.allclose(b.grad.float(), bf.grad, atol=2e-1 * (tokens / 256) ** 0.5, rtol=5e-2)


@pytest.mark.parametrize("N,C,H,W", [(4, 64, 112, 112), (2, 8, 7, 9), (3, 16, 10, 10)])
def test_maxpool3x3s2_matches_torch(plane, N, C, H, W):
    from kungfu_b200.ops.pool import max_pool3x3s2

    torch.manual_seed(H * W)
    x = torch.randn(N, C, H, W, device="cuda").bfloat16().contiguous(memory_format=torch.channels_last).requires_grad_()
    y = max_pool3x3s2(x)
    dy = torch.randn_like(y)
    y.backward(dy)
    xr = x.detach().clone().requires_grad_()
    yr = torch.nn.functional.max_pool2d(xr, 3, 2, 1)
    yr.backward(dy)
    assert y.shape == yr.shape and torch.equal(y, yr)
    assert torch.allclose(x.grad.float(), xr.grad.float(), atol=1e-2, rtol=1e-2), (x.grad.float() - xr.grad.float()).abs().max()


def test_collective_kernels_on_one_rank_mesh(plane):
    """all-reduce (one-shot / two-shot / split-phase), fused S-SGD / SMA / PairAveraging, publish / pull (LDG + TMA + retry),
    broadcast, all-gather, barrier, completion wait: each launched on a one-rank mesh and compared with fp32 PyTorch math."""
    from kungfu_b200.selftest import run_single_rank_kernels

    res = run_single_rank_kernels(plane, verbose=True)
    assert not res["fails"], res["fails"]
    assert len(res["launched"]) >= 14


@pytest.mark.parametrize("H", [64, 768, 1024])
@pytest.mark.parametrize("with_res", [True, False])
def test_add_layernorm_matches_fp32(H, with_res):
    """Fused residual + LayerNorm forward/backward vs an fp32 PyTorch reference of the same op."""
    from kungfu_b200.ops.transformer import add_layernorm

    torch.manual_seed(1)
    rows = 515
    x = torch.randn(rows, H, device="cuda").bfloat16().requires_grad_()
    r = torch.randn(rows, H, device="cuda").bfloat16().requires_grad_() if with_res else None
    w = (1 + 0.1 * torch.randn(H, device="cuda")).bfloat16().requires_grad_()
    b = (0.1 * torch.randn(H, device="cuda")).bfloat16().requires_grad_()
    dy = torch.randn(rows, H, device="cuda").bfloat16()
    y = add_layernorm(x, r, w, b, 1e-5)
    y.backward(dy)
    xf, wf, bf = x.detach().float().requires_grad_(), w.detach().float().requires_grad_(), b.detach().float().requires_grad_()
    rf = r.detach().float().requires_grad_() if with_res else None
    zf = ((xf + rf) if with_res else xf).bfloat16().float()  # the kernel rounds the saved sum to storage precision (casts are differentiable)
    yf = torch.nn.functional.layer_norm(zf, (H,), wf, bf, 1e-5)
    yf.backward(dy.float())
    assert torch.allclose(y.float(), yf, atol=3e-2, rtol=3e-2)
    assert torch.allclose(x.grad.float(), xf.grad, atol=3e-2, rtol=3e-2)
    if with_res:
        assert torch.equal(x.grad, r.grad)
    # parameter gradients are sums over 515 rows: compare relative to their scale
    for got, ref in ((w.grad, wf.grad), (b.grad, bf.grad)):
        assert (got.float() - ref).abs().max() <= 2e-2 * ref.abs().max() + 1e-2


@pytest.mark.parametrize("H", [256, 1024])
def test_add_layernorm_folds_the_linear_bias(H):
    """`lin_bias` (the bias of the GEMM that produced x) is added inside the kernel; its gradient is the column sum of dz."""
    from kungfu_b200.ops.transformer import add_layernorm

    torch.manual_seed(5)
    rows = 777
    x = torch.randn(rows, H, device="cuda").bfloat16().requires_grad_()
    r = torch.randn(rows, H, device="cuda").bfloat16().requires_grad_()
    lb = (0.3 * torch.randn(H, device="cuda")).bfloat16().requires_grad_()
    w = (1 + 0.1 * torch.randn(H, device="cuda")).bfloat16().requires_grad_()
    b = (0.1 * torch.randn(H, device="cuda")).bfloat16().requires_grad_()
    dy = torch.randn(rows, H, device="cuda").bfloat16()
    y = add_layernorm(x, r, w, b, 1e-5, lin_bias=lb)
    y.backward(dy)
    xf, rf, lf, wf, bf = (t.detach().float().requires_grad_() for t in (x, r, lb, w, b))
    zf = (xf + lf + rf).bfloat16().float()
    yf = torch.nn.functional.layer_norm(zf, (H,), wf, bf, 1e-5)
    yf.backward(dy.float())
    assert torch.allclose(y.float(), yf, atol=3e-2, rtol=3e-2)
    assert torch.allclose(x.grad.float(), xf.grad, atol=3e-2, rtol=3e-2) and torch.equal(x.grad, r.grad)
    for got, ref in ((w.grad, wf.grad), (b.grad, bf.grad), (lb.grad, lf.grad)):
        assert (got.float() - ref).abs().max() <= 2e-2 * ref.abs().max() + 1e-2
    # the bias gradient is exactly the column sum of the dz the kernel stored (fp32 accumulation, one rounding)
    assert torch.allclose(lb.grad.float(), x.grad.float().sum(0), atol=2e-2 * float(x.grad.float().sum(0).abs().max()) + 1e-3, rtol=0)


@pytest.mark.parametrize("heads,d", [(4, 64), (16, 64)])
def test_qkv_projection_matches_autograd(heads, d):
    """Fused QKV projection: forward views equal the plain expression; backward (one pack kernel for cat + layout copy + bias column sum, then
    the two GEMMs) equals autograd's through the same attention."""
    from kungfu_b200.ops.transformer import qkv_projection

    torch.manual_seed(6)
    B, S, H = 3, 40, heads * d
    x = torch.randn(B, S, H, device="cuda").bfloat16().requires_grad_()
    w = (torch.randn(3 * H, H, device="cuda") / H ** 0.5).bfloat16().requires_grad_()
    b = (0.1 * torch.randn(3 * H, device="cuda")).bfloat16().requires_grad_()
    q, k, v = qkv_projection(x, w, b, heads)
    a = torch.nn.functional.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(B, S, H)
    da = torch.randn_like(a)
    a.backward(da)
    x2, w2, b2 = (t.detach().clone().requires_grad_() for t in (x, w, b))
    q2, k2, v2 = torch.nn.functional.linear(x2, w2, b2).view(B, S, 3, heads, d).permute(2, 0, 3, 1, 4)
    a2 = torch.nn.functional.scaled_dot_product_attention(q2, k2, v2).transpose(1, 2).reshape(B, S, H)
    a2.backward(da)
    assert torch.equal(q, q2) and torch.equal(k, k2) and torch.equal(v, v2) and torch.equal(a, a2)
    for got, ref in ((x.grad, x2.grad), (w.grad, w2.grad), (b.grad, b2.grad)):
        assert (got.float() - ref.float()).abs().max() <= 2e-2 * ref.float().abs().max() + 1e-3
    # non-contiguous / odd gradients take the repack fallback and must give the same answer
    from kungfu_b200.ops.transformer import _QKVProjection

    q, k, v = _QKVProjection.apply(x.detach().requires_grad_(), w.detach(), b.detach(), heads)
    g = torch.randn(B, heads, S, d + 8, device="cuda").bfloat16()[..., 4:d + 4]  # misaligned base pointer
    (q * g).sum().backward()


def test_gelu_rational_erf_accuracy():
    """The A&S 7.1.26 form used by bias_gelu stays within bf16 resolution of the exact erf GELU over the whole useful range."""
    from kungfu_b200.ops.transformer import bias_gelu

    x = torch.linspace(-9, 9, 1 << 16, device="cuda").half().view(-1, 1 << 10).requires_grad_()
    zero = torch.zeros(1 << 10, device="cuda").half()
    y = bias_gelu(x, zero)
    y.backward(torch.ones_like(y))
    xf = x.detach().double().requires_grad_()
    yf = torch.nn.functional.gelu(xf)
    yf.backward(torch.ones_like(yf))
    assert (y.double() - yf).abs().max() < 2e-3 and (x.grad.double() - xf.grad).abs().max() < 2e-3
    assert torch.isfinite(y).all() and torch.isfinite(x.grad).all()


@pytest.mark.parametrize("I", [128, 3072, 4096])
def test_bias_gelu_matches_fp32(I):
    from kungfu_b200.ops.transformer import bias_gelu

    torch.manual_seed(2)
    rows = 300
    x = torch.randn(rows, I, device="cuda").bfloat16().requires_grad_()
    b = (0.5 * torch.randn(I, device="cuda")).bfloat16().requires_grad_()
    dy = torch.randn(rows, I, device="cuda").bfloat16()
    y = bias_gelu(x, b)
    y.backward(dy)
    xf, bf = x.detach().float().requires_grad_(), b.detach().float().requires_grad_()
    yf = torch.nn.functional.gelu(xf + bf)
    yf.backward(dy.float())
    assert torch.allclose(y.float(), yf, atol=2e-2, rtol=2e-2)
    assert torch.allclose(x.grad.float(), xf.grad, atol=2e-2, rtol=2e-2)
    assert (b.grad.float() - bf.grad).abs().max() <= 2e-2 * bf.grad.abs().max() + 1e-2


def test_bert_fused_ops_match_unfused():
    """BERT (tiny config, bf16) with the fused residual+LayerNorm / bias+GELU kernels == the same weights through the PyTorch ops."""
    from kungfu_b200.models import bert

    torch.manual_seed(3)
    cfg = bert.BertConfig.tiny()
    m = bert.Bert(cfg).cuda().bfloat16()
    ids = torch.randint(0, cfg.vocab_size, (4, 32), device="cuda")
    y = torch.randint(0, 2, (4,), device="cuda")
    out = {}
    for fused in (True, False):
        cfg.fused_ops = fused
        for l in m.layers:
            l.fused_ops = fused
        m.zero_grad()
        logits = m(ids)
        torch.nn.functional.cross_entropy(logits.float(), y).backward()
        out[fused] = (logits.float().clone(), m.layers[0].ff1.bias.grad.float().clone(), m.layers[1].ln2.weight.grad.float().clone())
    for a, b in zip(out[True], out[False]):
        assert (a - b).abs().max() <= 5e-2 * b.abs().max() + 2e-2
