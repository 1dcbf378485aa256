"""Single-rank self-test of the collective / fused kernels.

Every kernel that normally talks to peers over NVLink accepts a one-rank mesh (peer 0 = itself): all-reduce (one-shot,
two-shot, split-phase), the fused S-SGD / SMA / PairAveraging steps, publish / pull (LDG and TMA), broadcast, all-gather, the
flag barrier and the completion wait. This module launches each of them once on the current GPU and checks the result
against plain PyTorch fp32 math, so that a box with ONE GPU still exercises (and a profiler still sees, by name) the code
paths the multi-GPU jobs run. `__graft_entry__.smoke()` and `tests/test_gpu_kernels.py` call it.

Counterpart in the reference: the np=1 `Peer` tests (tests/cpp/unit/test_kungfu.cpp) and the NCCL `simple_test`
(tests/cpp/integration/fake_nccl_trainer.cpp:34-63).
"""
from __future__ import annotations

from typing import Dict, List

import torch


def run_single_rank_kernels(plane=None, verbose: bool = False) -> Dict[str, object]:
    """Returns {"launched": [kernel family names], "fails": [...]}; raises nothing by itself."""
    from .runtime import get_runtime
    from .utils import dtypes

    rt = get_runtime()
    plane = plane or rt.gpu
    ext = plane.ext
    dev = plane.device
    st = torch.cuda.current_stream().cuda_stream
    n = 1 << 16  # bf16 elements: 128 KiB regions
    nb = n * 2
    arena = plane.new_arena(16 * nb + (4 << 20), "selftest")
    arena.ensure()
    mesh = arena.mesh
    launched: List[str] = []
    fails: List[str] = []
    BF16, F32 = dtypes.BF16, dtypes.F32
    THREADS, BLOCKS = 256, 4

    def region():
        off = arena.alloc(nb, align=4096)
        return off, arena.view(off, n, torch.bfloat16)

    def expect(name: str, ok: bool) -> None:
        launched.append(name)
        if not ok:
            fails.append(name)
        if verbose:
            print(f"[selftest] {name}: {'ok' if ok else 'FAIL'}")

    g = torch.Generator(device=dev).manual_seed(7)
    x = torch.randn(n, device=dev, generator=g).to(torch.bfloat16)

    # ---- all-reduce: one-shot (out of place), two-shot (in place), split-phase two-shot to another region --------------
    a_off, a = region()
    b_off, b = region()
    a.copy_(x)
    out = torch.zeros_like(x)
    ext.allreduce(mesh, 0, ext.ALGO_ONESHOT, a_off, out.data_ptr(), nb, BF16, dtypes.SUM, 1.0, 0, 0, BLOCKS, THREADS, st)
    expect("allreduce_oneshot", torch.equal(out, x))
    ll_max = 64 << 10
    ll_off = arena.alloc(ext.allreduce_ll_slot_bytes(ll_max), align=4096)
    for it in range(3):  # three operations: both slot parities and the counter hand-over
        src = (x[:ll_max // 2].float() + it).to(torch.bfloat16)
        out2 = torch.zeros_like(src)
        ext.allreduce_ll(mesh, 0, src.data_ptr(), out2.data_ptr(), ll_max, ll_off, ll_max, BF16, dtypes.SUM, 0.5, ext.FLAG_SCALE, st)
        good = torch.equal(out2, (src.float() * 0.5).to(torch.bfloat16))
        if it == 2 or not good:
            expect("allreduce_ll(flag-carrying pushes)", good)
            break
    acc = torch.zeros(2, device=dev)
    ext.allreduce(mesh, 0, ext.ALGO_TWOSHOT, a_off, 0, nb, BF16, dtypes.SUM, 0.5, acc.data_ptr(), ext.FLAG_SCALE | ext.FLAG_SUMSQ, BLOCKS,
                  THREADS, st)
    half = (x.float() * 0.5).to(torch.bfloat16)
    expect("allreduce_twoshot(scale,sumsq)", torch.equal(a, half) and abs(acc[0].item() / (x.float() ** 2).sum().item() - 1) < 1e-3)
    a.copy_(x)
    b.zero_()
    ext.barrier(mesh, 0, st)
    ext.allreduce_to(mesh, 0, ext.ALGO_TWOSHOT, a_off, b_off, nb, BF16, dtypes.SUM, 0.5, 0, ext.FLAG_SCALE | ext.FLAG_GATED, BLOCKS, THREADS, st)
    ext.wait_done(mesh, 0, 0, 1 * BLOCKS, st)
    expect("gate+allreduce_twoshot(split-phase,out-of-place)+wait_done", torch.equal(b, half) and torch.equal(a, x))

    # ---- broadcast / all-gather / barrier -----------------------------------------------------------------------------
    ext.broadcast(mesh, 0, ext.ALGO_TWOSHOT, a_off, nb, 0, BLOCKS, THREADS, st)
    expect("broadcast", torch.equal(a, x))
    b.zero_()
    ext.allgather(mesh, 0, ext.ALGO_TWOSHOT, a_off, b_off, nb, BLOCKS, THREADS, st)
    expect("allgather", torch.equal(b, x))
    ext.barrier(mesh, 0, st)
    launched.append("barrier")

    # ---- fused S-SGD: RS(grads) -> SGD(momentum) on fp32 master -> AG(params), plain and split-phase ---------------------
    hp = torch.zeros(16)
    lr, mom = 0.1, 0.9
    hp[0], hp[1], hp[5] = lr, mom, 1.0
    w0 = torch.randn(n, device=dev, generator=g)
    g_off, gr = region()
    p_off, pr = region()
    for gated in (False, True):
        hyper = hp.to(dev)
        master = w0.clone()
        s1 = torch.zeros(n, device=dev)
        pr.copy_(w0.to(torch.bfloat16))
        ref_w, ref_m = w0.clone(), torch.zeros(n, device=dev)
        for step in range(2):
            grad = torch.randn(n, device=dev, generator=g).to(torch.bfloat16)
            gr.copy_(grad)
            ext.hyper_tick(hyper.data_ptr(), st)
            if gated:
                ext.barrier(mesh, 0, st)
            ext.ssgd_fused(mesh, 0, ext.ALGO_TWOSHOT, g_off, p_off, BF16, nb, master.data_ptr(), s1.data_ptr(), 0, hyper.data_ptr(),
                           ext.OPT_SGD, 0, ext.FLAG_GATED if gated else 0, BLOCKS, THREADS, st)
            if gated:
                ext.wait_done(mesh, 0, 0, BLOCKS, st)
            ref_m = grad.float() if step == 0 else mom * ref_m + grad.float()
            ref_w = ref_w - lr * ref_m
        ok = torch.allclose(master, ref_w, atol=1e-5, rtol=1e-5) and torch.equal(pr, master.to(torch.bfloat16))
        expect("ssgd_fused" + ("(split-phase)" if gated else ""), ok)

    # ---- SMA: fused kernel, and out-of-place average + local blend kernel -----------------------------------------------
    hp2 = hp.clone()
    hp2[13] = 0.1  # alpha
    x_off, xr = region()
    for form in ("fused", "blend"):
        hyper = hp2.to(dev)
        s1 = torch.zeros(n, device=dev)
        pr.copy_(w0.to(torch.bfloat16))
        grad = torch.randn(n, device=dev, generator=g).to(torch.bfloat16)
        gr.copy_(grad)
        ext.hyper_tick(hyper.data_ptr(), st)
        if form == "fused":
            ext.sma_fused(mesh, 0, ext.ALGO_TWOSHOT, p_off, x_off, BF16, nb, gr.data_ptr(), BF16, s1.data_ptr(), 0, 0, BF16, hyper.data_ptr(),
                          ext.OPT_SGD, BLOCKS, THREADS, st)
        else:
            ext.barrier(mesh, 0, st)
            ext.allreduce_to(mesh, 0, ext.ALGO_TWOSHOT, p_off, x_off, nb, BF16, dtypes.SUM, 1.0, 0, ext.FLAG_SCALE | ext.FLAG_GATED, BLOCKS,
                             THREADS, st)
            ext.wait_done(mesh, 0, 0, BLOCKS, st)
            ext.sma_blend(pr.data_ptr(), xr.data_ptr(), BF16, nb, gr.data_ptr(), BF16, s1.data_ptr(), 0, hyper.data_ptr(), ext.OPT_SGD, 64, 256, st)
        p0 = w0.to(torch.bfloat16).float()
        ref = (p0 + 0.1 * (p0 - p0)) - lr * grad.float()  # one rank: the average is the model itself
        expect(f"sma_{form}", torch.allclose(pr.float(), ref, atol=2e-2, rtol=2e-2))

    # ---- model store: publish -> pull (LDG, TMA, retry form) -> pair averaging from the store and from a private copy ----
    pubs = [arena.alloc(nb, align=4096) for _ in range(3)]
    ver = arena.alloc(256, align=256)
    done = torch.zeros(4, dtype=torch.int32, device=dev)
    ok = torch.zeros(4, dtype=torch.int32, device=dev)
    src = x.clone()
    ext.publish(mesh, 1, pubs, ver, src.data_ptr(), nb, done.data_ptr(), BLOCKS, THREADS, st)
    dst = torch.zeros_like(x)
    ext.pull(mesh, 1, 0, pubs, ver, dst.data_ptr(), nb, ok.data_ptr(), 0, BLOCKS, THREADS, st)
    expect("publish+pull", torch.equal(dst, x) and int(ok[0].item()) == 1)
    dst.zero_()
    tgt = torch.zeros(1, dtype=torch.int32, device=dev)
    ext.pull_tma(mesh, 1, 0, pubs, ver, dst.data_ptr(), nb, ok.data_ptr(), 0, 4, st, tgt.data_ptr())
    ext.pull_tma(mesh, 1, 0, pubs, ver, dst.data_ptr(), nb, ok.data_ptr(), 1, 4, st, tgt.data_ptr())  # retry form: a no-op after a clean copy
    expect("pull_tma(+retry)", torch.equal(dst, x) and int(ok[2].item()) == 0)
    for private in (False, True):
        hyper = hp.to(dev)
        s1 = torch.zeros(n, device=dev)
        params = w0.to(torch.bfloat16).clone()
        grad = torch.randn(n, device=dev, generator=g).to(torch.bfloat16)
        ext.hyper_tick(hyper.data_ptr(), st)
        ext.pair_avg(mesh, 1, 0, pubs, ver, params.data_ptr(), BF16, nb, grad.data_ptr(), BF16, s1.data_ptr(), 0, hyper.data_ptr(), ext.OPT_SGD, 1,
                     done.data_ptr(), dst.data_ptr() if private else 0, ok.data_ptr() if private else 0, BLOCKS, THREADS, st, tgt.data_ptr())
        other = x if private or True else x
        ref = (0.5 * (w0.to(torch.bfloat16).float() + other.float())).to(torch.bfloat16).float() - lr * grad.float()
        expect("pair_avg" + ("(private copy)" if private else "(in-kernel read)"), torch.allclose(params.float(), ref, atol=2e-2, rtol=2e-2))
        # the next iteration's `other` is what this one published; keep the check simple by re-publishing x
        ext.publish(mesh, 1, pubs, ver, src.data_ptr(), nb, done.data_ptr(), BLOCKS, THREADS, st)
        ext.pull_tma(mesh, 1, 0, pubs, ver, dst.data_ptr(), nb, ok.data_ptr(), 0, 4, st, tgt.data_ptr())
    torch.cuda.synchronize()
    try:
        arena.check()
    except RuntimeError as e:
        fails.append(f"device status: {e}")
    return {"launched": launched, "fails": fails}
