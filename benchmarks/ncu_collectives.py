#!/usr/bin/env python
"""Workload for Nsight Compute captures of the collective / fused kernels in a LIVE multi-rank job.

Only ONE rank runs under ncu (scripts/gpu/ncu_collectives.sh starts the ranks by hand instead of torchrun); the others run
unprofiled. This works with the split-phase kernels because the profiled BODY kernels never wait for a peer: ncu may replay
them as often as it likes (each replay re-reads peer memory over NVLink and re-signals completion — harmless extra counts), and
the one-warp gate / wait kernels that do wait are filtered out of the capture (-k regex). Kernels launched per round, ResNet-50
sized (25.6 M bf16 elements):
    ssgd_fused (NVLS or two-shot)   allreduce (out-of-place, 1/np folded in = the SMA average)   pull_tma   pair_avg (private copy)
With WORLD_SIZE=1 the same script profiles the one-rank forms (incl. the LL one-shot, which polls and must not be replayed live).
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from kungfu_b200.ops import collective as coll  # noqa: E402
from kungfu_b200.runtime import get_runtime  # noqa: E402
from kungfu_b200.utils import dtypes  # noqa: E402


def main():
    rt = get_runtime()
    N, me = rt.size, rt.rank
    torch.cuda.set_device(rt.cuda_index())
    dev = torch.device("cuda", rt.cuda_index())
    plane = rt.gpu
    ext = plane.ext
    st = torch.cuda.current_stream()
    n = 25_600_000 // 4096 * 4096
    S = n * 2
    arena = plane.new_arena(7 * S + (8 << 20), "ncu").ensure()
    mesh = arena.mesh
    g_off, p_off, x_off = (arena.alloc(S, align=4096) for _ in range(3))
    pubs = [arena.alloc(S, align=4096) for _ in range(3)]
    ver = arena.alloc(256, align=256)
    arena.view(g_off, n, torch.bfloat16).normal_(0, 1e-3)
    params = arena.view(p_off, n, torch.bfloat16)
    params.normal_(0, 2e-2)
    hp = torch.zeros(16)
    hp[0], hp[1], hp[5], hp[7], hp[8], hp[9], hp[14] = 1e-4, 0.9, 1.0 / N, 0.9, 0.999, 1e-8, 1.0
    hyper = hp.to(dev)
    ext.hyper_tick(hyper.data_ptr(), st.cuda_stream)
    BF16 = dtypes.BF16
    algo = ext.ALGO_NVLS if (arena.has_multicast and N > 1) else ext.ALGO_TWOSHOT
    per_rank = ((n // 8 + N - 1) // N) * 8
    master = torch.zeros(per_rank + 4096, device=dev)
    mom = torch.zeros(per_rank + 4096, device=dev)
    s1 = torch.zeros(n, device=dev)
    s2 = torch.zeros(n, device=dev)
    mfull = params.float().clone()
    done = torch.zeros(4, dtype=torch.int32, device=dev)
    ok = torch.zeros(4, dtype=torch.int32, device=dev)
    pulled = torch.empty(S, dtype=torch.uint8, device=dev)
    tgt = torch.full((1,), (me + 1) % N, dtype=torch.int32, device=dev)
    ext.publish(mesh, 1, pubs, ver, params.data_ptr(), S, done.data_ptr(), 64, 512, st.cuda_stream)
    torch.cuda.synchronize()
    coll.barrier()
    blocks = int(os.environ.get("KUNGFU_COMM_BLOCKS", "32"))
    bucket = 8 << 20
    for rnd in range(int(os.environ.get("KF_NCU_ROUNDS", "2"))):
        nb = bucket
        b = coll.pick_blocks(nb, N, True, blocks)
        ext.barrier(mesh, 0, st.cuda_stream)
        ext.ssgd_fused(mesh, 0, algo, g_off, p_off, BF16, nb, master.data_ptr(), mom.data_ptr(), 0, hyper.data_ptr(), ext.OPT_SGD, 0, ext.FLAG_GATED, b,
                       512, st.cuda_stream)
        ext.wait_done(mesh, 0, 0, N * b, st.cuda_stream)
        b2 = coll.pick_blocks(S, N, True, blocks)
        ext.barrier(mesh, 0, st.cuda_stream)
        ext.allreduce_to(mesh, 0, algo, p_off, x_off, S, BF16, dtypes.SUM, 1.0 / N, 0, ext.FLAG_SCALE | ext.FLAG_GATED, b2, 512, st.cuda_stream)
        ext.wait_done(mesh, 0, 0, N * b2, st.cuda_stream)
        ext.pull_tma(mesh, 1, 0, pubs, ver, pulled.data_ptr(), S, ok.data_ptr(), 0, 24, st.cuda_stream, tgt.data_ptr())
        ext.pair_avg(mesh, 1, 0, pubs, ver, params.data_ptr(), BF16, S, arena.view(g_off, n, torch.bfloat16).data_ptr(), BF16, s1.data_ptr(), s2.data_ptr(),
                     hyper.data_ptr(), ext.OPT_ADAM, 1, done.data_ptr(), pulled.data_ptr(), ok.data_ptr(), 148, 512, st.cuda_stream, tgt.data_ptr(),
                     mfull.data_ptr())
        ext.sma_blend(params.data_ptr(), arena.view(x_off, n, torch.bfloat16).data_ptr(), BF16, S, arena.view(g_off, n, torch.bfloat16).data_ptr(), BF16,
                      s1.data_ptr(), 0, hyper.data_ptr(), ext.OPT_SGD, 4 * plane.sm_count, 512, st.cuda_stream)
        if N == 1:
            ll_off = arena.alloc(ext.allreduce_ll_slot_bytes(64 << 10), align=4096, name="ll")
            src = params[: (64 << 10) // 2].clone()
            out = torch.empty_like(src)
            ext.allreduce_ll(mesh, 0, src.data_ptr(), out.data_ptr(), 64 << 10, ll_off, 64 << 10, BF16, dtypes.SUM, 1.0, 0, st.cuda_stream)
        torch.cuda.synchronize()
    coll.barrier()
    if me == 0:
        print("ncu workload done", flush=True)


if __name__ == "__main__":
    main()
