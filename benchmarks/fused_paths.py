#!/usr/bin/env python
"""Every fused optimizer path timed in isolation at model scale, with its roofline fraction (BASELINE.json: "each fused path
reported as achieved fraction of its roofline — the slower of its compute at peak and its bytes over NVLink at link bandwidth").

Paths (bf16 parameters / gradients, fp32 master + state where the optimizer keeps them):
  ssgd      gate -> ssgd_fused (reduce-scatter grads + SGD on the owned shard + all-gather params) -> wait, 8 MiB buckets
  sma_avg   gate -> out-of-place all-reduce params -> scratch (1/np folded in) -> wait          [runs under forward+backward]
  sma_blend local blend + momentum-SGD step over the arena                                       [HBM-bound]
  pull      TMA pull of a peer's published model into a private buffer                           [runs under forward+backward]
  pair_step local average + AdamW + publish                                                      [HBM-bound]
NVLink bytes per GPU and direction: NVLS all-reduce = S (1 + 1/N) (switch reduces and multicasts; checked against the link counters),
two-shot = 2 (N-1)/N S, pull = S. Every row also carries the bytes the hardware link counters saw per launch (`nvidia-smi nvlink -gt d`).
Rooflines: NVLink 770 GB/s measured peer copy (900 nominal); HBM copy from MEASURED_PEAKS.json (fallback 6650 GB/s).

  torchrun --nproc-per-node N benchmarks/fused_paths.py [--models resnet50,bert-large] [--out profiles/fused_paths_N.json]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from kungfu_b200.benchmarks import model_sizes  # noqa: E402
from kungfu_b200.ops import collective as coll  # noqa: E402
from kungfu_b200.runtime import get_runtime  # noqa: E402
from kungfu_b200.utils import dtypes  # noqa: E402
from kungfu_b200.utils.timing import ClockSampler, max_over_ranks, nvlink_counters  # noqa: E402


def hbm_peak():
    try:
        d = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(d.get("hbm_gbs") or d.get("hbm_copy_gbs")), "measured"
    except Exception:
        return 6650.0, "fallback"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--models", default="resnet50,bert-large")
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--blocks", type=int, default=32)
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    rt = get_runtime()
    N, me = rt.size, rt.rank
    torch.cuda.set_device(rt.cuda_index())
    dev = torch.device("cuda", rt.cuda_index())
    plane = rt.gpu
    ext = plane.ext
    st = torch.cuda.current_stream()
    hbm, hbm_src = hbm_peak()
    BF16 = dtypes.BF16
    rows = []
    sampler = ClockSampler(rt.cuda_index()).start() if me == 0 else None

    last_links = {}

    def timeit(fn, iters):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        rt.control.barrier()
        c0 = nvlink_counters(rt.cuda_index()) if me == 0 else None
        rt.control.barrier()  # rank 0 has read its counters (a subprocess: ~0.1 s) — nobody starts the timed loop before that
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(iters):
            fn()
        t1.record()
        torch.cuda.synchronize()
        rt.control.barrier()  # every rank's traffic of this loop has crossed the links before they are read again
        c1 = nvlink_counters(rt.cuda_index()) if me == 0 else None
        last_links.clear()
        if c0 and c1:
            last_links.update({"nvlink_tx_bytes_per_iter_counted": round((c1["tx"] - c0["tx"]) / iters), "nvlink_rx_bytes_per_iter_counted": round((c1["rx"] - c0["rx"]) / iters)})
        return max_over_ranks(t0.elapsed_time(t1) / iters) * 1e3  # us

    for model in a.models.split(","):
        n = sum(model_sizes.sizes(model))
        n = (n + 4095) // 4096 * 4096
        S = n * 2
        arena = plane.new_arena(6 * S + (8 << 20), "fp-" + model).ensure()
        mesh = arena.mesh
        g_off, p_off, x_off = (arena.alloc(S, align=4096) for _ in range(3))
        pubs = [arena.alloc(S, align=4096) for _ in range(3)]
        ver = arena.alloc(256, align=256)
        grads, params = arena.view(g_off, n, torch.bfloat16), arena.view(p_off, n, torch.bfloat16)
        grads.normal_(0, 1e-3)
        params.normal_(0, 2e-2)
        hp = torch.zeros(16)
        hp[0], hp[1], hp[5], hp[7], hp[8], hp[9], hp[13], hp[14] = 1e-4, 0.9, 1.0 / N, 0.9, 0.999, 1e-8, 0.1, 1.0
        hyper = hp.to(dev)
        ext.hyper_tick(hyper.data_ptr(), st.cuda_stream)
        algo = ext.ALGO_NVLS if (arena.has_multicast and N > 1) else ext.ALGO_TWOSHOT
        # bytes per GPU and direction: NVLS = S (1 + 1/N) (reduce-scatter: the switch reads every rank's whole message once, S out /
        # S/N in; all-gather: S/N out / S in — confirmed by the link counters), two-shot P2P = 2 (N-1)/N S
        nv_factor = (1.0 + 1.0 / N) if algo == ext.ALGO_NVLS else 2.0 * (N - 1) / max(N, 1)
        lanes = 8
        per_rank = ((n // lanes + N - 1) // N) * lanes
        master = torch.zeros(per_rank + 4096, device=dev)
        mom = torch.zeros(per_rank + 4096, device=dev)
        bucket = 8 << 20
        offs = list(range(0, S, bucket))

        def ssgd():
            moff = 0
            for o in offs:
                nb = min(bucket, S - o)
                blocks = coll.pick_blocks(nb, N, True, a.blocks)
                shard = ((nb // 16 + N - 1) // N) * 8
                if N > 1:
                    ext.barrier(mesh, 0, st.cuda_stream)
                ext.ssgd_fused(mesh, 0, algo, g_off + o, p_off + o, BF16, nb, master.data_ptr() + 4 * moff, mom.data_ptr() + 4 * moff, 0,
                               hyper.data_ptr(), ext.OPT_SGD, 0, ext.FLAG_GATED if N > 1 else 0, blocks, 512, st.cuda_stream)
                if N > 1:
                    ext.wait_done(mesh, 0, 0, N * blocks, st.cuda_stream)
                moff += shard

        def sma_avg():
            blocks = coll.pick_blocks(S, N, True, a.blocks)
            ext.barrier(mesh, 0, st.cuda_stream)
            ext.allreduce_to(mesh, 0, algo, p_off, x_off, S, BF16, dtypes.SUM, 1.0 / N, 0, ext.FLAG_SCALE | ext.FLAG_GATED, blocks, 512, st.cuda_stream)
            ext.wait_done(mesh, 0, 0, N * blocks, st.cuda_stream)

        s1_full = torch.zeros(n, device=dev)
        s2_full = torch.zeros(n, device=dev)
        m_full = params.float().clone()
        xs = arena.view(x_off, n, torch.bfloat16)

        def sma_blend():
            ext.sma_blend(params.data_ptr(), xs.data_ptr(), BF16, S, grads.data_ptr(), BF16, s1_full.data_ptr(), 0, hyper.data_ptr(), ext.OPT_SGD,
                          4 * plane.sm_count, 512, st.cuda_stream)

        done = torch.zeros(4, dtype=torch.int32, device=dev)
        ok = torch.zeros(4, dtype=torch.int32, device=dev)
        pulled = torch.empty(S, dtype=torch.uint8, device=dev)
        tgt = torch.full((1,), (me + 1) % N, dtype=torch.int32, device=dev)
        ext.publish(mesh, 1, pubs, ver, params.data_ptr(), S, done.data_ptr(), 64, 512, st.cuda_stream)
        torch.cuda.synchronize()
        coll.barrier()
        pull_blocks = int(os.environ.get("KUNGFU_PULL_BLOCKS", "24"))

        def pull(nb=pull_blocks):
            ext.pull_tma(mesh, 1, 0, pubs, ver, pulled.data_ptr(), S, ok.data_ptr(), 0, nb, st.cuda_stream, tgt.data_ptr())

        def pair_step():
            ext.pair_avg(mesh, 1, 0, pubs, ver, params.data_ptr(), BF16, S, grads.data_ptr(), BF16, s1_full.data_ptr(), s2_full.data_ptr(),
                         hyper.data_ptr(), ext.OPT_ADAM, 1, done.data_ptr(), pulled.data_ptr(), ok.data_ptr(), 148, 512, st.cuda_stream, tgt.data_ptr(),
                         m_full.data_ptr())

        def flat_adam():  # the plain local AdamW step the pair step is compared with
            ext.flat_opt(params.data_ptr(), BF16, grads.data_ptr(), BF16, m_full.data_ptr(), s1_full.data_ptr(), s2_full.data_ptr(), hyper.data_ptr(),
                         ext.OPT_ADAM, n, st.cuda_stream)

        def add(path, us, nvlink_bytes=0, hbm_bytes=0, note=""):
            row = {"model": model, "params": n, "path": path, "us": round(us, 1), "note": note, **last_links}
            if nvlink_bytes:
                gbs = nvlink_bytes / (us * 1e-6) / 1e9
                row.update({"nvlink_bytes_per_dir": int(nvlink_bytes), "nvlink_gbs": round(gbs, 1), "frac_of_770_measured": round(gbs / 770, 3),
                            "frac_of_900_nominal": round(gbs / 900, 3)})
            if hbm_bytes:
                gbs = hbm_bytes / (us * 1e-6) / 1e9
                row.update({"hbm_bytes": int(hbm_bytes), "hbm_gbs": round(gbs, 1), f"frac_of_hbm_{hbm_src}_{hbm:.0f}": round(gbs / hbm, 3)})
            rows.append(row)
            if me == 0:
                print(json.dumps(row), flush=True)

        if N > 1:
            # ssgd: NVLink S*nv_factor; HBM per rank: read S grads + shard master/mom rw (16 B/N per param) + write S params
            add("ssgd(gate+fused+wait, 8MiB buckets)", timeit(ssgd, a.iters), nvlink_bytes=S * nv_factor, note=f"algo={'nvls' if algo == ext.ALGO_NVLS else 'twoshot'} blocks<={a.blocks}")
            add("sma_avg(gate+allreduce_to+wait)", timeit(sma_avg, a.iters), nvlink_bytes=S * nv_factor, note=f"blocks<={a.blocks}")
            for nb in (8, 24, 74, 148):
                add(f"pull_tma(blocks={nb})", timeit(lambda nb=nb: pull(nb), a.iters), nvlink_bytes=S, note="every rank pulls a different peer concurrently")
        add("sma_blend(local, momentum-SGD)", timeit(sma_blend, a.iters), hbm_bytes=n * (2 + 2 + 2 + 4 + 2 + 4), note="r p,avg,g,m; w p,m")
        add("pair_step(local avg + AdamW + publish)", timeit(pair_step, a.iters), hbm_bytes=n * (2 + 2 + 4 + 4 + 4 + 2 + 2 + 4 + 4 + 4),
            note="r other,g,master,m,v; w p,pub,master,m,v")
        add("flat_adamw(local, for comparison)", timeit(flat_adam, a.iters), hbm_bytes=n * (2 + 4 + 4 + 4 + 2 + 4 + 4 + 4), note="r g,master,m,v; w p,master,m,v")
        torch.cuda.synchronize()
        arena.check()
        del arena
    clocks = sampler.stop() if sampler is not None else None
    if me == 0:
        out = {"bench": "fused_paths", "n_gpus": N, "hbm_peak_gbs": hbm, "hbm_peak_source": hbm_src, "nvlink_measured_peer_copy_gbs": 770,
               "nvlink_nominal_gbs": 900, "rows": rows, "clocks": clocks}
        if a.out:
            os.makedirs(os.path.dirname(os.path.abspath(a.out)), exist_ok=True)
            with open(a.out, "w") as f:
                json.dump(out, f, indent=1)
        print("RESULT: " + json.dumps({"n_gpus": N, "paths": len(rows), "clocks": clocks}))
    coll.barrier()


if __name__ == "__main__":
    main()
