"""BERT-large + PairAveragingOptimizer (BASELINE.json config #4): sequences/sec of the whole job, the exposed cost of the
gossip, torn reads, and the pull's achieved NVLink read bandwidth against its roofline.

The step (bf16 model, fp32 master weights + AdamW state inside our kernel, fused residual+LayerNorm / bias+GELU kernels, flash
attention and GEMMs from the libraries) is captured into ONE CUDA graph; the target of each step's pull is a device word
refreshed before every replay.

  --impl ours    zero_grad() starts a TMA pull of the target's published model on the side stream (hidden under forward+backward);
                 after backward ONE local kernel averages, applies AdamW and publishes (triple-buffered store, version words)
  --impl nccl    the same model / kernels / graph with an NCCL gossip: every step rank r sends its model to r+k and receives the
                 model of r-k with ncclSend/ncclRecv (k cycles through 1..N-1), then the same local average+AdamW kernel
  --impl local   no communication at all (upper bound: what the gossip costs is `local` minus the arm)

    torchrun --nproc-per-node 8 benchmarks/bert_pairavg.py [--model large|base|tiny] [--batch 32] [--seq 128] [--out f.json]
"""
import argparse
import json
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import kungfu_b200.torch as kf  # noqa: E402
from kungfu_b200.models import bert  # noqa: E402
from kungfu_b200.runtime import get_runtime  # noqa: E402
from kungfu_b200.trainer import GraphedStep  # noqa: E402
from kungfu_b200.utils.timing import ClockSampler, max_over_ranks  # noqa: E402


class NcclGossip:
    """Baseline arm: synchronous pairwise model exchange over ncclSend/ncclRecv, then our local average+AdamW kernel."""

    def __init__(self, model, lr):
        import torch.distributed as dist

        self.dist = dist
        self.rt = get_runtime()
        # reuse the engine for the flat parameter arena, fp32 master / AdamW state and the local kernel; its own gossip is unused
        self.opt = kf.optimizers.PairAveragingOptimizer(torch.optim.AdamW(model.parameters(), lr=lr), named_parameters=model.named_parameters(),
                                                        overlap=False)
        self.eng = self.opt.engine
        self.k = 0
        s = self.eng.sets[0]
        self.recv = torch.empty_like(s.params)
        self.send = torch.empty_like(s.params)  # NCCL sends from an ordinary allocation (the parameters live in VMM-mapped arena memory)
        self.ok = torch.tensor([1, 0, 0, 0], dtype=torch.int32, device=s.params.device)  # "version 1, clean": average it in

    def zero_grad(self):
        self.eng.zero_grad()

    @torch.no_grad()
    def step(self):
        eng, dist, rt = self.eng, self.dist, self.rt
        eng._open_step()
        s = eng.sets[0]
        N = rt.size
        self.k = self.k % (N - 1) + 1
        self.send.copy_(s.params)
        ops = [dist.P2POp(dist.isend, self.send, (rt.rank + self.k) % N), dist.P2POp(dist.irecv, self.recv, (rt.rank - self.k) % N)]
        for w in dist.batch_isend_irecv(ops):
            w.wait()
        from kungfu_b200.utils import dtypes

        cur = torch.cuda.current_stream()
        tick = torch.cuda.Event()
        tick.record(eng.plane.comm_stream)
        cur.wait_event(tick)
        ptr = lambda t: t.data_ptr() if t is not None else 0
        lay = s.layout
        eng.ext.pair_avg(eng.arena.mesh, 1, rt.rank, [s.x_off, s.x_off2, s.x_off3], s.ver_off, s.params.data_ptr(), dtypes.dtype_code(lay.dtype),
                         lay.nbytes, s.grads.data_ptr(), dtypes.dtype_code(lay.dtype), ptr(s.s1[0]), ptr(s.s2[0]), s.hyper.data_ptr(),
                         eng.ext.OPT_ADAM, 1, s.done_counter.data_ptr(), self.recv.data_ptr(), self.ok.data_ptr(), 148, 512, cur.cuda_stream, 0,
                         ptr(s.master[0]))
        eng.finish()

    @property
    def engine(self):
        return self.eng


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="large")
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--seq", type=int, default=128)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "nccl", "local"])
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--eager-ops", action="store_true", help="PyTorch LayerNorm / GELU instead of the fused kernels")
    ap.add_argument("--out", default="")
    ap.add_argument("--breakdown", default="", help="rank 0: CUPTI kernel-time breakdown of 3 steps -> this markdown file")
    a = ap.parse_args()
    rt = get_runtime()
    N = rt.size
    torch.cuda.set_device(rt.cuda_index())
    dev = torch.device("cuda", rt.cuda_index())
    torch.manual_seed(rt.rank)  # replicas start different: gossip has work to do
    cfg = getattr(bert.BertConfig, a.model)()
    cfg.fused_ops = not a.eager_ops
    model = bert.Bert(cfg).to(dev).bfloat16()
    lr = 1e-4
    if a.impl == "nccl" and N > 1:
        import torch.distributed as dist

        os.environ.setdefault("TORCH_NCCL_ASYNC_ERROR_HANDLING", "0")
        if not dist.is_initialized():
            dist.init_process_group("nccl", device_id=dev)
        dist.all_reduce(torch.ones(1 << 16, device=dev))
        opt = NcclGossip(model, lr)
    elif a.impl == "local" or N == 1:
        if N > 1:
            # nobody to talk to: the Python bucket path honours the nranks override below (the native engine asks the mesh); it hands the
            # gradients over with one batched copy per bucket like the native one, and runs the same fused AdamW kernel
            os.environ["KUNGFU_NATIVE_ENGINE"] = "0"
        opt = kf.optimizers.SynchronousSGDOptimizer(torch.optim.AdamW(model.parameters(), lr=lr), named_parameters=model.named_parameters())
        if N > 1:
            opt.engine.nranks = 1
    else:
        opt = kf.optimizers.PairAveragingOptimizer(torch.optim.AdamW(model.parameters(), lr=lr), named_parameters=model.named_parameters())
    ids = torch.randint(0, cfg.vocab_size, (a.batch, a.seq), device=dev)
    y = torch.randint(0, cfg.num_labels, (a.batch,), device=dev)
    loss_buf = torch.zeros((), device=dev)

    def body():
        opt.zero_grad()
        loss = F.cross_entropy(model(ids).float(), y)
        loss.backward()
        opt.step()
        loss_buf.copy_(loss.detach())

    step = GraphedStep(body, opt, warmup_iters=max(3, a.warmup), use_graph=not a.no_graph,
                       capture_error_mode="thread_local" if a.impl == "nccl" else "global")
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    sampler = ClockSampler(rt.cuda_index()).start() if rt.rank == 0 else None
    rt.control.barrier()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(a.steps):
        step()
    t1.record()
    torch.cuda.synchronize()
    ms = max_over_ranks(t0.elapsed_time(t1)) / a.steps
    clocks = sampler.stop() if sampler is not None else None
    eng = opt.engine
    eng.check()
    torn = eng.torn_reads() if a.impl == "ours" and N > 1 else 0
    torn_total = int(max_over_ranks(float(torn)))
    nbytes = sum(s.layout.nbytes for s in eng.sets)
    pull_ms = None
    if a.impl == "ours" and N > 1:
        # the pull alone, all ranks at once (every rank reads a different peer): achieved NVLink read bandwidth
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        rt.control.barrier()
        eng.set_pair_target((rt.rank + 1) % N)
        ss = eng.plane.side_stream
        torch.cuda.synchronize()
        g0.record(ss)
        for _ in range(5):
            for s in eng.sets:
                eng.ext.pull_tma(eng.arena.mesh, 1, 0, [s.x_off, s.x_off2, s.x_off3], s.ver_off, s.pulled.data_ptr(), s.layout.nbytes,
                                 s.pull_ok.data_ptr(), 0, int(os.environ.get("KUNGFU_PULL_BLOCKS", "24")), ss.cuda_stream, eng._target_dev.data_ptr())
        g1.record(ss)
        torch.cuda.synchronize()
        pull_ms = max_over_ranks(g0.elapsed_time(g1)) / 5
    if a.breakdown:
        from torch.profiler import ProfilerActivity, profile

        torch.cuda.synchronize()
        rt.control.barrier()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for _ in range(3):
                step()
            torch.cuda.synchronize()
        rt.control.barrier()
        if rt.rank == 0:
            evs = [e for e in prof.key_averages() if getattr(e, "self_device_time_total", 0) > 0]
            tot = sum(e.self_device_time_total for e in evs)
            evs.sort(key=lambda e: -e.self_device_time_total)
            lines = [f"# BERT-{a.model} step (batch {a.batch} x seq {a.seq}, impl {a.impl}, {N} GPUs): kernel device time from CUPTI, 3 graph replays\n",
                     f"sum of kernel time per step: {tot / 3 / 1e3:.2f} ms (streams overlap: the step itself is {ms:.2f} ms)\n",
                     "| ms/step | share | calls/step | kernel |", "|---:|---:|---:|---|"]
            for e in evs[:28]:
                lines.append(f"| {e.self_device_time_total / 3 / 1e3:.3f} | {100 * e.self_device_time_total / tot:.1f}% | {e.count / 3:.0f} | `{e.key[:100]}` |")
            ours = sum(e.self_device_time_total for e in evs if "kf::" in e.key)
            lines.append(f"\nkungfu_b200 kernels (`kf::`): {100 * ours / tot:.1f}% of kernel time")
            os.makedirs(os.path.dirname(os.path.abspath(a.breakdown)), exist_ok=True)
            open(a.breakdown, "w").write("\n".join(lines) + "\n")
    if rt.rank == 0:
        out = {"metric": "BERT PairAveraging sequences/sec", "value": round(N * a.batch / (ms / 1e3), 1), "unit": "sequences/sec", "impl": a.impl,
               "n_gpus": N, "model": a.model, "params": sum(p.numel() for p in model.parameters()), "dtype": "bf16 (fp32 master + AdamW state)",
               "batch_per_gpu": a.batch, "seq": a.seq, "steps": a.steps, "ms_per_step": round(ms, 3), "cuda_graph": step.graph is not None,
               "graph_error": step.graph_error, "fused_ops": cfg.fused_ops, "model_bytes": nbytes, "torn_reads_max_over_ranks": torn_total,
               "loss": round(float(loss_buf.item()), 4), "clocks": clocks, "data": "synthetic"}
        if pull_ms is not None:
            out.update({"pull_ms_isolated": round(pull_ms, 3), "pull_nvlink_read_gbs": round(nbytes / (pull_ms * 1e-3) / 1e9, 1),
                        "pull_frac_of_770_measured": round(nbytes / (pull_ms * 1e-3) / 1e9 / 770, 3)})
        print("RESULT: " + json.dumps(out), flush=True)
        if a.out:
            os.makedirs(os.path.dirname(os.path.abspath(a.out)), exist_ok=True)
            with open(a.out, "w") as f:
                json.dump(out, f, indent=1)
    torch.cuda.synchronize()
    rt.control.barrier()


if __name__ == "__main__":
    main()
