"""BERT-large + PairAveragingOptimizer (BASELINE.json config #4): sequences/sec and the gossip kernel's achieved NVLink
read bandwidth against its roofline (bytes of the published model / 770 GB/s measured peer bandwidth).

    torchrun --nproc-per-node 8 benchmarks/bert_pairavg.py [--model large|base|tiny] [--batch 32] [--seq 128]
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
from kungfu_b200.utils.timing import max_over_ranks  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="large")
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--seq", type=int, default=128)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    a = ap.parse_args()
    rt = get_runtime()
    torch.cuda.set_device(rt.cuda_index())
    torch.manual_seed(rt.rank)  # replicas start different: gossip has work to do
    cfg = getattr(bert.BertConfig, a.model)()
    model = bert.Bert(cfg).cuda()  # fp32 parameters (the gossip blob is the fp32 model, 1.34 GB for BERT-large); bf16 autocast compute
    opt = kf.optimizers.PairAveragingOptimizer(torch.optim.AdamW(model.parameters(), lr=1e-4), named_parameters=model.named_parameters())
    ids = torch.randint(0, cfg.vocab_size, (a.batch, a.seq), device="cuda")
    y = torch.randint(0, cfg.num_labels, (a.batch,), device="cuda")

    def step():
        opt.zero_grad()
        with torch.autocast("cuda", dtype=torch.bfloat16):
            loss = F.cross_entropy(model(ids).float(), y)
        loss.backward()
        opt.step()

    for _ in range(a.warmup):
        step()
    torch.cuda.synchronize()
    rt.control.barrier()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(a.steps):
        step()
    t1.record()
    torch.cuda.synchronize()
    ms = max_over_ranks(t0.elapsed_time(t1)) / a.steps
    eng = opt.engine
    eng.check()
    # the gossip kernel alone
    nbytes = sum(s.layout.nbytes for s in eng.sets)
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tgt = (rt.rank + 1) % rt.size
    g0.record()
    for _ in range(5):
        eng.pair_step(tgt, do_step=True)
        eng.finish()
    g1.record()
    torch.cuda.synchronize()
    kms = max_over_ranks(g0.elapsed_time(g1)) / 5
    if rt.rank == 0:
        print("RESULT: %.1f %s" % (rt.size * a.batch / (ms / 1e3), json.dumps({
            "metric": "BERT PairAveraging sequences/sec", "np": rt.size, "model": a.model, "params": sum(p.numel() for p in model.parameters()),
            "batch": a.batch, "seq": a.seq, "ms_per_step": round(ms, 3), "pair_avg_kernel_ms": round(kms, 3), "model_bytes": nbytes,
            "pair_avg_nvlink_read_gbs": round(nbytes / (kms * 1e-3) / 1e9, 1), "roofline_gbs": 770,
            "torn_reads": eng.arena.mesh.status()[5]})))
    torch.cuda.synchronize()


if __name__ == "__main__":
    main()
