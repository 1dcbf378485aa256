"""Real (CUPTI, not replayed) per-kernel device time of one ResNet-50 training step: which kernels own the step.

    python benchmarks/profile_step.py [--batch 256] [--out profiles/resnet50_step_cupti.md]
"""
import argparse
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import kungfu_b200.torch as kf  # noqa: E402
from kungfu_b200.models import resnet  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--out", default="")
    ap.add_argument("--shapes", action="store_true")
    a = ap.parse_args()
    torch.cuda.set_device(0)
    torch.backends.cudnn.benchmark = True
    torch.manual_seed(0)
    model = resnet.to_b200(resnet.resnet50(), torch.bfloat16, "cuda")
    opt = kf.optimizers.SynchronousSGDOptimizer(torch.optim.SGD(model.parameters(), lr=0.01), named_parameters=model.named_parameters())
    x = torch.randn(a.batch, 3, 224, 224, device="cuda").bfloat16().contiguous(memory_format=torch.channels_last)
    y = torch.randint(0, 1000, (a.batch,), device="cuda")

    def step():
        opt.zero_grad()
        F.cross_entropy(model(x).float(), y).backward()
        opt.step()

    for _ in range(5):
        step()
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile

    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU], record_shapes=a.shapes) as prof:
        for _ in range(a.steps):
            step()
        torch.cuda.synchronize()
    evs = [e for e in prof.key_averages(group_by_input_shape=a.shapes) if e.device_time_total > 0 and e.device_type.name == "CUDA"]
    if not evs:
        evs = [e for e in prof.key_averages() if getattr(e, "device_time_total", 0) > 0]
    total = sum(e.self_device_time_total for e in evs)
    evs.sort(key=lambda e: -e.self_device_time_total)
    lines = [f"# ResNet-50 bf16 step (batch {a.batch}, 1 x B200): kernel device time from CUPTI (torch.profiler), {a.steps} eager steps\n",
             f"total kernel time per step: {total / a.steps / 1e3:.2f} ms\n", "| ms/step | share | calls/step | kernel |", "|---:|---:|---:|---|"]
    for e in evs[:32]:
        lines.append(f"| {e.self_device_time_total / a.steps / 1e3:.3f} | {100 * e.self_device_time_total / total:.1f}% | {e.count / a.steps:.0f} | `{e.key[:110]}` |")
    if a.shapes:
        lines.append("\n## aten::copy_ / contiguous calls by input shape (layout conversions)\n")
        for e in sorted(prof.key_averages(group_by_input_shape=True), key=lambda e: -e.device_time_total):
            if e.key in ("aten::copy_", "aten::contiguous", "aten::clone", "aten::add_", "aten::add") and e.device_time_total > 50 * a.steps:
                lines.append(f"- {e.key} {e.input_shapes} x{e.count / a.steps:.0f}/step: {e.device_time_total / a.steps / 1e3:.3f} ms/step")
    ours = sum(e.self_device_time_total for e in evs if "kf::" in e.key)
    lines.append(f"\nkungfu_b200 kernels (`kf::`): {100 * ours / total:.1f}% of kernel time")
    text = "\n".join(lines)
    print(text)
    if a.out:
        os.makedirs(os.path.dirname(os.path.abspath(a.out)), exist_ok=True)
        open(a.out, "w").write(text + "\n")


if __name__ == "__main__":
    main()
