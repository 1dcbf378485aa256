"""GPU elastic worker (run under `kungfu-run -w` on a multi-GPU box): the fused S-SGD engine keeps training across
1 -> 2 -> 1 resizes. Every rank feeds the SAME batch, so the averaged gradient — and therefore the whole trajectory —
is independent of the world size and can be checked against a plain torch.optim reference: this validates that fp32
master weights and momentum survive the re-sharding, that newcomers inherit the state, and that every symmetric arena
re-discovers its peers (VMM re-import + multicast re-bind)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import kungfu_b200.python as kfp  # noqa: E402
import kungfu_b200.torch as kf  # noqa: E402
from kungfu_b200 import ops  # noqa: E402
from kungfu_b200.elastic import ElasticState, sync_after_resize  # noqa: E402
from kungfu_b200.ops.adapt import StepBasedSchedule  # noqa: E402


def mlp():
    torch.manual_seed(5)
    return torch.nn.Sequential(torch.nn.Linear(32, 64), torch.nn.Tanh(), torch.nn.Linear(64, 8)).cuda()


def batch(step):
    g = torch.Generator(device="cuda").manual_seed(1000 + step)
    return torch.randn(16, 32, device="cuda", generator=g), torch.randint(0, 8, (16,), device="cuda", generator=g)


def main():
    schedule = StepBasedSchedule(sys.argv[1] if len(sys.argv) > 1 else "1:4,2:4,1:4")
    torch.cuda.set_device(kf.get_cuda_index())
    m = mlp()
    opt = kf.optimizers.SynchronousSGDOptimizer(torch.optim.SGD(m.parameters(), lr=0.05, momentum=0.9), named_parameters=m.named_parameters(),
                                                bucket_mb=0.004)
    assert opt.engine is not None and opt.engine.mode == "ssgd"
    state = sync_after_resize(ElasticState(), m, opt)
    sizes = []
    while state.step < schedule.total_steps:
        x, y = batch(state.step)
        opt.zero_grad()
        torch.nn.functional.cross_entropy(m(x), y).backward()
        opt.step()
        torch.cuda.synchronize()
        opt.engine.check()
        state.step += 1
        sizes.append(kf.current_cluster_size())
        assert all(ops.consensus(p.data) for p in m.parameters()), f"replicas diverged at step {state.step}"
        want = schedule(state.step) if state.step < schedule.total_steps else kf.current_cluster_size()
        if want != kf.current_cluster_size():
            changed, detached = kfp.resize(want)
            if detached:
                print(json.dumps({"detached": True, "step": state.step}), flush=True)
                return
            if changed:
                state = sync_after_resize(state, m, opt)
    if kf.current_rank() == 0:
        ref = mlp()
        ro = torch.optim.SGD(ref.parameters(), lr=0.05, momentum=0.9)
        for s in range(schedule.total_steps):
            x, y = batch(s)
            ro.zero_grad()
            torch.nn.functional.cross_entropy(ref(x), y).backward()
            ro.step()
        diff = max((a - b).abs().max().item() for a, b in zip(ref.parameters(), m.parameters()))
        mesh = opt.engine.arena.mesh
        print(json.dumps({"gpu_elastic": True, "steps": state.step, "max_diff_vs_reference": diff, "ok": diff < 5e-5, "mesh_epoch": mesh.epoch,
                          "multicast": bool(mesh.has_multicast), "sizes_seen": sorted(set(sizes))}), flush=True)


if __name__ == "__main__":
    main()
