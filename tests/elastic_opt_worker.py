"""Elastic schedule with each distributed optimizer (CPU tensors): survivors and joiners must stay in step through resizes.
Usage under `kungfu-run -w`: elastic_opt_worker.py <sma|pair|ada|gns|variance> <schedule> <max_step>"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import kungfu_b200.python as kfp  # noqa: E402
import kungfu_b200.torch as kf  # noqa: E402
from kungfu_b200.elastic import KungFuElasticTrainHook  # noqa: E402

which, schedule, max_step = sys.argv[1], sys.argv[2], int(sys.argv[3])
m = torch.nn.Linear(4, 2)
if which == "elastichook":
    # sample-count driven loop (reference: experimental/hook/elastic.py ElasticHook): resize at given steps, stop after 2 x 48 samples
    from kungfu_b200.elastic import ElasticHook

    opt = kf.optimizers.SynchronousSGDOptimizer(torch.optim.SGD(m.parameters(), lr=0.01), named_parameters=m.named_parameters())
    hook = ElasticHook(local_batch_size=4, epochs=2, epoch_size=48, model=m, optimizer=opt)
    hook.begin()
    plan = {int(k): int(v) for k, v in (kv.split("=") for kv in schedule.split(","))}
    sizes = set()
    while True:
        opt.zero_grad()
        m(torch.randn(4, 4)).pow(2).mean().backward()
        opt.step()
        sizes.add(kfp.current_cluster_size())
        if not hook.after_step(plan.get(hook.state.step + 1)):
            break
    if kfp.detached():
        print(json.dumps({"detached": True}), flush=True)
    elif kfp.current_rank() == 0:
        print(json.dumps({"done": which, "trained": hook.state.trained_samples, "steps": hook.state.step, "sizes": sorted(sizes)}), flush=True)
    sys.exit(0)
if which == "shrink":
    # shrink-and-continue: `schedule` = "<rank that dies>@<step>[,<rank>@<step>...]"; the survivors must finish `max_step` steps with identical
    # models. Ranks are those of the cluster at the time of the crash.
    from kungfu_b200.elastic import ElasticState, ShrinkOnFailure, sync_after_resize

    if os.environ.get("KUNGFU_TEST_JOINER_CRASH") == "1" and kfp.current_cluster_version() > 0:
        os._exit(3)  # a replacement worker that never makes it into the job
    torch.manual_seed(1234 + kfp.current_rank())
    dev = torch.device("cpu")
    if os.environ.get("KUNGFU_TEST_CUDA") == "1":  # the fused GPU engine: bf16 model, sharded fp32 master weights + momentum
        from kungfu_b200.runtime import get_runtime as _grt

        dev = torch.device("cuda", _grt().cuda_index())
        torch.cuda.set_device(dev)
        m = torch.nn.Sequential(torch.nn.Linear(4, 512), torch.nn.ReLU(), torch.nn.Linear(512, 2)).to(dev).bfloat16()
    crashes = [tuple(int(x) for x in c.split("@")) for c in schedule.split(",")]
    opt = kf.optimizers.SynchronousSGDOptimizer(torch.optim.SGD(m.parameters(), lr=0.05, momentum=0.9), named_parameters=m.named_parameters())
    state = ElasticState()
    sync_after_resize(state, m, opt)
    refill = os.environ.get("KUNGFU_TEST_REFILL") == "1"  # follow the config server: the launcher proposes a replacement after a loss
    guard = ShrinkOnFailure(state, m, opt, timeout_s=20, follow_config_server=refill, poll_every=2)
    sizes = []

    hang = os.environ.get("KUNGFU_TEST_HANG") == "1"  # the faulty worker hangs inside a batch instead of dying
    if hang:
        from kungfu_b200.cmd import monitor_batch_begin, monitor_batch_end

    def one_step():
        if hang:
            monitor_batch_begin()
        if (kfp.current_rank(), state.step) in crashes and len(guard.failures) == sum(1 for r, t in crashes if t < state.step):
            if hang:
                import time

                time.sleep(3600)  # never comes back: only the launcher's hang watch can end this
            os._exit(17)  # dies mid-training without saying goodbye
        opt.zero_grad()
        m(torch.randn(3, 4, device=dev, dtype=next(m.parameters()).dtype)).float().pow(2).mean().backward()
        opt.step()
        if dev.type == "cuda":
            torch.cuda.synchronize()
        if refill:
            import time

            time.sleep(0.1)  # a step of a real model takes time; the launcher proposes the replacement about a second after the loss
        if hang:
            monitor_batch_end()

    while state.step < max_step and not guard.detached:
        if guard.run(one_step):
            state.step += 1
            sizes.append(kfp.current_cluster_size())
    # every survivor holds the same model: compare a checksum through the (new) session
    flat = torch.cat([p.detach().reshape(-1) for p in m.parameters()]).double().cpu()
    from kungfu_b200.runtime import get_runtime

    ctl = get_runtime().control
    lo = ctl.all_reduce_cpu_(flat.clone(), "min", "shrink::lo")
    hi = ctl.all_reduce_cpu_(flat.clone(), "max", "shrink::hi")
    print(json.dumps({"done": which, "rank": kfp.current_rank(), "step": state.step, "final_size": kfp.current_cluster_size(),
                      "sizes": sorted(set(sizes)), "failures": [(a, b) for a, b, _ in guard.failures], "regrown": [(a, b) for a, b, _ in guard.regrown],
                      "recover_s": [round(t, 3) for _, _, t in guard.failures], "identical": bool(torch.equal(lo, hi)),
                      "finite": bool(torch.isfinite(flat).all())}), flush=True)
    sys.exit(0)
mk = {
    "sma": lambda o, n: kf.optimizers.SynchronousAveragingOptimizer(o, named_parameters=n),
    "pair": lambda o, n: kf.optimizers.PairAveragingOptimizer(o, named_parameters=n),
    "ada": lambda o, n: kf.optimizers.AdaptiveSGDOptimizer(o, n, change_step=3),
    "gns": lambda o, n: kf.optimizers.MonitorGradientNoiseScaleOptimizer(o, n, device_batch_size=3),
    "variance": lambda o, n: kf.optimizers.MonitorGradientVarianceOptimizer(o, n),
}[which]
opt = mk(torch.optim.SGD(m.parameters(), lr=0.01), list(m.named_parameters()))
hook = KungFuElasticTrainHook(schedule, max_step, m, opt)
hook.begin()
sizes = set()
while True:
    opt.zero_grad()
    m(torch.randn(3, 4)).pow(2).mean().backward()
    opt.step()
    sizes.add(kfp.current_cluster_size())
    if not hook.after_step():
        break
if kfp.detached():
    print(json.dumps({"detached": True}), flush=True)
elif kfp.current_rank() == 0:
    print(json.dumps({"done": which, "step": hook.state.step, "sizes": sorted(sizes),
                      "finite": bool(all(torch.isfinite(p).all() for p in m.parameters()))}), flush=True)
