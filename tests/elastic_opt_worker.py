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
