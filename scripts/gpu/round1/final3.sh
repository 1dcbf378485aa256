timeout 50 python - <<'PY'
import torch, time
t0=time.time()
import __graft_entry__ as g
g.smoke()
import kungfu_b200.torch as kf
from kungfu_b200 import variables as kv, ops
m = torch.nn.Sequential(torch.nn.Linear(64, 64), torch.nn.ReLU(), torch.nn.Linear(64, 8)).cuda()
opt = kf.optimizers.MonitorGradientNoiseScaleOptimizer(torch.optim.SGD(m.parameters(), lr=0.1), m.named_parameters(), device_batch_size=16)
for _ in range(3):
    opt.zero_grad(); m(torch.randn(16, 64, device="cuda")).pow(2).mean().backward(); opt.step()
print("gns", opt.noise_scale(), kv.eval_gradient_noise_scale(), "allreduce", float(ops.all_reduce(torch.ones(4, device="cuda")).sum()))
print("final smoke ok in %.1fs" % (time.time()-t0))
PY
