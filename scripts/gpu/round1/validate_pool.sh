echo "=== tests"; timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "maxpool or resnet or trainer or bn" 2>&1 | tail -3
echo "=== nan"; timeout 120 python - <<'PY'
import torch
from kungfu_b200.ops.pool import max_pool3x3s2
x = torch.randn(2, 16, 9, 11, device="cuda").bfloat16()
x[0, 3, 4, 5] = float("nan"); x[1, :, 0, 0] = float("-inf")
x = x.contiguous(memory_format=torch.channels_last).requires_grad_()
y = max_pool3x3s2(x); dy = torch.ones_like(y); y.backward(dy)
xr = x.detach().clone().requires_grad_(); yr = torch.nn.functional.max_pool2d(xr, 3, 2, 1); yr.backward(dy)
same_nan = torch.equal(torch.isnan(y), torch.isnan(yr)); ok = torch.equal(torch.nan_to_num(y.float(), 7.0), torch.nan_to_num(yr.float(), 7.0))
print("nan fwd ok:", same_nan and ok, "grad maxdiff:", (x.grad.float() - xr.grad.float()).abs().max().item())
PY
echo "=== bench1"; timeout 400 python bench.py --steps 30 --warmup 5 2>&1 | grep "^{" | tail -1
echo "=== cupti"; timeout 400 python benchmarks/profile_step.py --out gpurun_out/resnet50_step_cupti.md 2>&1 | grep -i "maxpool\|total kernel" | cut -c1-150
