echo "=== tests"; timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q 2>&1 | tail -3
echo "=== bench1"; timeout 400 python bench.py --steps 30 --warmup 5 2>&1 | grep "^{" | tail -1
echo "=== cupti"; timeout 400 python benchmarks/profile_step.py --out gpurun_out/resnet50_step_cupti.md 2>&1 | grep -i "total kernel" | cut -c1-150
