# 1-GPU check: kernel tests, smoke, headline bench, CUPTI breakdown (one gpurun call)
mkdir -p gpurun_out
echo "=== tests"; timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q 2>&1 | tail -4
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
echo "=== bench1"; timeout 400 python bench.py --steps 30 --warmup 5 2>&1 | grep "^{" | tail -1
echo "=== cupti"; timeout 400 python benchmarks/profile_step.py --out gpurun_out/resnet50_step_cupti.md 2>&1 | grep -v Warn | head -24 | cut -c1-150
