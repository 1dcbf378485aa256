mkdir -p gpurun_out
echo "=== pytest -m gpu"; timeout 150 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
echo "=== smoke"; timeout 60 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
echo "=== vgg16"; timeout 70 python benchmarks/system/benchmark_kungfu_torch.py --model vgg16 --batch-size 64 --num-iters 2 --num-batches-per-iter 5 --num-warmup-batches 3 2>&1 | grep -E "RESULT|Error|error" | tail -2
echo "=== inception_v3"; timeout 70 python benchmarks/system/benchmark_kungfu_torch.py --model inception_v3 --batch-size 64 --num-iters 2 --num-batches-per-iter 5 --num-warmup-batches 3 2>&1 | grep -E "RESULT|Error|error" | tail -2
